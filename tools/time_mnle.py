"""Time and check the MNLE potential kernels at configs[3] (T=50, C=1024) on cuda:0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    torch.cuda.set_device(0)
    print(json.dumps(bench.mnle_bench(torch.device("cuda:0"), with_cpu="--cpu" in sys.argv), indent=1))
