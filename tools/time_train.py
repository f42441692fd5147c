"""Time one MNLE training step (4096-row minibatch, run_config.py:12) on cuda:0 and the same step
through the CPU spec under torch autograd (fp32, all host threads)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    torch.cuda.set_device(0)
    rows = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
    print(json.dumps(bench.mnle_train_bench(torch.device("cuda:0"), with_cpu="--no-cpu" not in sys.argv, rows=rows)))
