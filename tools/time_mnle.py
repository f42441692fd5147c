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
    if "--rows" in sys.argv:
        from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
        est = DeviceMNLE(PackedMNLE.from_params(bench.random_mnle_params(0)))
        R = 1 << 20
        z = bench.build_workload(R, 0, torch.device("cuda:0"))
        x = torch.stack([torch.rand(R, device="cuda") * 3 + 0.2, torch.randint(0, 3, (R,), device="cuda").float()], 1)
        for kernel in ("tc", "simt"):
            for _ in range(2):
                est.log_prob(x, condition=z, kernel=kernel)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                est.log_prob(x, condition=z, kernel=kernel)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"rows API, R = {R}: {kernel} {ms:.3f} ms per call, {R / ms * 1e3:.3e} rows/s")

