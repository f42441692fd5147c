"""Time one MNLE training step (4096-row minibatch, run_config.py:12) on the device and the same
step through the CPU spec under torch autograd (fp32, all host threads)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ddm_oracle as orc  # noqa: E402
from oracle import mnle_spec as ms  # noqa: E402
from sbi_for_diffusion_models_b200.mnle_train import MNLETrainer  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    N = 1 << 18
    p = ms.init_params(0)
    theta = orc.prior_sample(N, seed=2)
    rs = np.random.RandomState(0)
    pulses = torch.from_numpy(np.where(rs.rand(N, 80) < 0.5, 1.0, -1.0).astype(np.float32))
    cond = torch.cat([theta, pulses], 1)
    x = torch.from_numpy(np.stack([np.exp(rs.uniform(-3, 2.1, N)), rs.randint(0, 3, N)], 1).astype(np.float32))
    tr = MNLETrainer(3, cond_mean=p["cond_mean"], cond_std=p["cond_std"], mu_y=0.35, sigma_y=1.1, init=p)
    xd, cd = x.cuda(), tr.standardise(cond)
    g = torch.Generator(device="cuda").manual_seed(0)
    idx = [torch.randperm(N, device="cuda", generator=g)[:R].contiguous() for _ in range(8)]
    for i in range(5):
        tr.nll(xd, cd, idx[i % 8]); tr.adam()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 50
    e0.record()
    for i in range(steps):
        tr.nll(xd, cd, idx[i % 8]); tr.adam()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / steps
    e0.record()
    for i in range(steps):
        tr.nll(xd, cd, idx[i % 8], grad=False)
    e1.record()
    torch.cuda.synchronize()
    ms_fwd = e0.elapsed_time(e1) / steps
    out = {"rows": R, "ms_per_step": ms_step, "ms_forward_only": ms_fwd, "rows_per_s": R / ms_step * 1e3,
           "dense_tflops_fwd_bwd": 3 * 0.818e6 * R / ms_step * 1e3 / 1e12, "loss": float(tr.stats[0])}
    if "--no-cpu" not in sys.argv:
        frozen = ("cond_mean", "cond_std", "flow.mu_y", "flow.sigma_y")
        pc = {k: v.clone().requires_grad_(k not in frozen) for k, v in p.items()}
        opt = torch.optim.Adam([v for k, v in pc.items() if k not in frozen], lr=5e-4)
        xs, cs = x[:R], cond[:R]
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            opt.zero_grad()
            loss = -ms.log_prob(pc, xs, cs).mean()
            loss.backward()
            torch.nn.utils.clip_grad_norm_([v for k, v in pc.items() if k not in frozen], 5.0)
            opt.step()
            ts.append(time.perf_counter() - t0)
        out["cpu_spec_autograd_ms_per_step"] = 1e3 * float(np.median(ts[1:]))
        out["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
