"""Per-stage timeline of CTA 0 of the tcgen05 MNLE kernel (clock64 stamps, cycles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sbi_for_diffusion_models_b200 import _native
from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE

torch.cuda.set_device(0)
est = DeviceMNLE(PackedMNLE.from_params(bench.random_mnle_params(0)))
T, C = 50, 1024
torch.manual_seed(0)
theta = torch.rand(C, 5, device="cuda") + 0.2
x = torch.stack([torch.rand(T) * 2 + 0.3, torch.randint(0, 3, (T,)).float()], 1).cuda()
pulses = (torch.randint(0, 2, (T, 80)).float() * 2 - 1).cuda()
for _ in range(3):
    est.loglik_sum(theta, x, pulses, kernel="tc")
trace = torch.zeros(34 * 2 * 8, dtype=torch.int64, device="cuda")
_native.lib().mnle_tc_set_trace(trace.data_ptr())
est.loglik_sum(theta, x, pulses, kernel="tc")
torch.cuda.synchronize()
_native.lib().mnle_tc_set_trace(None)
tr = trace.cpu().view(34, 2, 8)
t0 = int(tr[0, 0, 0])
print("stage tile | A-ready  issued(+wts) mma-issued at-stage loaded | D-seen  epi-done | mma+hop  epilogue")
for s in range(34):
    for X in range(2):
        a, b, c, d, e, f, g, h = [int(v) - t0 for v in tr[s, X][:8]]
        print(f"{s:3d} {X} | {a:7d} {b:7d} {e:7d} {f:7d} {g:7d} | {c:7d} {d:7d} | {c - b:6d} {d - c:6d}")
print("total", int(tr.max()) - t0)
