"""Distribution of per-row |tc - simt| on the sharpened net (scale 2), T=50 x C=1024."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ddm_oracle as orc, mnle_spec as ms
from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
p = ms.init_params(1, scale=2.0)
p64 = ms.cast_params(p, torch.float64)
est = DeviceMNLE(PackedMNLE.from_params(p))
C, T = 1024, 50
theta = orc.prior_sample(C, seed=3)
pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(123)), 0, T, 80, 0.75))
x, _ = orc.sim_rng_c(np.repeat(np.array([[0.45, 0.6, 1.3, 14.0, 0.25]], np.float32), T, 0), pulses.numpy(), 7)
x = torch.from_numpy(x)
A = torch.stack([est.loglik_sum(theta, x[t:t+1], pulses[t:t+1], kernel="tc") for t in range(T)]).double()
B = torch.stack([est.loglik_sum(theta, x[t:t+1], pulses[t:t+1], kernel="simt") for t in range(T)]).double()
W = torch.stack([ms.loglik_sum(p64, theta, x[t:t+1], pulses[t:t+1]) for t in range(T)])
ea, eb = (A - W).abs(), (B - W).abs()
print("tc   vs f64: mean %.2e  p99 %.2e  p99.9 %.2e  max %.2e" % (ea.mean(), ea.flatten().quantile(0.99), ea.flatten().quantile(0.999), ea.max()))
print("simt vs f64: mean %.2e  p99 %.2e  p99.9 %.2e  max %.2e" % (eb.mean(), eb.flatten().quantile(0.99), eb.flatten().quantile(0.999), eb.max()))
idx = ea.flatten().topk(8).indices
for i in idx:
    t, c = divmod(int(i), C)
    print(f"t {t} c {c}: tc {A[t,c]:.5f} simt {B[t,c]:.5f} f64 {W[t,c]:.5f} x {x[t].tolist()} theta {[round(v,4) for v in theta[c].tolist()]}")
print("sum rel err tc  ", ((A.sum(0) - W.sum(0)).abs() / W.sum(0).abs()).max().item())
print("sum rel err simt", ((B.sum(0) - W.sum(0)).abs() / W.sum(0).abs()).max().item())
