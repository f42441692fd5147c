"""Per-row error of the tc and simt MNLE kernels against the float64 spec."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ddm_oracle as orc, mnle_spec as ms
from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE

for seed, scale in ((0, 1.0), (1, 2.0)):
    p = ms.init_params(seed, scale=scale)
    p64 = ms.cast_params(p, torch.float64)
    est = DeviceMNLE(PackedMNLE.from_params(p))
    theta = orc.prior_sample(700, seed=11)
    pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(123)), 0, 40, 80, 0.75))
    x, _ = orc.sim_rng_c(np.repeat(np.array([[0.45, 0.6, 1.3, 14.0, 0.25]], np.float32), 40, 0), pulses.numpy(), 7)
    x = torch.from_numpy(x)
    for t in (0, 7, 14):
        a = est.loglik_sum(theta, x[t:t + 1], pulses[t:t + 1], kernel="tc").double()
        b = est.loglik_sum(theta, x[t:t + 1], pulses[t:t + 1], kernel="simt").double()
        w = ms.loglik_sum(p64, theta, x[t:t + 1], pulses[t:t + 1])
        ea, eb = (a - w).abs(), (b - w).abs()
        i = int(ea.argmax())
        print(f"scale {scale} t {t}: tc max {ea.max():.2e} mean {ea.mean():.2e} | simt max {eb.max():.2e} mean {eb.mean():.2e} | "
              f"worst row {i}: tc {a[i]:.5f} simt {b[i]:.5f} spec {w[i]:.5f} theta {theta[i].tolist()}")
