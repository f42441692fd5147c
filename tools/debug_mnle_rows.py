import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ddm_oracle as orc, mnle_spec as ms
from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
p = ms.init_params(0, scale=1.0); p64 = ms.cast_params(p, torch.float64)
est = DeviceMNLE(PackedMNLE.from_params(p))
R=3000
theta = orc.prior_sample(R, seed=2)
pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(1)), 0, R, 80, 0.75))
cond = torch.cat([theta, pulses], dim=1)
rs = np.random.RandomState(0)
x = torch.from_numpy(np.stack([np.exp(rs.uniform(-3, 2.1, R)), rs.randint(0, 3, R)], 1).astype(np.float32))
x[:5, 0] = torch.tensor([1e-6, 8.0, 7.999999, 1e-3, 3e-5])
got = est.log_prob(x.unsqueeze(0), condition=cond)[0].double()
w64 = ms.log_prob(p64, x, cond); w32 = ms.log_prob(p, x, cond).double()
err = (got-w64).abs()
idx = err.argsort(descending=True)[:12]
for i in idx.tolist():
    print(i, x[i].tolist(), theta[i].tolist(), "got", got[i].item(), "w64", w64[i].item(), "w32", w32[i].item(), "err", err[i].item())
print("mean err", err.mean().item(), "p99", err.quantile(0.99).item(), "n>1e-3", int((err>1e-3).sum()))
