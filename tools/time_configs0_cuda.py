#!/usr/bin/env python
"""configs[0] with the prior's parameters on the GPU, and simulate_observed_session (T = 50): wall time and cProfile."""
import contextlib, cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sbi_for_diffusion_models_b200 import data_simulator as ds
from sbi_for_diffusion_models_b200.priors import build_prior_theta
from sbi_for_diffusion_models_b200.proposals import ExtendedProposal, PulseSequenceProposal
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
prior = build_prior_theta(dev)

def once(seed):
    prop = ExtendedProposal(prior, PulseSequenceProposal(80, 0.75, seed=seed, device=dev), device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        return ds.simulate_training_set_with_conditions(prop, 10_000, 4096, dev, mu_sensory=1.0, p_success=0.75, P=80, log_rt=False, seed=seed)
def session(seed):
    return ds.simulate_observed_session(torch.tensor([0.45, 0.6, 1.3, 14.0, 0.25]), 50, dev, mu_sensory=1.0, p_success=0.75, P=80, seed=seed, log_rt=False)
for name, fn in (("configs0 (cuda prior)", once), ("simulate_observed_session T=50", session)):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(20):
        t0 = time.perf_counter(); fn(10 + i); ts.append(time.perf_counter() - t0)
    print(name, "ms: min %.3f median %.3f" % (min(ts) * 1e3, sorted(ts)[10] * 1e3))
    pr = cProfile.Profile(); pr.enable(); fn(99); pr.disable()
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(14); print(st.getvalue()[:3800])
