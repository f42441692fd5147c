#!/usr/bin/env python
"""Where the time of the README-default run goes (configs[0]: 10 000 trials in batches of 4096 through
simulate_training_set_with_conditions): cProfile of the third call."""
import contextlib, cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sbi_for_diffusion_models_b200 import data_simulator as ds
from sbi_for_diffusion_models_b200.priors import build_prior_theta
from sbi_for_diffusion_models_b200.proposals import ExtendedProposal, PulseSequenceProposal
dev = torch.device("cuda", 0); torch.cuda.set_device(0)

def once(seed):
    prop = ExtendedProposal(build_prior_theta(), PulseSequenceProposal(80, 0.75, seed=seed, device=dev), device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        return ds.simulate_training_set_with_conditions(prop, 10_000, 4096, dev, mu_sensory=1.0, p_success=0.75, P=80, log_rt=False, seed=seed)
for i in range(3):
    once(i)
torch.cuda.synchronize()
ts = []
for i in range(5):
    t0 = time.perf_counter(); once(10 + i); ts.append(time.perf_counter() - t0)
print("ms per call:", [round(t * 1e3, 3) for t in ts])
pr = cProfile.Profile(); pr.enable(); once(99); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(22); print(st.getvalue()[:4500])
