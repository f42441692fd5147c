#!/usr/bin/env python
"""simulate_training_set_with_conditions end to end (device proposal -> CPU (z, x)) at 1e7 trials for several launch
block sizes (data_simulator._LAUNCH_ROWS): the last block's copy and host rebuild are not overlapped with anything."""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sbi_for_diffusion_models_b200 import data_simulator as ds
from sbi_for_diffusion_models_b200.priors import build_prior_theta
from sbi_for_diffusion_models_b200.proposals import ExtendedProposal, PulseSequenceProposal
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000

def once(seed, n):
    prop = ExtendedProposal(build_prior_theta(dev), PulseSequenceProposal(80, 0.75, seed=seed, device=dev), device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        return ds.simulate_training_set_with_conditions(prop, n, 1 << 18, dev, mu_sensory=1.0, p_success=0.75, P=80, log_rt=False, seed=seed)

once(0, 1 << 20); torch.cuda.synchronize()
for rows in (1 << 22, 1 << 21, 1 << 20, 1 << 19, 1 << 22):
    ds._LAUNCH_ROWS = rows
    ts = []
    for i in range(3):
        t0 = time.perf_counter(); z, x = once(1 + i, N); ts.append(time.perf_counter() - t0); del z, x
    print(f"_LAUNCH_ROWS = 2^{rows.bit_length() - 1}: {min(ts) * 1e3:.1f} ms (runs {[round(t * 1e3, 1) for t in ts]})")
