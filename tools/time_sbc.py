#!/usr/bin/env python
"""Time run_sbc on the trained estimator and count potential evaluations.
    python tools/time_sbc.py [datasets] [samples] [warmup]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sbi_for_diffusion_models_b200 import mnle, samplers
from sbi_for_diffusion_models_b200.priors import build_prior_theta
from sbi_for_diffusion_models_b200.run_config import RunConfig

D = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
W = int(sys.argv[3]) if len(sys.argv) > 3 else 100
torch.cuda.set_device(0)
est, _ = bench.trained_mnle()
prior = build_prior_theta()
evals = {"n": 0}
orig = samplers.VectorizedSliceSampler.run
def counting(self, *a, **kw):
    out = orig(self, *a, **kw)
    evals["n"] += self.n_evals
    return out
samplers.VectorizedSliceSampler.run = counting
mnle.run_sbc(RunConfig(WARMUP_STEPS=2), prior_theta=prior, density_estimator=est, num_datasets=2, posterior_samples_per_dataset=128, save=False)
torch.cuda.synchronize()
evals["n"] = 0
t0 = time.perf_counter()
out = mnle.run_sbc(RunConfig(WARMUP_STEPS=W), prior_theta=prior, density_estimator=est, num_datasets=D, posterior_samples_per_dataset=S, save=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
sweeps = W + -(-S // 128)
print(f"D={D} S={S} W={W}: {dt:.2f} s, {evals['n']} potential calls ({evals['n'] / (sweeps * 5):.1f} per coordinate update), "
      f"{dt / evals['n'] * 1e3:.3f} ms per call; rank means / S {(out['ranks'].mean(0) / S).round(3).tolist()}")
