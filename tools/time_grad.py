#!/usr/bin/env python
"""Value + gradient of the MNLE potential at configs[3] (T=50, C=1024, trained estimator), reverse mode on tcgen05:
a few calls for an ncu launch list (python tools/time_grad.py [C])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
from sbi_for_diffusion_models_b200.simulator import simulate_trials

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = 50
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mnle_trained.npz"))
est = DeviceMNLE(PackedMNLE(d["packed"], int(d["n_choices"])))
z = bench.build_workload(max(C, T), 0, dev)
th = z[:C, :5].contiguous()
pl = z[:T, 5:].contiguous()
xo = simulate_trials(z[:T, :5], pl, seed=3)
for _ in range(3):
    est.loglik_sum_and_grad(th, xo, pl, kernel="tc")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    val, grad = est.loglik_sum_and_grad(th, xo, pl, kernel="tc")
e1.record(); torch.cuda.synchronize()
print("value+grad, T=%d C=%d: %.3f ms per call" % (T, C, e0.elapsed_time(e1) / 5), float(val.sum()), float(grad.abs().sum()))
