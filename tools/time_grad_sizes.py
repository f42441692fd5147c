#!/usr/bin/env python
"""Value + gradient of the MNLE potential: reverse mode on tcgen05 against forward mode on the CUDA cores over a
range of (T, C), graph-replayed and eager (the threshold of DeviceMNLE.GRAD_TC_MIN_ROWS)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
from sbi_for_diffusion_models_b200.simulator import simulate_trials

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mnle_trained.npz"))
est = DeviceMNLE(PackedMNLE(d["packed"], int(d["n_choices"])))
z = bench.build_workload(4096, 0, dev)
for T, C in ((1, 1), (8, 4), (50, 1), (50, 2), (50, 8), (50, 32), (50, 128), (50, 1024)):
    th = z[:C, :5].contiguous(); pl = z[:T, 5:].contiguous()
    xo = simulate_trials(z[:T, :5], pl, seed=3)
    res = {}
    for k in ("tc", "simt"):
        fn = lambda: est.loglik_sum_and_grad(th, xo, pl, kernel=k)
        g = bench._time_graph(fn, reps=5)
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); e = (time.perf_counter() - t0) / 20 * 1e3
        res[k] = (g, e)
    fwd = bench._time_graph(lambda: est.loglik_sum(th, xo, pl), reps=5)
    print(f"T={T:3d} C={C:5d}: tc graph {res['tc'][0]:.3f} eager {res['tc'][1]:.3f} | simt graph {res['simt'][0]:.3f} eager {res['simt'][1]:.3f} | forward only {fwd:.3f} ms")
