#!/usr/bin/env python
"""Latency of small launches: the producer / consumer kernel against the throughput kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sbi_for_diffusion_models_b200 import _native
from sbi_for_diffusion_models_b200.simulator import simulate_trials
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
L = _native.lib()
z = bench.build_workload(1 << 17, 0, dev)
for n in (50, 1000, 10000, 16384, 32768, 65536):
    res = []
    for cap in (1 << 20, 0):
        L.ddm_sim_set_small_batch_max(cap)
        zz = z[:n]
        for i in range(3): simulate_trials(zz[:, :5], zz[:, 5:], seed=i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(5): simulate_trials(zz[:, :5], zz[:, 5:], seed=10 + i)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 5)
    print(n, "small %.3f ms  throughput %.3f ms" % tuple(res))
