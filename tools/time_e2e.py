#!/usr/bin/env python
"""Time the host-z streaming pipeline for several (chunk, streams) settings."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sbi_for_diffusion_models_b200.simulator import HostPipeline, Schedule

n = 1 << 22
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
z = bench.build_workload(n, 0, dev)
zh = torch.empty((n, 85), dtype=torch.float32, pin_memory=True); zh.copy_(z); del z
xh = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
sched = Schedule.from_constants(1.0)
t_nd = zh[:, 4].clamp(0.0, sched.t_nd_hi)
for chunk, ns in ((1 << 20, 1 << 22), (1 << 19, 1 << 22), (1 << 18, 1 << 22), (1 << 17, 1 << 22), (1 << 18, 1 << 21)):
    pipe = HostPipeline(85, chunk=chunk, max_batch=ns, device=dev)
    for i in range(2):
        pipe.run(zh, xh, sched=sched, seed=1 + i); pipe.synchronize()
    t0 = time.perf_counter()
    for i in range(4):
        pipe.run(zh, xh, sched=sched, seed=10 + i); pipe.synchronize()
    dt = (time.perf_counter() - t0) / 4
    steps = int(torch.round((xh[:, 0] - t_nd) / sched.dt).to(torch.int64).sum())
    print(json.dumps({"lib": os.path.basename(os.environ.get("DDM_B200_LIB", "default")), "chunk": chunk, "max_batch": ns,
                      "ms": dt * 1e3, "steps_per_s": steps / dt, "h2d_GBps": n * 340 / dt / 1e9}))
    del pipe
