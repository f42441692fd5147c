#!/usr/bin/env python
"""Time the simulator kernel of whichever library DDM_B200_LIB points at (variant sweeps).
Prints one JSON line with useful steps/s and a checksum of x (all variants must agree)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sbi_for_diffusion_models_b200.simulator import Schedule, simulate_trials

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
z = bench.build_workload(n, 0, dev)
x = torch.empty((n, 2), device=dev)
sched = Schedule.from_constants(1.0)
for i in range(2):
    simulate_trials(z[:, :5], z[:, 5:], seed=100 + i, out=x, schedule=sched)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms, steps = [], 0
for i in range(3):
    e0.record()
    _, st = simulate_trials(z[:, :5], z[:, 5:], seed=7 + i, out=x, schedule=sched, return_stats=True)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
    steps += st.useful_steps
chk = int(x.view(torch.int32).to(torch.int64).sum())
print(json.dumps({"lib": os.environ.get("DDM_B200_LIB", "default"), "trials": n, "steps_per_s": steps / (sum(ms) * 1e-3),
                  "ms": ms, "checksum": chk, "lane_eff": st.lane_efficiency}))
