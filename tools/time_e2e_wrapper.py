"""Where the time of data_simulator.sim_wrapper(host z) goes: pinned allocation, pipeline, sync."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sbi_for_diffusion_models_b200 import data_simulator as ds

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
z = bench.build_workload(n, 0, dev)
zh = torch.empty((n, 85), dtype=torch.float32, pin_memory=True); zh.copy_(z); del z
torch.cuda.synchronize()
for i in range(3):
    ds.sim_wrapper(zh, mu_sensory=1.0, p_success=0.75, P=80, log_rt=False, seed=i)
x = None
ts = []
for i in range(6):
    t0 = time.perf_counter()
    x = ds.sim_wrapper(zh, mu_sensory=1.0, p_success=0.75, P=80, log_rt=False, seed=10 + i)
    ts.append((time.perf_counter() - t0) * 1e3)
print("sim_wrapper ms per call:", [round(t, 2) for t in ts])
t0 = time.perf_counter(); a = torch.empty((n, 2), dtype=torch.float32, pin_memory=True); t1 = time.perf_counter()
b = torch.empty((n, 2), dtype=torch.float32, pin_memory=True); t2 = time.perf_counter()
del a; c = torch.empty((n, 2), dtype=torch.float32, pin_memory=True); t3 = time.perf_counter()
print("pinned alloc ms: fresh %.2f fresh %.2f cached %.2f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
