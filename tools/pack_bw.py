"""Host packer throughput (ddm_pack_z_host) on this machine's cores, next to a plain pinned H2D copy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sbi_for_diffusion_models_b200 import _native

L = _native.lib()
N, P = 1 << 22, 80
z = torch.ones((N, 85), dtype=torch.float32).pin_memory()
out = torch.ones((N, 8), dtype=torch.int32).pin_memory()
print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
for nt in (1, 4, 8, 16, 32):
    best = 1e9
    for rep in range(3):
        t = time.perf_counter()
        L.ddm_pack_z_host(z.data_ptr(), 85, N, P, out.data_ptr(), nt)
        best = min(best, time.perf_counter() - t)
    print(f"{nt:3d} threads {best * 1e3:8.2f} ms {N * 340 / best / 1e9:7.1f} GB/s")
if torch.cuda.is_available():
    d = torch.empty_like(z, device="cuda")
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(z, non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print(f"pinned H2D {dt * 1e3:.2f} ms {N * 340 / dt / 1e9:.1f} GB/s")

# the ingest loop alone (pack chunk -> enqueue copies), no kernel running
if torch.cuda.is_available():
    import ctypes
    chunk = 1 << 18
    pk = torch.empty((N, 8), dtype=torch.int32, device="cuda")
    ready = torch.zeros(1, dtype=torch.int64, device="cuda")
    marks = (torch.arange(1, N // chunk + 1, dtype=torch.int64) * chunk).pin_memory()
    cs = torch.cuda.Stream()
    got = ctypes.c_int64(0)
    for nt in (8, 16):
        for rep in range(4):
            torch.cuda.synchronize(); t = time.perf_counter()
            L.ddm_ingest_packed(z.data_ptr(), 85, N, P, chunk, out.data_ptr(), pk.data_ptr(), ready.data_ptr(),
                                marks.data_ptr(), nt, cs.cuda_stream, ctypes.byref(got))
            t1 = time.perf_counter() - t
            torch.cuda.synchronize(); t2 = time.perf_counter() - t
        print(f"ingest loop {nt} threads: host {t1 * 1e3:.2f} ms, copies done {t2 * 1e3:.2f} ms")
    for ch in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
        best = 1e9
        for rep in range(3):
            t = time.perf_counter()
            for a in range(0, N, ch):
                L.ddm_pack_z_host(z.data_ptr() + a * 340, 85, ch, P, out.data_ptr() + a * 32, 16)
            best = min(best, time.perf_counter() - t)
        print(f"pack in chunks of {ch}: {best * 1e3:.2f} ms")
