"""Cycles per tcgen05.mma (M=128, K=16, bf16) for the operand layouts the MNLE kernel uses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sbi_for_diffusion_models_b200 import _native
L = _native.lib()
dev = "cuda"
A = torch.randn(128, 128, device=dev)
for N in (128, 80, 16):
    B = torch.randn(N, 128, device=dev)
    D = torch.zeros((128, N), device=dev)
    for base, name in ((100, "A smem"), (200, "A tmem")):
        res = []
        for reps in (1, 5, 20, 40):
            L.mnle_tc_selftest(A.data_ptr(), B.data_ptr(), N, base + reps, 0, 0, 0, D.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            res.append((reps * 24, D[0, 0].item()))
        slope = (res[-1][1] - res[1][1]) / (res[-1][0] - res[1][0])
        print(f"N={N} {name}: {res} -> {slope:.1f} cycles per MMA (floor {128 * N / 256:.0f})")
