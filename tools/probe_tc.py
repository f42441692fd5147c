import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sbi_for_diffusion_models_b200 import _native
L = _native.lib()
torch.manual_seed(0)
dev = "cuda"
def run(N, passes, lbo_a=0, lbo_b=0, sbo=0):
    A = torch.randn(128, 128, device=dev); B = torch.randn(N, 128, device=dev)
    D = torch.full((128, N), float("nan"), device=dev)
    rc = L.mnle_tc_selftest(A.data_ptr(), B.data_ptr(), N, passes, lbo_a, lbo_b, sbo, D.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref64 = A.double() @ B.double().T
    refbf = A.bfloat16().double() @ B.bfloat16().double().T
    return rc, (D.double() - ref64).abs().max().item(), (D.double() - refbf).abs().max().item()
for N in (128, 80, 16):
    for passes in (1, 3, 11, 13):
        print("N", N, "passes", passes, "rc, err_vs_fp64, err_vs_bf16prod:", run(N, passes))
print("swapped lbo/sbo:", run(128, 1, lbo_a=128, lbo_b=128, sbo=2048))
