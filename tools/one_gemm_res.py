"""fc2 + bias + fp32 residual GEMM of a cfg3 block (M=188160, N=768, K=3072) once, for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple3d_former_b200 import _lib as L
T, D = 188160, 768
a = torch.randn(T, 4 * D, device="cuda").bfloat16()
w = (torch.randn(D, 4 * D, device="cuda") * 0.02).bfloat16()
b = torch.zeros(D, device="cuda")
res = torch.randn(T, D, device="cuda")
for _ in range(3):
    y = L.gemm(a, w, bias=b, residual=res, out_dtype=torch.float32)
    y2 = L.gemm(a, w)  # same shape, plain bf16 epilogue
torch.cuda.synchronize()
print("done")
