"""Launches the two activation-epilogue GEMMs of a cfg3 block once (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple3d_former_b200 import _lib as L
T, D = 188160, 768
x = torch.randn(T, D, device="cuda").bfloat16()
w1 = (torch.randn(4 * D, D, device="cuda") * 0.02).bfloat16()
b1 = torch.zeros(4 * D, device="cuda")
pre = torch.empty(T, 4 * D, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    a = L.gemm(x, w1, bias=b1, epilogue=L.EPI_GELU, aux_out=pre)          # fwd fc1 + GELU
    dy = L.gemm(x, w1, b_mn=True, epilogue=L.EPI_DGELU, aux_in=pre) if False else None
w2 = (torch.randn(D, 4 * D, device="cuda") * 0.02).bfloat16()            # stored [D, 4D]: B operand MN-major for dX
dy16 = torch.randn(T, D, device="cuda").bfloat16()
for _ in range(2):
    dpre = L.gemm(dy16, w2, b_mn=True, epilogue=L.EPI_DGELU, aux_in=pre)  # dX of fc2 with dGELU epilogue
res = torch.randn(T, D, device="cuda")
wp = (torch.randn(D, D, device="cuda") * 0.02).bfloat16()
for _ in range(2):
    y = L.gemm(x, wp, bias=b1[:D].contiguous(), residual=res, out_dtype=torch.float32)  # proj + residual (fp32)
torch.cuda.synchronize()
print("done")
