"""One weight-gradient GEMM (dW fc1 of cfg3: M=3072, N=768, K=188160, both operands MN-major, accumulate) for ncu A/B runs."""
import os, sys
root = os.environ.get("S3D_TREE", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root)
import torch
from simple3d_former_b200 import _lib as L
T, D = 188160, 768
a = torch.randn(T, 4 * D, device="cuda").bfloat16()
b = torch.randn(T, D, device="cuda").bfloat16()
out = torch.zeros(4 * D, D, device="cuda")
for _ in range(3):
    L.gemm(a, b, a_mn=True, b_mn=True, out=out, residual=out)
a2 = torch.randn(T, D, device="cuda").bfloat16()
b2 = torch.randn(T, 4 * D, device="cuda").bfloat16()
out2 = torch.zeros(D, 4 * D, device="cuda")
for _ in range(3):
    L.gemm(a2, b2, a_mn=True, b_mn=True, out=out2, residual=out2)
torch.cuda.synchronize()
print("done", L.LIB_PATH)
