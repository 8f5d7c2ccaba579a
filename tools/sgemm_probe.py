"""Per-call times of s3d_sgemm_f32 on the point-stem / head shapes of cfg4 (R = 128 x 1024 points, q = 48)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple3d_former_b200 import _lib as L

def t(fn, reps=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3

R, q = 131072, 48
x6 = torch.randn(R, 6, device="cuda"); x3 = torch.randn(R, 3, device="cuda")
w6 = torch.randn(q, 6, device="cuda"); wq = torch.randn(q, q, device="cuda"); b = torch.randn(q, device="cuda")
h = torch.randn(R, q, device="cuda").relu(); df = torch.randn(R, q, device="cuda")
ones = torch.ones(1, R, device="cuda")
out = torch.empty(R, q, device="cuda")
print("fwd  x6 W^T relu   [R,6]x[6,48]  :", round(t(lambda: L.sgemm(x6, w6.t(), bias=b, relu=True, out=out)), 1), "us")
print("fwd  h  W^T        [R,48]x[48,48]:", round(t(lambda: L.sgemm(h, wq.t(), bias=b, out=out)), 1), "us")
print("bwd  dh = df W gate[R,48]x[48,48]:", round(t(lambda: L.sgemm(df, wq, gate=h, out=out)), 1), "us")
print("bwd  dW = df^T h   [48,R]x[R,48] :", round(t(lambda: L.sgemm(df.t(), h)), 1), "us")
print("bwd  dW = dh^T x6  [48,R]x[R,6]  :", round(t(lambda: L.sgemm(df.t(), x6)), 1), "us")
print("bwd  db = 1^T df   [1,R]x[R,48]  :", round(t(lambda: L.sgemm(ones, df)), 1), "us")
print("torch h @ W^T                     :", round(t(lambda: torch.addmm(b, h, wq.t())), 1), "us")
print("torch df^T @ h                    :", round(t(lambda: df.t() @ h), 1), "us")
