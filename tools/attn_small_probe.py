"""Stage-1 attention of cfg3 (12544 x 3 heads, 15 tokens, head_dim 256): forward / backward time and HBM rate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple3d_former_b200 import _lib as L
B, N, H, dh = 12544, 15, 3, 256
E = H * dh
qkv = (torch.randn(B * N, 3 * E, device="cuda") * 0.5).bfloat16()
o = torch.empty(B * N, E, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device="cuda"); delta = torch.empty_like(lse)
do = torch.randn(B * N, E, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
qs, os_ = (N * 3 * E, dh, 3 * E), (N * E, dh, E)
b, db = qkv.data_ptr(), dqkv.data_ptr()
def t(fn, reps=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
f = t(lambda: L.attn_fwd(b, b + 2 * E, b + 4 * E, o, lse, B, H, N, dh, qs, os_, dh ** -0.5))
w = t(lambda: L.attn_bwd(b, b + 2 * E, b + 4 * E, o, do, lse, delta, db, db + 2 * E, db + 4 * E, B, H, N, dh, qs, os_, dh ** -0.5))
T = B * N
print(f"fwd {f:.3f} ms  {T * E * 2 * 4 / f / 1e9:.2f} TB/s   bwd {w:.3f} ms  {T * E * 2 * 8 / w / 1e9:.2f} TB/s")
