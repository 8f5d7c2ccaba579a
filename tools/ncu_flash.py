"""ncu target: the group_embed flash-attention kernels (S = 12544, 15 columns, 4 heads of 192), one forward and one
backward launch each without and with attention-probability dropout (p = 0.1).

    ncu --set full --clock-control none --import-source on -k regex:"fa_" -o gpurun_out/r02_flash python tools/ncu_flash.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from simple3d_former_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
S, Nb, H, dh = 12544, int(os.environ.get("NB", "15")), 4, 192
E = H * dh
seed = torch.tensor([20210915], dtype=torch.int32, device=dev)
qkv = (torch.randn(S * Nb, 3 * E, device=dev, generator=g) * 0.5).bfloat16()
dout = (torch.randn(S * Nb, E, device=dev, generator=g)).bfloat16()
o = torch.empty(S * Nb, E, device=dev, dtype=torch.bfloat16)
lse = torch.empty(Nb, H, S, device=dev, dtype=torch.float32)
qs, os_ = (3 * E, dh, Nb * 3 * E), (E, dh, Nb * E)
b = qkv.data_ptr()
dqkv = torch.empty_like(qkv)
delta = torch.empty_like(lse)
db = dqkv.data_ptr()
for drop in ([False, True] if os.environ.get("DROP", "both") == "both" else [os.environ["DROP"] == "1"]):
    kw = dict(drop_seed=seed, drop_site=1, drop_p=0.1) if drop else {}
    L.attn_fwd(b, b + 2 * E, b + 4 * E, o, lse, Nb, H, S, dh, qs, os_, dh ** -0.5, **kw)
    L.attn_bwd(b, b + 2 * E, b + 4 * E, o, dout, lse, delta, db, db + 2 * E, db + 4 * E, Nb, H, S, dh, qs, os_, dh ** -0.5, **kw)
torch.cuda.synchronize()
print("done")
