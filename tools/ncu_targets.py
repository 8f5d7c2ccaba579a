"""One launch of every hot kernel of the cfg3 step at its real shape, for `ncu --set full` captures:

    ncu --set full --clock-control none --import-source on -o gpurun_out/r01_cfg3_kernels python tools/ncu_targets.py

(stage-1 tokens T = 64 * 196 * 15 = 188160, D = 768, 3 heads of 256; group_embed attention S = 12544, 4 heads of 192).
Numbers printed under ncu are not bench values; bench.py measures the same launches with CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from simple3d_former_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
T, D = 188160, 768
g = torch.Generator(device=dev).manual_seed(0)


def rn(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


x16 = rn(T, D)
x32 = rn(T, D, dtype=torch.float32)
w_fc1 = rn(4 * D, D, scale=0.02)
w_fc2 = rn(D, 4 * D, scale=0.02)
w_qkv = rn(3 * D, D, scale=0.02)
w_proj = rn(D, D, scale=0.02)
b4 = torch.zeros(4 * D, device=dev)
b3 = torch.zeros(3 * D, device=dev)
b1 = torch.zeros(D, device=dev)
pre = torch.empty(T, 4 * D, device=dev, dtype=torch.bfloat16)

# ---- GEMMs of one encoder block
a16 = L.gemm(x16, w_fc1, bias=b4, epilogue=L.EPI_GELU, aux_out=pre)                      # fc1 + bias + GELU
dpre = L.gemm(x16, w_fc2, b_mn=True, epilogue=L.EPI_DGELU, aux_in=pre)                   # dX of fc2 with dGELU
y = L.gemm(a16, w_fc2, bias=b1, residual=x32, out_dtype=torch.float32)                   # fc2 + bias + residual (fp32)
dw = torch.zeros(4 * D, D, device=dev)
L.gemm(dpre, x16, a_mn=True, b_mn=True, out=dw, residual=dw)                             # dW fc1 (split-K, accumulate)
y2 = L.gemm(x16, w_proj, bias=b1, residual=x32, out_dtype=torch.float32)                 # proj + bias + residual (fp32)
qkv = L.gemm(x16, w_qkv, bias=b3)                                                        # qkv projection
del a16, y, y2, pre

# ---- encoder attention core on 15-token sequences (3 heads of 256)
B, N, H, dh = 12544, 15, 3, 256
E = H * dh
out = torch.empty(B, N, E, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device=dev, dtype=torch.float32)
qs, os_ = (N * 3 * E, dh, 3 * E), (N * E, dh, E)
base = qkv.data_ptr()
L.attn_fwd(base, base + 2 * E, base + 4 * E, out, lse, B, H, N, dh, qs, os_, dh ** -0.5)
dqkv = torch.empty_like(qkv)
delta = torch.empty_like(lse)
dbase = dqkv.data_ptr()
L.attn_bwd(base, base + 2 * E, base + 4 * E, out, x16.view(B, N, E), lse, delta, dbase, dbase + 2 * E, dbase + 4 * E, B, H,
           N, dh, qs, os_, dh ** -0.5)

# ---- LayerNorm forward / backward, bias-gradient column sums
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
h16, _, _, mean, rstd = L.layernorm_fwd(x32, gamma, beta, 1e-6)
L.layernorm_bwd(x16, x32, gamma, mean, rstd, dres=x32, want_bf16=True)
L.colsum(dpre)
del dpre, dqkv

# ---- group_embed attention (sequence-first, S = 12544, 15 "batch" columns, 4 heads of 192)
S, Nb, Hg, dg = 12544, 15, 4, 192
Eg = Hg * dg
qkv_g = rn(S * Nb, 3 * Eg)
o_g = torch.empty(S * Nb, Eg, device=dev, dtype=torch.bfloat16)
lse_g = torch.empty(Nb, Hg, S, device=dev, dtype=torch.float32)
qs_g, os_g = (3 * Eg, dg, Nb * 3 * Eg), (Eg, dg, Nb * Eg)
bg = qkv_g.data_ptr()
L.attn_fwd(bg, bg + 2 * Eg, bg + 4 * Eg, o_g, lse_g, Nb, Hg, S, dg, qs_g, os_g, dg ** -0.5)
dqkv_g = torch.empty_like(qkv_g)
delta_g = torch.empty_like(lse_g)
dbg = dqkv_g.data_ptr()
do_g = rn(S * Nb, Eg)
L.attn_bwd(bg, bg + 2 * Eg, bg + 4 * Eg, o_g, do_g, lse_g, delta_g, dbg, dbg + 2 * Eg, dbg + 4 * Eg, Nb, Hg, S, dg, qs_g,
           os_g, dg ** -0.5)
torch.cuda.synchronize()
print("done")
