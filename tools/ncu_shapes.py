"""ncu target for the bench's `roofline.ncu` table: ONE launch of every top (kernel, shape) of the cfg3 step through the
C ABI, in a fixed order. The order (shape key -> kernel-name patterns) is written to gpurun_out/ncu_shapes_order.json and
tools/ncu_table.py joins it with the capture:

    ncu --set full --clock-control none -k regex:"gemm_bf16|fa_|attn_|layernorm|colsum" -o /tmp/r02_shapes \\
        python tools/ncu_shapes.py
    python tools/ncu_table.py gpurun_out/r02_shapes.ncu-rep gpurun_out/ncu_shapes_order.json   # -> profiles/ncu_table.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from simple3d_former_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
order = []


def rn(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


def gemm(M, N, K, a_mn=0, b_mn=0, epi=0, f32=0, residual=False, aux_out=False):
    a = rn(K, M) if a_mn else rn(M, K)
    b = rn(K, N) if b_mn else rn(N, K)
    kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), epilogue=epi, out_dtype=torch.float32 if f32 else torch.bfloat16)
    if epi in (L.EPI_GELU, L.EPI_RELU):
        kw["bias"] = rn(N, dtype=torch.float32)
    if epi == L.EPI_GELU and aux_out:
        kw["aux_out"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if epi in (L.EPI_DGELU, L.EPI_DRELU):
        kw["aux_in"] = rn(M, N)
    if residual:
        kw["residual"] = rn(M, N, dtype=torch.float32)
        kw["bias"] = rn(N, dtype=torch.float32)
    if a_mn and b_mn:  # weight gradient: accumulate in place like the trainer's gradient sinks (split-K red.add)
        out = torch.zeros(M, N, device=dev)
        kw.update(out=out, residual=out)
        kw.pop("out_dtype")
    L.gemm(a, b, **kw)
    order.append((f"s3d_gemm_bf16[M={M},N={N},K={K},a_mn={a_mn},b_mn={b_mn},epi={epi},f32={f32}]", ["gemm_bf16_kernel"]))
    torch.cuda.synchronize()


T, D = 188160, 768
gemm(T, 4 * D, D, epi=L.EPI_GELU, aux_out=True)            # fc1 + bias + GELU (bf16 out + bf16 pre-activation)
gemm(T, 4 * D, D, b_mn=1, epi=L.EPI_DGELU)                 # dX of fc2 with the dGELU epilogue
gemm(T, D, 4 * D, f32=1, residual=True)                    # fc2 + bias + fp32 residual
gemm(T, D, 4 * D, b_mn=1)                                  # dX of fc1
gemm(4 * D, D, T, a_mn=1, b_mn=1, f32=1)                   # dW fc1
gemm(D, 4 * D, T, a_mn=1, b_mn=1, f32=1)                   # dW fc2
gemm(T, 3 * D, D)                                          # qkv + bias
gemm(T, D, D, f32=1, residual=True)                        # proj + bias + fp32 residual
gemm(T, D, 3 * D, b_mn=1)                                  # dX of qkv
gemm(3 * D, D, T, a_mn=1, b_mn=1, f32=1)                   # dW qkv
gemm(T, D, D, b_mn=1)                                      # dX of proj

x = rn(T, D, dtype=torch.float32)
gam, bet = rn(D, dtype=torch.float32), rn(D, dtype=torch.float32)
_, _, _, mean, rstd = L.layernorm_fwd(x, gam, bet, 1e-6)
order.append(("s3d_layernorm_fwd", ["layernorm_fwd"]))
L.layernorm_bwd(rn(T, D), x, gam, mean, rstd, dres=rn(T, D, dtype=torch.float32), want_bf16=True)
order.append(("s3d_layernorm_bwd", ["layernorm_bwd"]))
L.colsum(rn(T, 4 * D))
order.append(("s3d_colsum_bf16", ["colsum"]))
del x
torch.cuda.synchronize()


def attn(B, N, H, dh, seqfirst, drop):
    E = H * dh
    seed = torch.tensor([20210915], dtype=torch.int32, device=dev)
    if seqfirst:
        qkv = rn(N * B, 3 * E, scale=0.5)
        qs, os_ = (3 * E, dh, B * 3 * E), (E, dh, B * E)
    else:
        qkv = rn(B * N, 3 * E, scale=0.5)
        qs, os_ = (N * 3 * E, dh, 3 * E), (N * E, dh, E)
    o = torch.empty(B * N, E, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev, dtype=torch.float32)
    b = qkv.data_ptr()
    kw = dict(drop_seed=seed, drop_site=1, drop_p=0.1) if drop else {}
    L.attn_fwd(b, b + 2 * E, b + 4 * E, o, lse, B, H, N, dh, qs, os_, dh ** -0.5, **kw)
    tc_f = drop or N >= 64          # forward: tcgen05 for every sequence of at least one 64-key block
    # backward: single score pass (flash dQ kernel that spills P o mask / dS + two batched GEMMs; head_dim 256: spill-only
    # kernel + three GEMMs) from N = 128; the 15-token sequences keep the warp-per-sequence kernel
    tc = drop or N >= 128
    order.append((f"s3d_attn_fwd[B={B},H={H},N={N},dh={dh},drop={int(drop)}]", ["fa_fwd" if tc_f else "attn_fwd"]))
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    db = dqkv.data_ptr()
    L.attn_bwd(b, b + 2 * E, b + 4 * E, o, rn(B * N, E), lse, delta, db, db + 2 * E, db + 4 * E, B, H, N, dh, qs, os_,
               dh ** -0.5, **kw)
    pats = (["fa_delta", "fa_bwd_dq", "gemm_bf16_kernel", "gemm_bf16_kernel"] + (["gemm_bf16_kernel"] if dh == 256 else [])) if tc else (
        ["attn_bwd_small"] if N <= 16 else ["attn_delta", "attn_bwd_dq", "attn_bwd_dkv"])
    order.append((f"s3d_attn_bwd[B={B},H={H},N={N},dh={dh},drop={int(drop)}]", pats))
    torch.cuda.synchronize()


attn(15, 12544, 4, 192, True, True)     # group_embed, dropout p = 0.1 (what the timed training step runs)
attn(15, 12544, 4, 192, True, False)    # same without dropout (eval / p = 0)
attn(12544, 15, 3, 256, False, False)   # stage 1: 15-token sequences
attn(64, 197, 3, 256, False, False)     # stage 2
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "ncu_shapes_order.json"), "w") as f:
    json.dump(order, f, indent=1)
print("launched", len(order), "labelled calls")
