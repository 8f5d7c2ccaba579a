"""Small invocations of the kernels added in round 1, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitizer_targets.py

Shapes are chosen to hit tails (rows / channels not multiples of the vector widths, partial chunks, padded sequences)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from simple3d_former_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def rn(*s):
    return torch.randn(*s, generator=g).to(dev)


if which in ("all", "ln"):
    # bulk-async pipelined LayerNorm (T >= 148 * warps * 4 rows) forward / backward, bf16 and fp32 dy, with / without dres
    for (T, D) in [(9500, 384), (7200, 768), (9473, 192)]:
        x, gam, bet = rn(T, D), rn(D), rn(D)
        y16, y32, _, mean, rstd = L.layernorm_fwd(x, gam, bet, 1e-6, want_f32=True)
        L.layernorm_fwd(x, gam, bet, 1e-5, addend=rn(T, D), want_sum=True, want_f32=True)
        L.layernorm_bwd(rn(T, D).bfloat16(), x, gam, mean, rstd, dres=rn(T, D), want_bf16=True)
        L.layernorm_bwd(rn(T, D), x, gam, mean, rstd)
    L.colsum(rn(5001, 776).bfloat16())
    L.colsum(rn(300, 70).bfloat16())
if which in ("all", "attn"):
    # tiny-sequence attention forward / backward (persistent, double-buffered), padded rows
    for (B, H, N, dh) in [(700, 3, 15, 256), (50, 3, 5, 256), (33, 4, 9, 192), (40, 6, 16, 64)]:
        E = H * dh
        qkv = (rn(B, N, 3, H, dh) * 0.5).bfloat16()
        out = torch.empty(B, N, E, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, N, device=dev)
        qs, os_ = (N * 3 * E, dh, 3 * E), (N * E, dh, E)
        base = qkv.data_ptr()
        L.attn_fwd(base, base + 2 * E, base + 4 * E, out, lse, B, H, N, dh, qs, os_, dh ** -0.5)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        db = dqkv.data_ptr()
        L.attn_bwd(base, base + 2 * E, base + 4 * E, out, rn(B, N, E).bfloat16(), lse, delta, db, db + 2 * E, db + 4 * E, B, H,
                   N, dh, qs, os_, dh ** -0.5)
if which in ("all", "points"):
    from simple3d_former_b200.models import TransitionUp
    from simple3d_former_b200.pointnet_util import PointNetSetAbstraction
    for (B, N, S, K, Cf, C) in [(2, 300, 77, 8, 96, 192), (3, 256, 256, 16, 48, 96), (2, 130, 65, 16, 16, 40)]:
        sa = PointNetSetAbstraction(S, 0, K, Cf + 3, [C, C], False, knn=True).to(dev).train()
        sa.fps_start = torch.zeros(B, dtype=torch.long, device=dev)
        pts = rn(B, N, Cf).requires_grad_(True)
        _, y = sa((torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev), pts)
        y.sum().backward()
    tu = TransitionUp(64, 32, 40).to(dev).train()
    xyz2 = (torch.rand(3, 250, 3, generator=g) * 2 - 1).to(dev)
    p1, p2 = rn(3, 61, 64).requires_grad_(True), rn(3, 250, 32).requires_grad_(True)
    tu(xyz2[:, :61].contiguous(), p1, xyz2, p2).sum().backward()
if which in ("all", "binvox"):
    import binvox_np as BO
    from simple3d_former_b200 import binvox_rw as P
    rng = np.random.default_rng(1)
    files = [BO.write(rng.random((V, V, V)) < 0.2) for V in (30, 30)] + [BO.write(np.zeros((30, 30, 30), bool))]
    P.load_voxel_batch(files)
    P.load_voxel_batch([BO.write(rng.random((64, 64, 64)) < 0.05)], dtype=torch.float32, fix_coords=False)
if which in ("all", "dropout"):
    seed = torch.tensor([77], dtype=torch.int32, device=dev)
    L.dropout_add(rn(333, 96), rn(333, 96), seed, 2, 0.1)
    L.dropout_bf16(rn(333, 96).bfloat16(), seed, 3, 0.1)
    B, H, N, dh = 2, 2, 300, 64
    E = H * dh
    qkv = (rn(B, N, 3, H, dh) * 0.5).bfloat16()
    out = torch.empty(B, N, E, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev)
    qs, os_ = (N * 3 * E, dh, 3 * E), (N * E, dh, E)
    base = qkv.data_ptr()
    L.attn_fwd(base, base + 2 * E, base + 4 * E, out, lse, B, H, N, dh, qs, os_, dh ** -0.5, drop_seed=seed, drop_site=1, drop_p=0.1)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    db = dqkv.data_ptr()
    L.attn_bwd(base, base + 2 * E, base + 4 * E, out, rn(B, N, E).bfloat16(), lse, delta, db, db + 2 * E, db + 4 * E, B, H, N,
               dh, qs, os_, dh ** -0.5, drop_seed=seed, drop_site=1, drop_p=0.1)
if which in ("all", "round2"):
    # round 2: single-score-pass attention backward (spill stores through rank-4 tensor maps + batched panel GEMMs) at
    # sizes with partial query tiles / key blocks, both layouts; fp32 strided GEMM; staged voxel gather; LN backward with
    # the fused column sums
    os.environ["S3D_FA_SPILL_MIN_N"] = "1"
    seed = torch.tensor([5], dtype=torch.int32, device=dev)
    for (B, H, N, dh, seqfirst, p_drop) in [(2, 2, 300, 64, False, 0.1), (3, 4, 130, 192, True, 0.0), (1, 3, 257, 64, False, 0.0),
                                            (2, 4, 200, 96, True, 0.1), (2, 4, 129, 48, True, 0.0)]:
        E = H * dh
        if seqfirst:
            qkv = (rn(N * B, 3 * E) * 0.5).bfloat16()
            qs, os_ = (3 * E, dh, B * 3 * E), (E, dh, B * E)
        else:
            qkv = (rn(B * N, 3 * E) * 0.5).bfloat16()
            qs, os_ = (N * 3 * E, dh, 3 * E), (N * E, dh, E)
        out = torch.empty(B * N, E, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, N, device=dev)
        base = qkv.data_ptr()
        kw = dict(drop_seed=seed, drop_site=1, drop_p=p_drop) if p_drop > 0 else {}
        L.attn_fwd(base, base + 2 * E, base + 4 * E, out, lse, B, H, N, dh, qs, os_, dh ** -0.5, **kw)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        db = dqkv.data_ptr()
        assert L.lib().s3d_attn_bwd_workspace_bytes(B, H, N, dh) > 0
        L.attn_bwd(base, base + 2 * E, base + 4 * E, out, rn(B * N, E).bfloat16(), lse, delta, db, db + 2 * E, db + 4 * E,
                   B, H, N, dh, qs, os_, dh ** -0.5, workspace=True, **kw)
    for (M, N, K) in [(1000, 48, 6), (130, 70, 33), (48, 48, 9000), (1, 48, 5000)]:
        L.sgemm(rn(M, K), rn(K, N), bias=rn(N))
        L.sgemm(rn(K, M).t(), rn(N, K).t(), relu=True)
    L.sgemm(rn(300, 20), rn(20, 50), gate=rn(300, 50))
    for (B, V, c, p, zsum, dt) in [(2, 30, 6, 5, True, torch.float32), (1, 128, 9, 14, False, torch.uint8),
                                   (2, 30, 6, 5, False, torch.int32), (1, 33, 4, 8, False, torch.uint8)]:
        vox = (torch.rand(B, 1, V, V, V, generator=g) < 0.1).to(dt).to(dev)
        L.voxel_patch_gather(vox, c, p, (c ** 3 + 63) // 64 * 64, zsum)
    for (T, D) in [(9500, 384), (333, 192)]:
        x, gam, bet = rn(T, D), rn(D), rn(D)
        _, _, _, mean, rstd = L.layernorm_fwd(x, gam, bet, 1e-6)
        L.layernorm_bwd(rn(T, D).bfloat16(), x, gam, mean, rstd, dres=rn(T, D), want_bf16=True, want_dxsum=True)
torch.cuda.synchronize()
print("sanitizer targets done:", which)
