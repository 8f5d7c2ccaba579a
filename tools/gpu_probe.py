"""Bring-up probe: runs each kernel family against plain torch on the GPU and prints max errors.
Every group runs in its own subprocess (a trapped kernel kills only its own CUDA context).

    python tools/gpu_probe.py            # run all groups
    python tools/gpu_probe.py gemm_k     # run one group in-process
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["attn_tc", "attn_spill", "sgemm", "gemm_epi_perf", "gemm_small", "gemm_perf", "gemm_k", "gemm_mn", "gemm_epi", "gemm_batched", "ln", "attn_fwd", "attn_bwd", "elementwise", "points"]


def rel_err(a, b):
    import torch
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def report(name, err, tol):
    print(f"  [{'PASS' if err <= tol else 'FAIL'}] {name}: rel_err={err:.3e} (tol {tol:g})", flush=True)


def g_gemm_perf():
    """TFLOP/s of the shapes that dominate cfg3 / cfg2, per (BN, cluster, splits) variant (CUDA events, L2-cold-ish)."""
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(0)
    shapes = [("fwd qkv  cfg3", 188160, 2304, 768, False, False), ("fwd fc1  cfg3", 188160, 3072, 768, False, False),
              ("fwd fc2  cfg3", 188160, 768, 3072, False, False), ("dX  fc1  cfg3", 188160, 768, 3072, False, True),
              ("dW  fc1  cfg3", 3072, 768, 188160, True, True), ("dW  qkv  cfg3", 2304, 768, 188160, True, True),
              ("fwd fc1  cfg2", 1664, 1536, 384, False, False), ("dW  fc1  cfg2", 1536, 384, 1664, True, True)]
    for (name, M, N, K, amn, bmn) in shapes:
        a = torch.randn((K, M) if amn else (M, K), device="cuda").bfloat16()
        b = torch.randn((K, N) if bmn else (N, K), device="cuda").bfloat16()
        dw = amn and bmn
        out = torch.empty(M, N, device="cuda", dtype=torch.float32 if dw else torch.bfloat16)
        variants = [(256, 1, 0), (256, 2, 0), (256, 4, 0), (128, 2, 0), (128, 1, 0)]
        if dw:
            variants += [(256, 2, 1), (256, 1, 1), (128, 2, 0)]
        for (bn, cl, sp) in variants:
            if N < bn:
                continue
            for _ in range(2):
                L.gemm(a, b, a_mn=amn, b_mn=bmn, out=out, force_bn=bn, force_cluster=cl, force_splits=sp)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if M * N * K > 1e11 else 20
            s.record()
            for _ in range(reps):
                L.gemm(a, b, a_mn=amn, b_mn=bmn, out=out, force_bn=bn, force_cluster=cl, force_splits=sp)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / reps
            print(f"  [PERF] {name} M{M} N{N} K{K} BN{bn} cl{cl} split{sp}: {ms * 1e3:9.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)


def g_gemm_epi_perf():
    """TFLOP/s of the epilogue-heavy cfg3 shapes (GELU + pre-activation save, dGELU, fp32 residual), default heuristics."""
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(0)
    T, D = 188160, 768

    def t(fn, reps=5):
        for _ in range(2):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    x = torch.randn(T, D, device="cuda").bfloat16()
    w1 = torch.randn(4 * D, D, device="cuda").bfloat16()
    b1 = torch.randn(4 * D, device="cuda")
    pre = torch.empty(T, 4 * D, device="cuda", dtype=torch.bfloat16)
    act = torch.empty(T, 4 * D, device="cuda", dtype=torch.bfloat16)
    ms = t(lambda: L.gemm(x, w1, bias=b1, epilogue=L.EPI_GELU, aux_out=pre, out=act))
    print(f"  [PERF] fc1 + bias + GELU (+pre-activation): {ms * 1e3:8.1f} us {2.0 * T * 4 * D * D / ms / 1e9:7.1f} TFLOP/s", flush=True)
    ms = t(lambda: L.gemm(x, w1, bias=b1, out=act))
    print(f"  [PERF] fc1 + bias (plain epilogue):         {ms * 1e3:8.1f} us {2.0 * T * 4 * D * D / ms / 1e9:7.1f} TFLOP/s", flush=True)
    dy = torch.randn(T, D, device="cuda").bfloat16()
    w2 = torch.randn(D, 4 * D, device="cuda").bfloat16()
    ms = t(lambda: L.gemm(dy, w2, b_mn=True, epilogue=L.EPI_DGELU, aux_in=pre, out=act))
    print(f"  [PERF] dX of fc2 with dGELU:                {ms * 1e3:8.1f} us {2.0 * T * 4 * D * D / ms / 1e9:7.1f} TFLOP/s", flush=True)
    res = torch.randn(T, D, device="cuda")
    o32 = torch.empty(T, D, device="cuda")
    b2 = torch.randn(D, device="cuda")
    ms = t(lambda: L.gemm(act, w2, bias=b2, residual=res, out=o32))
    print(f"  [PERF] fc2 + bias + fp32 residual:          {ms * 1e3:8.1f} us {2.0 * T * 4 * D * D / ms / 1e9:7.1f} TFLOP/s", flush=True)
    wp = torch.randn(D, D, device="cuda").bfloat16()
    ms = t(lambda: L.gemm(x, wp, bias=b2, residual=res, out=o32))
    print(f"  [PERF] proj + bias + fp32 residual:         {ms * 1e3:8.1f} us {2.0 * T * D * D / ms / 1e9:7.1f} TFLOP/s", flush=True)


def g_gemm_small():
    """GPU time per launch of small GEMMs (CUDA graph of 50 launches: no CPU launch overhead), vs cuBLAS in a graph."""
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(0)

    def graph_time(fn, n=50):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                fn()
        g.replay()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(4):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / (4 * n) * 1e3

    for (name, M, N, K, amn, bmn, f32) in [("1 tile 1 kblock", 128, 128, 64, False, False, False),
                                           ("1 tile 6 kblocks", 128, 128, 384, False, False, False),
                                           ("fwd qkv cfg2", 1664, 1152, 384, False, False, False),
                                           ("fwd fc1 cfg2", 1664, 1536, 384, False, False, False),
                                           ("fwd fc2 cfg2 f32", 1664, 384, 1536, False, False, True),
                                           ("dX fc2 cfg2", 1664, 1536, 384, False, True, False),
                                           ("dW fc1 cfg2 f32", 1536, 384, 1664, True, True, True)]:
        a = torch.randn((K, M) if amn else (M, K), device="cuda").bfloat16()
        b = torch.randn((K, N) if bmn else (N, K), device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
        res = []
        for (bn, cl, sp) in [(0, 0, 0), (128, 1, 1), (64, 1, 1), (256, 1, 1)]:
            if bn > N and bn != 0:
                continue
            us = graph_time(lambda: L.gemm(a, b, a_mn=amn, b_mn=bmn, out=out, force_bn=bn, force_cluster=cl, force_splits=sp))
            res.append(f"BN{bn}/cl{cl}/sp{sp}: {us:6.1f} us")
        am = a.t() if amn else a
        bm = b if bmn else b.t()
        ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        us = graph_time(lambda: torch.matmul(am, bm, out=ref))
        print(f"  [PERF] {name:18s} M{M} N{N} K{K}: " + " | ".join(res) + f" | cuBLAS {us:6.1f} us", flush=True)


def g_gemm_k():
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(0)
    for (M, N, K) in [(128, 128, 64), (128, 256, 256), (256, 128, 512), (1664, 1152, 384), (1664, 384, 1536),
                      (200, 192, 192), (333, 40, 72), (12608, 768, 3072)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        ref = a.float() @ b.float().t()
        for bn in (64, 128, 256):
            out = L.gemm(a, b, out_dtype=torch.float32, force_bn=bn)
            torch.cuda.synchronize()
            report(f"gemm K-major M{M} N{N} K{K} BN{bn}", rel_err(out, ref), 2e-3)
        out = L.gemm(a, b)  # auto BN, bf16 out
        torch.cuda.synchronize()
        report(f"gemm K-major M{M} N{N} K{K} auto bf16", rel_err(out, ref), 1e-2)
        for (bn, cl) in ((128, 2), (256, 2), (256, 4), (256, 1)):
            out = L.gemm(a, b, out_dtype=torch.float32, force_bn=bn, force_cluster=cl)
            torch.cuda.synchronize()
            report(f"gemm K-major M{M} N{N} K{K} BN{bn} cluster{cl}", rel_err(out, ref), 2e-3)
        for sp in (2, 5):
            out = L.gemm(a, b, out_dtype=torch.float32, force_splits=sp)
            acc = ref.clone()
            L.gemm(a, b, out=acc, residual=acc, force_splits=sp)
            torch.cuda.synchronize()
            report(f"gemm K-major M{M} N{N} K{K} splitK{sp}", rel_err(out, ref), 2e-3)
            report(f"gemm K-major M{M} N{N} K{K} splitK{sp} accumulate", rel_err(acc, 2 * ref), 2e-3)


def g_gemm_mn():
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(1)
    for (M, N, K) in [(128, 128, 64), (256, 256, 256), (1536, 384, 1664), (384, 384, 1000), (192, 576, 333 * 8)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        ref = a.float() @ b.float().t()
        at = a.t().contiguous()  # [K, M]
        bt = b.t().contiguous()  # [K, N]
        for bn in (64, 128, 256):
            o1 = L.gemm(at, b, a_mn=True, out_dtype=torch.float32, force_bn=bn)
            o2 = L.gemm(a, bt, b_mn=True, out_dtype=torch.float32, force_bn=bn)
            o3 = L.gemm(at, bt, a_mn=True, b_mn=True, out_dtype=torch.float32, force_bn=bn)
            torch.cuda.synchronize()
            report(f"gemm A-MN     M{M} N{N} K{K} BN{bn}", rel_err(o1, ref), 2e-3)
            report(f"gemm B-MN     M{M} N{N} K{K} BN{bn}", rel_err(o2, ref), 2e-3)
            report(f"gemm A-MN B-MN M{M} N{N} K{K} BN{bn}", rel_err(o3, ref), 2e-3)
        for (bn, cl, sp) in ((256, 2, 0), (256, 4, 3), (128, 2, 7), (256, 1, 2)):
            o2 = L.gemm(a, bt, b_mn=True, out_dtype=torch.float32, force_bn=bn, force_cluster=cl, force_splits=sp)
            o3 = L.gemm(at, bt, a_mn=True, b_mn=True, out_dtype=torch.float32, force_bn=bn, force_cluster=cl, force_splits=sp)
            torch.cuda.synchronize()
            report(f"gemm B-MN      M{M} N{N} K{K} BN{bn} cluster{cl} split{sp}", rel_err(o2, ref), 2e-3)
            report(f"gemm A-MN B-MN M{M} N{N} K{K} BN{bn} cluster{cl} split{sp}", rel_err(o3, ref), 2e-3)


def g_gemm_epi():
    import torch
    import torch.nn.functional as F
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(2)
    M, N, K = 1664, 1536, 384
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    acc = a.float() @ b.float().t()
    out = L.gemm(a, b, bias=bias, out_dtype=torch.float32)
    report("bias", rel_err(out, acc + bias), 2e-3)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = L.gemm(a, b, bias=bias, epilogue=L.EPI_GELU, aux_out=pre)
    report("bias+gelu", rel_err(out, F.gelu(acc + bias)), 1e-2)
    report("bias+gelu preact", rel_err(pre, acc + bias), 1e-2)
    out = L.gemm(a, b, bias=bias, residual=res, out_dtype=torch.float32)
    report("bias+residual", rel_err(out, acc + bias + res), 2e-3)
    out = L.gemm(a, b, alpha=0.2, out_dtype=torch.float32)
    report("alpha", rel_err(out, 0.2 * acc), 2e-3)
    x = torch.randn(M, N, device="cuda").bfloat16()
    xr = x.float().requires_grad_(True)
    F.gelu(xr).sum().backward()
    out = L.gemm(a, b, epilogue=L.EPI_DGELU, aux_in=x, out_dtype=torch.float32)
    report("dgelu", rel_err(out, acc * xr.grad), 5e-3)
    acc_buf = res.clone()
    L.gemm(a, b, residual=acc_buf, out=acc_buf)
    report("accumulate in place", rel_err(acc_buf, acc + res), 2e-3)
    # N tail with vector epilogue disabled (N=40 head-like), M tail
    M, N, K = 77, 40, 384
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    out = L.gemm(a, b, bias=bias, out_dtype=torch.float32)
    report("tails M77 N40", rel_err(out, a.float() @ b.float().t() + bias), 2e-3)


def g_gemm_batched():
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(3)
    Bn, M, N, K = 6, 300, 200, 192
    a = torch.randn(Bn, M, K, device="cuda").bfloat16()
    b = torch.randn(Bn, N, K, device="cuda").bfloat16()
    out = L.gemm(a, b, out_dtype=torch.float32)
    report("batched K-major", rel_err(out, torch.bmm(a.float(), b.float().transpose(1, 2))), 2e-3)
    bt = b.transpose(1, 2).contiguous()
    out = L.gemm(a, bt, b_mn=True, out_dtype=torch.float32)
    report("batched B-MN", rel_err(out, torch.bmm(a.float(), b.float().transpose(1, 2))), 2e-3)


def g_ln():
    import torch
    import torch.nn.functional as F
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(4)
    # the last four shapes are large enough for the bulk-async pipelined variants (T >= 148 * warps * 4 rows); D = 1024 with
    # an fp32 dy does not fit their shared-memory ring and must fall back to the register-resident kernel
    for (T, D) in [(1664, 384), (1000, 768), (333, 192), (50, 1024), (20000, 768), (12001, 192), (9500, 384), (9600, 1024)]:
        x = torch.randn(T, D, device="cuda") * 2 + 0.5
        g = torch.randn(D, device="cuda")
        b = torch.randn(D, device="cuda")
        y16, y32, _, mean, rstd = L.layernorm_fwd(x, g, b, 1e-6, want_f32=True)
        ref = F.layer_norm(x, (D,), g, b, 1e-6)
        report(f"ln fwd f32 T{T} D{D}", rel_err(y32, ref), 1e-5)
        report(f"ln fwd bf16 T{T} D{D}", rel_err(y16, ref), 1e-2)
        xr = x.clone().requires_grad_(True)
        gr = g.clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        dy = torch.randn(T, D, device="cuda")
        dres = torch.randn(T, D, device="cuda")
        F.layer_norm(xr, (D,), gr, br, 1e-6).backward(dy)
        dx, dx16, dg, db, dxs = L.layernorm_bwd(dy, x, g, mean, rstd, dres=dres, want_bf16=True, want_dxsum=True)
        report(f"ln bwd dx column sums T{T}", rel_err(dxs, dx.sum(0)), 1e-4)
        report(f"ln bwd dx T{T} D{D}", rel_err(dx, xr.grad + dres), 1e-4)
        report(f"ln bwd dx16 T{T} D{D}", rel_err(dx16, xr.grad + dres), 1e-2)
        report(f"ln bwd dgamma T{T} D{D}", rel_err(dg, gr.grad), 1e-4)
        report(f"ln bwd dbeta T{T} D{D}", rel_err(db, br.grad), 1e-4)
        dx, _, dg, db = L.layernorm_bwd(dy.bfloat16(), x, g, mean, rstd)
        report(f"ln bwd (bf16 dy) dx T{T} D{D}", rel_err(dx, xr.grad), 1e-2)
    for T in (100, 11000):
        x = torch.randn(T, 768, device="cuda")
        a = torch.randn(T, 768, device="cuda")
        g = torch.ones(768, device="cuda")
        b = torch.zeros(768, device="cuda")
        _, y32, s, _, _ = L.layernorm_fwd(x, g, b, 1e-5, addend=a, want_sum=True, want_f32=True)
        report(f"ln fwd fused add T{T}", rel_err(y32, F.layer_norm(x + a, (768,), g, b, 1e-5)), 1e-5)
        report(f"ln fwd fused add sum T{T}", rel_err(s, x + a), 1e-6)
    for (T, C) in [(300, 768), (5000, 3072), (777, 100), (20000, 2304)]:  # 16-byte and 4-byte column-sum kernels
        m = torch.randn(T, C, device="cuda").bfloat16()
        report(f"colsum T{T} C{C}", rel_err(L.colsum(m), m.float().sum(0)), 1e-4)
        acc = torch.ones(C, device="cuda")
        L.colsum(m, out=acc, accumulate=True)
        report(f"colsum accumulate T{T} C{C}", rel_err(acc, m.float().sum(0) + 1), 1e-4)


def _attn_ref(q, k, v, scale):
    import torch
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    p = s.softmax(-1)
    return p @ v.float(), torch.logsumexp(s, -1)


def _attn_cases():
    # (B, N, H, dh, layout)
    return [(3, 15, 3, 256, "timm"), (130, 15, 3, 256, "timm"), (4, 26, 6, 64, "timm"), (2, 197, 3, 256, "timm"),
            (2, 257, 3, 64, "timm"), (1, 513, 3, 64, "timm"), (2, 16, 3, 64, "timm"), (3, 15, 3, 64, "timm"),
            (2, 700, 4, 192, "seqfirst"), (15, 392, 4, 192, "seqfirst")]


def _make_qkv(B, N, H, dh, layout):
    import torch
    E = H * dh
    if layout == "timm":
        qkv = (torch.randn(B, N, 3, H, dh, device="cuda")).bfloat16()
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]  # [B,N,H,dh]
        qs = (N * 3 * E, dh, 3 * E)
        os_ = (N * E, dh, E)
        out = torch.empty(B, N, H, dh, device="cuda", dtype=torch.bfloat16)
        perm = lambda t: t.permute(0, 2, 1, 3)  # -> [B,H,N,dh]
    else:  # sequence-first [S=N, Nb=B, 3E]
        qkv = torch.randn(N, B, 3, H, dh, device="cuda").bfloat16()
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]  # [N,B,H,dh]
        qs = (3 * E, dh, B * 3 * E)
        os_ = (E, dh, B * E)
        out = torch.empty(N, B, H, dh, device="cuda", dtype=torch.bfloat16)
        perm = lambda t: t.permute(1, 2, 0, 3)
    return qkv, q, k, v, qs, os_, out, perm


def g_attn_tc():
    """tcgen05 flash attention (long sequences): correctness vs torch and TFLOP/s (S3D_ATTN_TC=0 -> mma.sync kernels)."""
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(11)
    print("  S3D_ATTN_TC =", os.environ.get("S3D_ATTN_TC", "1"), flush=True)
    for (B, N, H, dh, layout) in [(2, 1024, 4, 192, "seqfirst"), (3, 700, 4, 192, "seqfirst"), (2, 12544, 4, 192, "seqfirst"),
                                  (2, 513, 3, 64, "timm"), (3, 640, 2, 192, "timm"), (1, 2048, 3, 64, "timm"),
                                  (2, 700, 4, 96, "seqfirst"), (2, 600, 4, 48, "seqfirst"), (3, 100, 4, 96, "seqfirst"),
                                  (2, 333, 4, 48, "timm"), (4, 197, 3, 256, "timm"), (3, 257, 3, 64, "timm"),
                                  (2, 64, 3, 256, "timm"), (2, 1000, 2, 256, "seqfirst"), (5, 65, 6, 64, "timm")]:
        qkv, q, k, v, qs, os_, out, perm = _make_qkv(B, N, H, dh, layout)
        lse = torch.empty(B, H, N, device="cuda")
        scale = dh ** -0.5
        L.attn_fwd(q, k, v, out, lse, B, H, N, dh, qs, os_, scale)
        torch.cuda.synchronize()
        ro, rl = _attn_ref(perm(q), perm(k), perm(v), scale)
        report(f"attn fwd out B{B} N{N} H{H} dh{dh} {layout}", rel_err(perm(out), ro), 1.5e-2)
        report(f"attn fwd lse B{B} N{N} H{H} dh{dh} {layout}", rel_err(lse, rl), 1e-3)
        del ro, rl
    # backward
    for (B, N, H, dh, layout) in [(2, 1024, 4, 192, "seqfirst"), (3, 700, 4, 192, "seqfirst"), (1, 12544, 4, 192, "seqfirst"),
                                  (2, 513, 3, 64, "timm"), (3, 640, 2, 192, "timm"), (2, 700, 4, 96, "seqfirst"),
                                  (2, 600, 4, 48, "seqfirst"), (3, 100, 4, 96, "seqfirst"), (2, 333, 4, 48, "timm"),
                                  (3, 257, 3, 64, "timm"), (4, 197, 3, 256, "timm"), (2, 129, 6, 64, "timm")]:
        qkv, q, k, v, qs, os_, out, perm = _make_qkv(B, N, H, dh, layout)
        lse = torch.empty(B, H, N, device="cuda")
        delta = torch.empty(B, H, N, device="cuda")
        scale = dh ** -0.5
        L.attn_fwd(q, k, v, out, lse, B, H, N, dh, qs, os_, scale)
        dout = torch.randn_like(out.float()).bfloat16()
        dqkv = torch.zeros_like(qkv)
        dq, dk, dv = dqkv.select(2, 0), dqkv.select(2, 1), dqkv.select(2, 2)
        L.attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out, dout, lse, delta, dq.data_ptr(), dk.data_ptr(),
                   dv.data_ptr(), B, H, N, dh, qs, os_, scale)
        torch.cuda.synchronize()
        qr = perm(q).float().detach().requires_grad_(True)
        kr = perm(k).float().detach().requires_grad_(True)
        vr = perm(v).float().detach().requires_grad_(True)
        s = (qr @ kr.transpose(-1, -2)) * scale
        (s.softmax(-1) @ vr).backward(perm(dout).float())
        del s
        report(f"attn bwd dq B{B} N{N} H{H} dh{dh} {layout}", rel_err(perm(dq), qr.grad), 2e-2)
        report(f"attn bwd dk B{B} N{N} H{H} dh{dh} {layout}", rel_err(perm(dk), kr.grad), 2e-2)
        report(f"attn bwd dv B{B} N{N} H{H} dh{dh} {layout}", rel_err(perm(dv), vr.grad), 2e-2)
        del qr, kr, vr
    B, N, H, dh = 15, 12544, 4, 192
    qkv, q, k, v, qs, os_, out, perm = _make_qkv(B, N, H, dh, "seqfirst")
    lse = torch.empty(B, H, N, device="cuda")
    delta = torch.empty(B, H, N, device="cuda")
    dout = torch.randn_like(out.float()).bfloat16()
    dqkv = torch.zeros_like(qkv)
    dq, dk, dv = dqkv.select(2, 0), dqkv.select(2, 1), dqkv.select(2, 2)

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    ms = timeit(lambda: L.attn_fwd(q, k, v, out, lse, B, H, N, dh, qs, os_, dh ** -0.5))
    print(f"  [PERF] attn fwd group_embed shape B15 H4 S12544 dh192: {ms:.2f} ms  {4.0 * B * H * N * N * dh / ms / 1e9:.1f} TFLOP/s", flush=True)
    ms = timeit(lambda: L.attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out, dout, lse, delta, dq.data_ptr(),
                                   dk.data_ptr(), dv.data_ptr(), B, H, N, dh, qs, os_, dh ** -0.5))
    print(f"  [PERF] attn bwd group_embed shape B15 H4 S12544 dh192: {ms:.2f} ms  {10.0 * B * H * N * N * dh / ms / 1e9:.1f} TFLOP/s "
          f"(algorithmic 5 GEMMs; 7 executed)", flush=True)
    for (Bs, Ns, Hs, ds) in [(64, 197, 3, 256), (128, 257, 3, 64), (32, 513, 3, 64)]:
        qkv2, q2, k2, v2, qs2, os2, out2, _ = _make_qkv(Bs, Ns, Hs, ds, "timm")
        lse2 = torch.empty(Bs, Hs, Ns, device="cuda")
        delta2 = torch.empty(Bs, Hs, Ns, device="cuda")
        do2 = torch.randn_like(out2.float()).bfloat16()
        dqkv2 = torch.zeros_like(qkv2)
        ms = timeit(lambda: L.attn_fwd(q2, k2, v2, out2, lse2, Bs, Hs, Ns, ds, qs2, os2, ds ** -0.5), reps=20)
        print(f"  [PERF] attn fwd timm B{Bs} N{Ns} H{Hs} dh{ds}: {ms * 1e3:.1f} us  {4.0 * Bs * Hs * Ns * Ns * ds / ms / 1e9:.1f} TFLOP/s", flush=True)
        ms = timeit(lambda: L.attn_bwd(q2.data_ptr(), k2.data_ptr(), v2.data_ptr(), out2, do2, lse2, delta2,
                                       dqkv2.select(2, 0).data_ptr(), dqkv2.select(2, 1).data_ptr(), dqkv2.select(2, 2).data_ptr(),
                                       Bs, Hs, Ns, ds, qs2, os2, ds ** -0.5), reps=20)
        print(f"  [PERF] attn bwd timm B{Bs} N{Ns} H{Hs} dh{ds}: {ms * 1e3:.1f} us  {10.0 * Bs * Hs * Ns * Ns * ds / ms / 1e9:.1f} TFLOP/s", flush=True)
    seed = torch.tensor([20210915], dtype=torch.int32, device="cuda")
    ms = timeit(lambda: L.attn_fwd(q, k, v, out, lse, B, H, N, dh, qs, os_, dh ** -0.5, drop_seed=seed, drop_site=1, drop_p=0.1))
    print(f"  [PERF] attn fwd group_embed shape, dropout p=0.1: {ms:.2f} ms  {4.0 * B * H * N * N * dh / ms / 1e9:.1f} TFLOP/s", flush=True)
    ms = timeit(lambda: L.attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out, dout, lse, delta, dq.data_ptr(),
                                   dk.data_ptr(), dv.data_ptr(), B, H, N, dh, qs, os_, dh ** -0.5, drop_seed=seed,
                                   drop_site=1, drop_p=0.1))
    print(f"  [PERF] attn bwd group_embed shape, dropout p=0.1: {ms:.2f} ms  {10.0 * B * H * N * N * dh / ms / 1e9:.1f} TFLOP/s "
          f"(algorithmic)", flush=True)


def g_attn_spill():
    """Single-score-pass flash backward (dQ kernel spills P o mask / dS, dK / dV are batched GEMMs): against torch without
    dropout, and against the recomputing two-kernel form with the SAME dropout mask; then the group_embed shape timed in
    both forms."""
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(12)
    old = os.environ.get("S3D_FA_SPILL_MIN_N")
    os.environ["S3D_FA_SPILL_MIN_N"] = "1"
    try:
        seed = torch.tensor([777], dtype=torch.int32, device="cuda")
        for (B, N, H, dh, layout, p_drop) in [(2, 1024, 4, 192, "seqfirst", 0.0), (3, 700, 4, 192, "seqfirst", 0.1),
                                              (1, 1088, 4, 192, "seqfirst", 0.0), (2, 1100, 3, 64, "timm", 0.0),
                                              (3, 257, 3, 64, "timm", 0.1), (2, 700, 4, 96, "seqfirst", 0.0),
                                              (2, 600, 4, 48, "seqfirst", 0.1), (1, 130, 1, 64, "timm", 0.25),
                                              (3, 640, 2, 192, "timm", 0.1), (2, 2176, 4, 192, "seqfirst", 0.1),
                                              (4, 197, 3, 256, "timm", 0.0), (2, 300, 2, 256, "seqfirst", 0.0),
                                              (1, 128, 1, 256, "timm", 0.0)]:
            assert L.lib().s3d_attn_bwd_workspace_bytes(B, H, N, dh) > 0
            qkv, q, k, v, qs, os_, out, perm = _make_qkv(B, N, H, dh, layout)
            lse = torch.empty(B, H, N, device="cuda")
            delta = torch.empty(B, H, N, device="cuda")
            scale = dh ** -0.5
            kw = dict(drop_seed=seed, drop_site=1, drop_p=p_drop) if p_drop > 0 else {}
            L.attn_fwd(q, k, v, out, lse, B, H, N, dh, qs, os_, scale, **kw)
            dout = torch.randn_like(out.float()).bfloat16()
            res = []
            for ws in (True, False):
                dqkv = torch.full_like(qkv, float("nan"))
                dq, dk, dv = dqkv.select(2, 0), dqkv.select(2, 1), dqkv.select(2, 2)
                L.attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out, dout, lse, delta, dq.data_ptr(), dk.data_ptr(),
                           dv.data_ptr(), B, H, N, dh, qs, os_, scale, workspace=ws, **kw)
                torch.cuda.synchronize()
                res.append((perm(dq).float(), perm(dk).float(), perm(dv).float()))
            tag = f"B{B} N{N} H{H} dh{dh} {layout} p{p_drop}"
            for name, a, b in zip(("dq", "dk", "dv"), res[0], res[1]):
                report(f"attn bwd spill vs two-kernel {name} {tag}", rel_err(a, b), 1e-2)
            if p_drop == 0.0:
                qr = perm(q).float().detach().requires_grad_(True)
                kr = perm(k).float().detach().requires_grad_(True)
                vr = perm(v).float().detach().requires_grad_(True)
                s_ = (qr @ kr.transpose(-1, -2)) * scale
                (s_.softmax(-1) @ vr).backward(perm(dout).float())
                del s_
                for name, a, b in zip(("dq", "dk", "dv"), res[0], (qr.grad, kr.grad, vr.grad)):
                    report(f"attn bwd spill vs torch {name} {tag}", rel_err(a, b), 2e-2)
                del qr, kr, vr
    finally:
        if old is None:
            os.environ.pop("S3D_FA_SPILL_MIN_N", None)
        else:
            os.environ["S3D_FA_SPILL_MIN_N"] = old
    if os.environ.get("S3D_PROBE_PERF", "1") == "0":
        return
    B, N, H, dh = 15, 12544, 4, 192
    qkv, q, k, v, qs, os_, out, perm = _make_qkv(B, N, H, dh, "seqfirst")
    lse = torch.empty(B, H, N, device="cuda")
    delta = torch.empty(B, H, N, device="cuda")
    dout = torch.randn_like(out.float()).bfloat16()
    dqkv = torch.zeros_like(qkv)
    dq, dk, dv = dqkv.select(2, 0), dqkv.select(2, 1), dqkv.select(2, 2)
    seed = torch.tensor([20210915], dtype=torch.int32, device="cuda")

    def timeit(fn, reps=4):
        for _ in range(2):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    Bs, Ns, Hs, ds = 64, 197, 3, 256  # stage 2 of cfg3: head_dim 256 -> spill-only form (dQ as a third GEMM) vs mma.sync
    qkv2, q2, k2, v2, qs2, os2, out2, _ = _make_qkv(Bs, Ns, Hs, ds, "timm")
    lse2 = torch.empty(Bs, Hs, Ns, device="cuda")
    delta2 = torch.empty(Bs, Hs, Ns, device="cuda")
    do2 = torch.randn_like(out2.float()).bfloat16()
    dqkv2 = torch.zeros_like(qkv2)
    L.attn_fwd(q2, k2, v2, out2, lse2, Bs, Hs, Ns, ds, qs2, os2, ds ** -0.5)
    for ws in (True, False):
        ms = timeit(lambda: L.attn_bwd(q2.data_ptr(), k2.data_ptr(), v2.data_ptr(), out2, do2, lse2, delta2,
                                       dqkv2.select(2, 0).data_ptr(), dqkv2.select(2, 1).data_ptr(),
                                       dqkv2.select(2, 2).data_ptr(), Bs, Hs, Ns, ds, qs2, os2, ds ** -0.5, workspace=ws), reps=20)
        print(f"  [PERF] attn bwd timm B{Bs} N{Ns} H{Hs} dh{ds} {'tcgen05 spill-only + 3 GEMMs' if ws else 'mma.sync'}: "
              f"{ms * 1e3:.1f} us  {10.0 * Bs * Hs * Ns * Ns * ds / ms / 1e9:.1f} TFLOP/s", flush=True)
    for p_drop in (0.0, 0.1):
        kw = dict(drop_seed=seed, drop_site=1, drop_p=p_drop) if p_drop > 0 else {}
        L.attn_fwd(q, k, v, out, lse, B, H, N, dh, qs, os_, dh ** -0.5, **kw)
        for ws in (True, False):
            ms = timeit(lambda: L.attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out, dout, lse, delta, dq.data_ptr(),
                                           dk.data_ptr(), dv.data_ptr(), B, H, N, dh, qs, os_, dh ** -0.5, workspace=ws, **kw))
            print(f"  [PERF] attn bwd group_embed shape p={p_drop} {'single score pass + 2 GEMMs' if ws else 'two-kernel form'}: "
                  f"{ms:.2f} ms  {10.0 * B * H * N * N * dh / ms / 1e9:.1f} TFLOP/s (algorithmic)", flush=True)


def g_sgemm():
    """fp32 CUDA-core GEMM (point stem, heads) against torch fp32 (highest precision), strided views, split-K, epilogues;
    then the two autograd nodes built on it against the nn modules they replace."""
    import torch
    import torch.nn as nn
    from simple3d_former_b200 import _lib as L
    from simple3d_former_b200 import functional as Fn
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(21)
    dev = "cuda"
    for (M, N, K) in [(131072, 48, 6), (1000, 48, 48), (64, 40, 768), (48, 48, 70000), (1, 48, 131072), (130, 70, 33)]:
        a = torch.randn(M, K, device=dev)
        b = torch.randn(K, N, device=dev)
        bias = torch.randn(N, device=dev)
        report(f"sgemm {M}x{N}x{K}", rel_err(L.sgemm(a, b, bias=bias), a @ b + bias), 2e-5)
        at = torch.randn(K, M, device=dev)
        bt = torch.randn(N, K, device=dev)
        report(f"sgemm strided views {M}x{N}x{K}", rel_err(L.sgemm(at.t(), bt.t()), at.t() @ bt.t()), 2e-5)
    a = torch.randn(300, 20, device=dev)
    b = torch.randn(20, 50, device=dev)
    g = torch.randn(300, 50, device=dev)
    c0 = torch.randn(300, 50, device=dev)
    report("sgemm relu", rel_err(L.sgemm(a, b, relu=True), (a @ b).relu()), 2e-5)
    report("sgemm gate", rel_err(L.sgemm(a, b, gate=g), (a @ b) * (g > 0)), 2e-5)
    c = c0.clone()
    L.sgemm(a, b, out=c, accumulate=True, alpha=0.5)
    report("sgemm accumulate", rel_err(c, c0 + 0.5 * (a @ b)), 2e-5)
    a = torch.randn(50000, 24, device=dev)
    b = torch.randn(50000, 48, device=dev)
    c = c0[:24, :48].clone()
    L.sgemm(a.t(), b, out=c, accumulate=True)
    report("sgemm split-K accumulate", rel_err(c, c0[:24, :48] + a.t() @ b), 2e-5)
    # autograd nodes
    B, Np, dp, q, ncls = 4, 500, 6, 48, 40
    fc1 = nn.Sequential(nn.Linear(dp, q), nn.ReLU(), nn.Linear(q, q)).to(dev)
    fcp = nn.Sequential(nn.Linear(3, q), nn.ReLU(), nn.Linear(q, q)).to(dev)
    head = nn.Linear(q, ncls).to(dev)
    x = torch.randn(B, Np, dp, device=dev)
    w = torch.randn(B, ncls, device=dev)
    params = list(fc1.parameters()) + list(fcp.parameters()) + list(head.parameters())
    (head((fc1(x) + fcp(x[..., :3])).mean(1)) * w).sum().backward()
    want = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    f = Fn.PointStemFn.apply(x, x[..., :3].contiguous(), fc1[0].weight, fc1[0].bias, fc1[2].weight, fc1[2].bias,
                             fcp[0].weight, fcp[0].bias, fcp[2].weight, fcp[2].bias)
    report("PointStemFn forward", rel_err(f, fc1(x) + fcp(x[..., :3])), 2e-5)
    (Fn.LinearF32Fn.apply(f.mean(1), head.weight, head.bias) * w).sum().backward()
    for (n, p), g0 in zip(list(fc1.named_parameters()) + list(fcp.named_parameters()) + list(head.named_parameters()), want):
        report(f"stem / head grad {n} {tuple(p.shape)}", rel_err(p.grad, g0), 5e-5)


def g_attn_fwd():
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(5)
    for (B, N, H, dh, layout) in _attn_cases():
        qkv, q, k, v, qs, os_, out, perm = _make_qkv(B, N, H, dh, layout)
        lse = torch.empty(B, H, N, device="cuda")
        scale = dh ** -0.5
        L.attn_fwd(q, k, v, out, lse, B, H, N, dh, qs, os_, scale)
        torch.cuda.synchronize()
        ro, rl = _attn_ref(perm(q), perm(k), perm(v), scale)
        report(f"attn fwd out B{B} N{N} H{H} dh{dh} {layout}", rel_err(perm(out), ro), 1.5e-2)
        report(f"attn fwd lse B{B} N{N} H{H} dh{dh} {layout}", rel_err(lse, rl), 1e-3)


def g_attn_bwd():
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(6)
    for (B, N, H, dh, layout) in _attn_cases():
        qkv, q, k, v, qs, os_, out, perm = _make_qkv(B, N, H, dh, layout)
        lse = torch.empty(B, H, N, device="cuda")
        delta = torch.empty(B, H, N, device="cuda")
        scale = dh ** -0.5
        L.attn_fwd(q, k, v, out, lse, B, H, N, dh, qs, os_, scale)
        dout = torch.randn_like(out.float()).bfloat16()
        dqkv = torch.zeros_like(qkv)
        dq, dk, dv = dqkv.select(2, 0), dqkv.select(2, 1), dqkv.select(2, 2)
        L.attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out, dout, lse, delta, dq.data_ptr(), dk.data_ptr(),
                   dv.data_ptr(), B, H, N, dh, qs, os_, scale)
        torch.cuda.synchronize()
        qr = perm(q).float().detach().requires_grad_(True)
        kr = perm(k).float().detach().requires_grad_(True)
        vr = perm(v).float().detach().requires_grad_(True)
        s = (qr @ kr.transpose(-1, -2)) * scale
        (s.softmax(-1) @ vr).backward(perm(dout).float())
        report(f"attn bwd dq B{B} N{N} H{H} dh{dh} {layout}", rel_err(perm(dq), qr.grad), 2e-2)
        report(f"attn bwd dk B{B} N{N} H{H} dh{dh} {layout}", rel_err(perm(dk), kr.grad), 2e-2)
        report(f"attn bwd dv B{B} N{N} H{H} dh{dh} {layout}", rel_err(perm(dv), vr.grad), 2e-2)


def g_elementwise():
    import torch
    import torch.nn.functional as F
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(7)
    x = torch.randn(1000, 333, device="cuda")
    report("cast", rel_err(L.cast_bf16(x), x), 5e-3)
    report("transpose f32", rel_err(L.transpose_bf16(x), x.t()), 5e-3)
    report("transpose bf16", rel_err(L.transpose_bf16(x.bfloat16()), x.bfloat16().t()), 0)
    y = torch.randn(5000, 384, device="cuda").bfloat16()
    report("colsum", rel_err(L.colsum(y), y.float().sum(0)), 1e-5)
    # voxel patch gather == conv3d
    for (B, V, c, p, D, zsum) in [(4, 30, 6, 5, 384, True), (2, 128, 9, 14, 768, False), (2, 30, 6, 5, 384, False)]:
        vox = (torch.rand(B, 1, V, V, V, device="cuda") < 0.1).float()
        w = torch.randn(D, 1, c, c, c, device="cuda") * 0.05
        bias = torch.randn(D, device="cuda")
        K = c ** 3
        kpad = (K + 63) // 64 * 64
        P = L.voxel_patch_gather(vox, c, p, kpad, zsum)
        wp = torch.zeros(D, kpad, device="cuda")
        wp[:, :K] = w.reshape(D, K)
        out = L.gemm(P, wp.bfloat16(), bias=bias, alpha=(1.0 / p) if zsum else 1.0, out_dtype=torch.float32)
        ref = F.conv3d(vox, w.bfloat16().float(), bias, stride=c)
        if zsum:
            ref = ref.mean(4).flatten(2).transpose(1, 2).reshape(-1, D)
        else:
            ref = ref.flatten(2).transpose(1, 2).reshape(-1, D)
        report(f"patchify V{V} c{c} p{p} zsum{zsum}", rel_err(out, ref), 2e-3)
        # the patch matrix itself, bit-exact, for every input dtype the loaders produce
        cells = vox[:, 0, :p * c, :p * c, :p * c].reshape(B, p, c, p, c, p, c).permute(0, 1, 3, 5, 2, 4, 6)
        want = (cells.sum(3) if zsum else cells).reshape(-1, K)
        for dt in (torch.float32, torch.uint8, torch.int32):
            Pd = L.voxel_patch_gather(vox.to(dt), c, p, kpad, zsum)
            ok = torch.equal(Pd[:, :K].float(), want) and bool((Pd[:, K:] == 0).all())
            report(f"patch matrix exact V{V} c{c} zsum{zsum} {str(dt)[6:]}", 0.0 if ok else 1.0, 0)
    # adam
    p0 = torch.randn(10001, device="cuda")
    g = torch.randn(10001, device="cuda")
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    m = torch.zeros_like(p0)
    v = torch.zeros_like(p0)
    pm = p0.clone()
    sh = torch.empty(10001, device="cuda", dtype=torch.bfloat16)
    for step in (1, 2, 3):
        pr.grad = g.clone() * step
        opt.step()
        L.adam_step(pm, g * step, m, v, sh, 1e-3, 0.9, 0.999, 1e-8, 0.0, step)
    report("adam 3 steps", rel_err(pm, pr.detach()), 1e-6)
    report("adam shadow", rel_err(sh, pm), 5e-3)


def g_points():
    import torch
    from simple3d_former_b200 import _lib as L
    torch.manual_seed(8)
    for (B, N, S, K) in [(4, 1024, 1024, 16), (3, 1024, 256, 16), (2, 2048, 512, 16), (2, 300, 77, 3), (2, 5000, 100, 16)]:
        xyz = torch.rand(B, N, 3, device="cuda") * 2 - 1
        q = xyz[:, :S].contiguous() if S <= N else torch.rand(B, S, 3, device="cuda")
        idx, dist = L.knn(xyz, q, K, want_dist=True)
        xc, qc = xyz.cpu(), q.cpu()
        d = torch.sum((qc[:, :, None] - xc[:, None]) ** 2, dim=-1)
        ds, ref = d.sort(dim=-1, stable=True)
        ok = torch.equal(idx.cpu(), ref[:, :, :K])
        okd = torch.equal(dist.cpu(), ds[:, :, :K])
        print(f"  [{'PASS' if ok and okd else 'FAIL'}] knn B{B} N{N} S{S} K{K}: idx exact={ok} dist exact={okd}", flush=True)
        r2 = float(torch.tensor(0.2 ** 2, dtype=torch.float32))
        bq = L.ball_query(r2, 16, xyz, q).cpu()
        gi = torch.arange(N).view(1, 1, N).repeat(B, S, 1)
        gi[d > 0.2 ** 2] = N
        gi = gi.sort(dim=-1)[0][:, :, :16]
        first = gi[:, :, 0:1].repeat(1, 1, 16)
        gi[gi == N] = first[gi == N]
        print(f"  [{'PASS' if torch.equal(bq, gi) else 'FAIL'}] ball_query B{B} N{N} S{S}", flush=True)
    for (B, N, npoint) in [(4, 1024, 1024), (4, 1024, 256), (2, 2048, 512), (2, 3000, 64), (1, 8000, 32)]:
        xyz = torch.rand(B, N, 3, device="cuda") * 2 - 1
        start = torch.randint(0, N, (B,), device="cuda")
        got = L.fps(xyz, npoint, start).cpu()
        xc = xyz.cpu()
        cent = torch.zeros(B, npoint, dtype=torch.long)
        distance = torch.ones(B, N) * 1e10
        far = start.cpu().clone()
        bi = torch.arange(B)
        for i in range(npoint):
            cent[:, i] = far
            c = xc[bi, far, :].view(B, 1, 3)
            dd = torch.sum((xc - c) ** 2, -1)
            distance = torch.min(distance, dd)
            far = torch.max(distance, -1)[1]
        print(f"  [{'PASS' if torch.equal(got, cent) else 'FAIL'}] fps B{B} N{N} npoint{npoint}", flush=True)
    pts = torch.randn(3, 100, 51, device="cuda")
    idx = torch.randint(0, 100, (3, 40, 16), device="cuda")
    out = L.gather_rows(pts, idx)
    ref = torch.gather(pts, 1, idx.reshape(3, -1)[..., None].expand(-1, -1, 51))
    print(f"  [{'PASS' if torch.equal(out, ref) else 'FAIL'}] gather_rows", flush=True)
    go = torch.randn(3, 640, 51, device="cuda")
    gp = L.scatter_add_rows(go, idx, 100)
    refg = torch.zeros(3, 100, 51, device="cuda").scatter_add_(1, idx.reshape(3, -1)[..., None].expand(-1, -1, 51), go)
    report("scatter_add_rows", rel_err(gp, refg), 1e-5)


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        import torch
        print(f"== {name} on {torch.cuda.get_device_name(0)}", flush=True)
        globals()["g_" + name]()
        torch.cuda.synchronize()
        return
    def run(g, env_extra=None):
        t0 = time.time()
        env = dict(os.environ)
        env.update(env_extra or {})
        out = ""
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), g], capture_output=True, text=True,
                               timeout=300, env=env)
            out = r.stdout
            print(out, end="")
            if r.returncode != 0:
                out += "FAIL"
                print(f"  [CRASH] group {g} rc={r.returncode}\n" + r.stderr[-1500:])
        except subprocess.TimeoutExpired:
            out += "FAIL"
            print(f"  [TIMEOUT] group {g}")
        print(f"  ({g} {env_extra or ''}: {time.time() - t0:.1f}s)", flush=True)
        return out

    for g in GROUPS:
        out = run(g)
        if "FAIL" in out and g == "gemm_mn":
            # bring-up: try the alternative descriptor offset assignments for MN-major operands
            for lbo, sbo in [(1024, 8192), (8192, 128), (128, 8192), (1024, 1024)]:
                run(g, {"S3D_DBG_MN_LBO": str(lbo), "S3D_DBG_MN_SBO": str(sbo)})
        if "FAIL" in out and g == "gemm_k":
            for lbo, sbo in [(0, 1024), (1024, 1024), (16, 128)]:
                run(g, {"S3D_DBG_K_LBO": str(lbo), "S3D_DBG_K_SBO": str(sbo)})


if __name__ == "__main__":
    main()
