"""ctypes binding of libs3d_b200.so (C ABI declared in include/s3d_b200.h).

The product path has no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
PyTorch is used only to own device memory and streams; every pointer handed to the library is `tensor.data_ptr()`.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs3d_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

# name -> (restype, argtypes); must list every symbol of include/s3d_b200.h (tests/test_abi.py checks this).
_P = c_void_p
SIGNATURES = {
    "s3d_abi_version": (c_int, []),
    "s3d_error_string": (c_char_p, [c_int]),
    "s3d_gemm_bf16": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_float,
                              _P, _P, c_int64, c_int, _P, c_int64, _P, c_int64, c_int, c_int64, c_int64, c_int64,
                              c_int64, c_int, c_int, c_int, _P]),
    "s3d_sgemm_f32": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, _P,
                              c_int, _P, c_int64, c_int, _P]),
    "s3d_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_float, _P]),
    "s3d_layernorm_bwd": (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    "s3d_attn_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int64,
                             c_int64, c_int64, c_float, _P, c_uint32, c_float, _P]),
    "s3d_attn_bwd_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "s3d_attn_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int64, c_int64,
                             c_int64, c_int64, c_int64, c_int64, c_float, _P, c_uint32, c_float, _P, c_int64, _P]),
    "s3d_dropout_add_f32": (c_int, [_P, _P, _P, c_int64, c_int, _P, c_uint32, c_float, _P]),
    "s3d_dropout_bf16": (c_int, [_P, _P, c_int64, c_int, _P, c_uint32, c_float, _P]),
    "s3d_cast_f32_to_bf16": (c_int, [_P, _P, c_int64, _P]),
    "s3d_transpose_to_bf16": (c_int, [_P, c_int, _P, c_int, c_int, c_int64, c_int64, _P]),
    "s3d_colsum_bf16": (c_int, [_P, _P, c_int, c_int, c_int64, c_int, _P]),
    "s3d_voxel_patch_gather": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "s3d_sgd_momentum_step": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_int, _P, c_float, _P]),
    "s3d_adam_step": (c_int, [_P, _P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, _P,
                              c_float, _P]),
    "s3d_knn": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "s3d_ball_query": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_float, c_int, _P]),
    "s3d_fps": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "s3d_gather_rows": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "s3d_scatter_add_rows": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "s3d_sa_group_fwd_stats": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "s3d_sa_group_fwd_act": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_int, _P]),
    "s3d_sa_group_bwd_stats": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P,
                                       c_int, _P]),
    "s3d_sa_group_bwd_scatter": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P,
                                         _P, _P, _P, _P, c_int, _P]),
    "s3d_sa_group_reduce": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P]),
    "s3d_sa_pool_select": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "s3d_sa_dz2_expand": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "s3d_bn_rows_stats": (c_int, [_P, c_int64, c_int, _P, c_int, _P]),
    "s3d_bn_rows_bwd_stats": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, _P, c_int, _P]),
    "s3d_bn_finalize_fwd": (c_int, [_P, c_int, c_int, c_double, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, _P]),
    "s3d_bn_finalize_bwd": (c_int, [_P, c_int, c_int, c_double, c_int, _P, _P, _P, _P, c_int, _P]),
    "s3d_bn_relu_apply": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "s3d_bn_relu_bwd_apply": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "s3d_split_bf16x3": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, _P]),
    "s3d_binvox_scan": (c_int, [_P, _P, _P, _P, _P, c_int, _P]),
    "s3d_binvox_expand": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "s3d_three_nn_interp_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "s3d_three_nn_interp_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
}

_lib = None
LAUNCHES = 0  # number of C-ABI compute calls issued (bench.py reports kernel launches from this)


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libs3d_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libs3d_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the sm_100a CUDA extension is mandatory; there is no CPU or PyTorch fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
        if _lib.s3d_abi_version() != 4:
            raise RuntimeError("libs3d_b200.so ABI version mismatch")
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().s3d_error_string(code)
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else code} (code {code})")


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    """Every kernel is enqueued on the CURRENT device's current stream (see stream()): tensors must be CUDA tensors of
    that device, otherwise the launch would hand foreign pointers to the wrong GPU."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("simple3d_former_b200 ops need CUDA tensors (no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"tensor on {t.device} but the current CUDA device is cuda:{cur}: call "
                               "torch.cuda.set_device() (one process per GPU) before using simple3d_former_b200")


def call(name: str, *args) -> None:
    global LAUNCHES
    LAUNCHES += 1
    check(getattr(lib(), name)(*args), name)


# ------------------------------------------------------------------------------------------------------------------
# thin tensor-level wrappers (shape checks + pointer extraction only)
# ------------------------------------------------------------------------------------------------------------------
EPI_NONE, EPI_GELU, EPI_DGELU, EPI_RELU, EPI_DRELU = 0, 1, 2, 3, 4


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16, alpha=1.0, bias=None, residual=None,
         epilogue=EPI_NONE, aux_in=None, aux_out=None, force_bn=0, force_cluster=0, force_splits=0):
    """D[M,N] = epi(alpha * A @ B^T).  a: [M,K] (or [K,M] if a_mn); b: [N,K] (or [K,N] if b_mn); 2-D or batched 3-D."""
    _need_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    batched = a.dim() == 3
    if batched:
        assert b.dim() == 3 and a.shape[0] == b.shape[0]
        batch = a.shape[0]
        a2, b2 = a[0], b[0]
    else:
        batch = 1
        a2, b2 = a, b
    assert a2.stride(-1) == 1 and b2.stride(-1) == 1
    if a_mn:
        K, M = a2.shape
    else:
        M, K = a2.shape
    if b_mn:
        Kb, N = b2.shape
    else:
        N, Kb = b2.shape
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    if out is None:
        shape = (batch, M, N) if batched else (M, N)
        out = torch.empty(shape, device=a.device, dtype=out_dtype)
    o2 = out[0] if batched else out
    assert o2.shape == (M, N) and o2.stride(-1) == 1
    r2 = None
    if residual is not None:
        assert residual.dtype == torch.float32
        r2 = residual[0] if batched else residual
        assert r2.shape == (M, N) and r2.stride(-1) == 1
    call("s3d_gemm_bf16", ptr(a), ptr(b), ptr(out), M, N, K, a2.stride(0), b2.stride(0), o2.stride(0), int(a_mn),
         int(b_mn), int(out.dtype == torch.float32), float(alpha), ptr(bias), ptr(residual),
         r2.stride(0) if r2 is not None else 0, epilogue, ptr(aux_in), aux_in.stride(0) if aux_in is not None else 0,
         ptr(aux_out), aux_out.stride(0) if aux_out is not None else 0, batch,
         a.stride(0) if batched else 0, b.stride(0) if batched else 0, out.stride(0) if batched else 0,
         residual.stride(0) if (batched and residual is not None) else 0, force_bn, force_cluster, force_splits,
         stream())
    return out


def sgemm(a, b, *, out=None, bias=None, relu=False, gate=None, accumulate=False, alpha=1.0):
    """fp32 C[M,N] (+)= alpha * a[M,K] @ b[K,N] (+ bias) (ReLU) (zeroed where gate <= 0) on the CUDA cores; a and b may be
    arbitrary 2-D strided views (x.t() costs nothing). The thin fp32 layers of the path (point stem, heads)."""
    _need_cuda(a, b)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == 2 and b.dim() == 2
    M, K = a.shape
    K2, N = b.shape
    assert K == K2
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.shape == (M, N) and out.stride(1) == 1
    if gate is not None:
        assert gate.dtype == torch.float32 and gate.shape == (M, N) and gate.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    call("s3d_sgemm_f32", ptr(a), ptr(b), ptr(out), M, N, K, a.stride(0), a.stride(1), b.stride(0), b.stride(1),
         out.stride(0), float(alpha), ptr(bias), int(relu), ptr(gate), gate.stride(0) if gate is not None else 0,
         int(accumulate), stream())
    return out


def layernorm_fwd(x, gamma, beta, eps, *, addend=None, want_sum=False, want_bf16=True, want_f32=False, want_stats=True):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    D = x.shape[-1]
    T = x.numel() // D
    y16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    y32 = torch.empty_like(x) if want_f32 else None
    s = torch.empty_like(x) if (want_sum and addend is not None) else None
    mean = torch.empty(T, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(T, device=x.device, dtype=torch.float32) if want_stats else None
    if addend is not None:
        assert addend.dtype == torch.float32 and addend.is_contiguous() and addend.shape == x.shape
    call("s3d_layernorm_fwd", ptr(x), ptr(addend), ptr(s), ptr(gamma), ptr(beta), ptr(y16), ptr(y32), ptr(mean),
         ptr(rstd), T, D, float(eps), stream())
    return y16, y32, s, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, *, dres=None, want_bf16=False, dgamma=None, dbeta=None, dxsum=None,
                  want_dxsum=False):
    """dxsum: f32 [D] that the column sums of dx are ACCUMULATED into (bias gradient of the Linear whose output gradient
    dx is); want_dxsum=True allocates a zeroed one. Returns (dx, dx_bf16, dgamma, dbeta[, dxsum])."""
    _need_cuda(dy, x)
    assert x.dtype == torch.float32 and x.is_contiguous() and dy.is_contiguous()
    D = x.shape[-1]
    T = x.numel() // D
    dx = torch.empty_like(x)
    dx16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    need = (dgamma is None) + (dgamma is None) + (want_dxsum and dxsum is None)
    if need:
        z = torch.zeros((need, D), device=x.device, dtype=torch.float32)  # one fill for every fresh accumulator
        i = 0
        if dgamma is None:
            dgamma, dbeta = z[0], z[1]
            i = 2
        if want_dxsum and dxsum is None:
            dxsum = z[i]
    call("s3d_layernorm_bwd", ptr(dy), int(dy.dtype == torch.bfloat16), ptr(x), ptr(gamma), ptr(mean), ptr(rstd),
         ptr(dres), ptr(dx), ptr(dx16), ptr(dgamma), ptr(dbeta), ptr(dxsum), T, D, stream())
    if want_dxsum or dxsum is not None:
        return dx, dx16, dgamma, dbeta, dxsum
    return dx, dx16, dgamma, dbeta


def attn_fwd(q, k, v, out, lse, B, H, N, dh, qs, os_, scale, drop_seed=None, drop_site=0, drop_p=0.0):
    """drop_seed: int32 CUDA tensor [1] (device-resident seed) enabling attention-probability dropout with rate drop_p."""
    _need_cuda(out, lse, drop_seed)
    call("s3d_attn_fwd", ptr(q) if hasattr(q, "data_ptr") else q, ptr(k) if hasattr(k, "data_ptr") else k,
         ptr(v) if hasattr(v, "data_ptr") else v, ptr(out), ptr(lse), B, H, N, dh, qs[0], qs[1], qs[2], os_[0], os_[1],
         os_[2], float(scale), ptr(drop_seed), int(drop_site), float(drop_p), stream())


_ATTN_WS_MAX = int(float(os.environ.get("S3D_ATTN_WS_MAX_GB", "64")) * (1 << 30))


def attn_bwd(q, k, v, out, dout, lse, delta, dq, dk, dv, B, H, N, dh, qs, os_, scale, drop_seed=None, drop_site=0,
             drop_p=0.0, workspace=True):
    """workspace=True: long sequences run the single-score-pass backward (the library says how much scratch it wants for
    the shape; it is allocated here for the duration of the call, up to S3D_ATTN_WS_MAX_GB, default 64 GiB);
    workspace=False forces the recomputing two-kernel form."""
    _need_cuda(out, dout, lse, delta, drop_seed)
    ws, ws_ptr, ws_bytes = None, None, 0
    if workspace:
        need = int(lib().s3d_attn_bwd_workspace_bytes(B, H, N, dh))
        if 0 < need <= _ATTN_WS_MAX:
            ws = torch.empty(need + 1024, device=out.device, dtype=torch.uint8)
            ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
            ws_bytes = need
    call("s3d_attn_bwd", q, k, v, ptr(out), ptr(dout), ptr(lse), ptr(delta), dq, dk, dv, B, H, N, dh, qs[0], qs[1],
         qs[2], os_[0], os_[1], os_[2], float(scale), ptr(drop_seed), int(drop_site), float(drop_p), ws_ptr, ws_bytes,
         stream())
    del ws  # allocated and consumed on the current stream: the caching allocator reuses it in stream order


def dropout_add(x, residual, seed, site, p):
    """out = residual + mask * x / (1 - p) on fp32 [rows, cols] (counter-based mask keyed by the device seed)."""
    _need_cuda(x, seed)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.is_contiguous() and residual.shape == x.shape
    out = torch.empty_like(x)
    call("s3d_dropout_add_f32", ptr(x), ptr(residual), ptr(out), x.shape[0], x.shape[1], ptr(seed), int(site), float(p),
         stream())
    return out


def dropout_bf16(x, seed, site, p, inplace=False):
    _need_cuda(x, seed)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 2
    out = x if inplace else torch.empty_like(x)
    call("s3d_dropout_bf16", ptr(x), ptr(out), x.shape[0], x.shape[1], ptr(seed), int(site), float(p), stream())
    return out


def cast_bf16(x, out=None):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    call("s3d_cast_f32_to_bf16", ptr(x), ptr(out), x.numel(), stream())
    return out


def transpose_bf16(x, out=None):
    """out[C,R] = bf16(x[R,C]) for a 2-D f32/bf16 matrix."""
    _need_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    R, C = x.shape
    if out is None:
        out = torch.empty((C, R), device=x.device, dtype=torch.bfloat16)
    call("s3d_transpose_to_bf16", ptr(x), int(x.dtype == torch.bfloat16), ptr(out), R, C, x.stride(0), out.stride(0),
         stream())
    return out


def colsum(x, out=None, accumulate=False):
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    T, C = x.shape
    if out is None:
        out = torch.empty(C, device=x.device, dtype=torch.float32)
        accumulate = False
    call("s3d_colsum_bf16", ptr(x), ptr(out), T, C, x.stride(0), int(accumulate), stream())
    return out


_VOXEL_DTYPES = {torch.float32: 0, torch.uint8: 1, torch.bool: 1, torch.int32: 2}


def voxel_patch_gather(x, cell, patch, kpad, zsum):
    _need_cuda(x)
    assert x.dtype in _VOXEL_DTYPES and x.is_contiguous() and x.dim() == 5 and x.shape[1] == 1
    B, _, V, _, _ = x.shape
    rows = B * patch * patch * (1 if zsum else patch)
    P = torch.empty((rows, kpad), device=x.device, dtype=torch.bfloat16)
    call("s3d_voxel_patch_gather", ptr(x), _VOXEL_DTYPES[x.dtype], ptr(P), B, V, cell, patch, kpad, int(zsum), stream())
    return P


def sgd_momentum_step(p, g, buf, shadow, lr, momentum, weight_decay, step, grad_scale=1.0, step_tensor=None):
    _need_cuda(p, g, buf)
    call("s3d_sgd_momentum_step", ptr(p), ptr(g), ptr(buf), ptr(shadow), p.numel(), float(lr), float(momentum),
         float(weight_decay), int(step), ptr(step_tensor), float(grad_scale), stream())


def adam_step(p, g, m, v, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, step_tensor=None):
    _need_cuda(p, g, m, v)
    call("s3d_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), ptr(shadow), p.numel(), float(lr), float(beta1), float(beta2),
         float(eps), float(weight_decay), int(step), ptr(step_tensor), float(grad_scale), stream())


def knn(xyz, query, K, want_dist=False):
    _need_cuda(xyz, query)
    xyz = xyz.contiguous().float()
    query = query.contiguous().float()
    B, N, _ = xyz.shape
    S = query.shape[1]
    idx = torch.empty((B, S, K), device=xyz.device, dtype=torch.long)
    dist = torch.empty((B, S, K), device=xyz.device, dtype=torch.float32) if want_dist else None
    call("s3d_knn", ptr(xyz), ptr(query), ptr(idx), ptr(dist), B, N, S, K, stream())
    return (idx, dist) if want_dist else idx


def ball_query(radius_sq_f32, nsample, xyz, query):
    _need_cuda(xyz, query)
    xyz = xyz.contiguous().float()
    query = query.contiguous().float()
    B, N, _ = xyz.shape
    S = query.shape[1]
    idx = torch.empty((B, S, nsample), device=xyz.device, dtype=torch.long)
    call("s3d_ball_query", ptr(xyz), ptr(query), ptr(idx), B, N, S, float(radius_sq_f32), nsample, stream())
    return idx


def fps(xyz, npoint, start):
    _need_cuda(xyz, start)
    xyz = xyz.contiguous().float()
    B, N, _ = xyz.shape
    start = start.contiguous().long()
    idx = torch.empty((B, npoint), device=xyz.device, dtype=torch.long)
    call("s3d_fps", ptr(xyz), ptr(start), ptr(idx), B, N, npoint, stream())
    return idx


def gather_rows(points, idx):
    _need_cuda(points, idx)
    B, N, C = points.shape
    M = idx.numel() // B
    out = torch.empty((B, M, C), device=points.device, dtype=torch.float32)
    call("s3d_gather_rows", ptr(points), ptr(idx), ptr(out), B, N, M, C, stream())
    return out


def scatter_add_rows(grad_out, idx, N):
    _need_cuda(grad_out, idx)
    B, M, C = grad_out.shape
    gp = torch.empty((B, N, C), device=grad_out.device, dtype=torch.float32)
    call("s3d_scatter_add_rows", ptr(grad_out), ptr(idx), ptr(gp), B, N, M, C, stream())
    return gp


# ------------------------------------------------------------------------------------------------------------------
# set abstraction / feature propagation (csrc/pointnet_fused.cu)
# ------------------------------------------------------------------------------------------------------------------
def _slots(work_items, C):
    """Number of partial-sum slots (CTA row-slices) for a statistics pass over `work_items` warp items of C channels."""
    nchunks = (C + 127) // 128
    # 148 SMs x 4 resident CTAs (256 threads, <= 64 registers): one full wave, no half-empty second wave
    return int(max(1, min((work_items + 7) // 8, 592 // nchunks)))


def _f32(*shape, device):
    return torch.empty(shape, device=device, dtype=torch.float32)


class SaGroup:
    """Arguments shared by the layer-1 passes of a set-abstraction layer (see include/s3d_b200.h)."""

    def __init__(self, uf, xyz, cxyz, idx, w1):
        _need_cuda(uf, xyz, cxyz, idx, w1)
        B, N, _ = xyz.shape
        S, K = idx.shape[1], idx.shape[2]
        C1 = uf.shape[-1]
        assert uf.dtype == torch.float32 and uf.is_contiguous() and uf.numel() == B * N * C1
        assert xyz.dtype == torch.float32 and xyz.is_contiguous() and cxyz.dtype == torch.float32 and cxyz.is_contiguous()
        assert cxyz.shape == (B, S, 3) and idx.dtype == torch.long and idx.is_contiguous()
        assert w1.dtype == torch.float32 and w1.dim() == 2 and w1.stride(1) == 1 and w1.shape[0] == C1
        self.t = (uf, xyz, cxyz, idx, w1)
        self.dims = (B, N, S, K, C1)
        self.P = _slots(B * S, C1)
        self.head = (ptr(uf), ptr(xyz), ptr(cxyz), ptr(idx), ptr(w1), w1.stride(0), B, N, S, K, C1)
        self.device = uf.device


def sa_group_fwd_stats(g):
    part = _f32(g.P, 2, g.dims[4], device=g.device)
    call("s3d_sa_group_fwd_stats", *g.head, ptr(part), g.P, stream())
    return part


def sa_group_fwd_act(g, scale, shift):
    B, N, S, K, C1 = g.dims
    a1 = torch.empty((B * S * K, C1), device=g.device, dtype=torch.bfloat16)
    call("s3d_sa_group_fwd_act", *g.head, ptr(scale), ptr(shift), ptr(a1), g.P, stream())
    return a1


def sa_group_bwd_stats(g, mean, rstd, a1, da1):
    part = _f32(g.P, 2, g.dims[4], device=g.device)
    assert a1.is_contiguous() and da1.is_contiguous() and da1.shape == a1.shape and da1.dtype == torch.bfloat16
    call("s3d_sa_group_bwd_stats", *g.head, ptr(mean), ptr(rstd), ptr(a1), ptr(da1), ptr(part), g.P, stream())
    return part


def sa_group_bwd_scatter(g, scale, mean, rstd, m1, m2, a1, da1):
    B, N, S, K, C1 = g.dims
    duf = torch.zeros((B * N, C1), device=g.device, dtype=torch.float32)
    P = max(1, min(g.P, 444 // ((C1 + 127) // 128)))  # this pass holds 80 registers: 3 resident CTAs per SM
    part = _f32(P, 3, C1, device=g.device)
    call("s3d_sa_group_bwd_scatter", *g.head, ptr(scale), ptr(mean), ptr(rstd), ptr(m1), ptr(m2), ptr(a1), ptr(da1),
         ptr(duf), ptr(part), P, stream())
    return duf, part


def sa_group_reduce(z2, G, K):
    _need_cuda(z2)
    assert z2.dtype == torch.float32 and z2.is_contiguous() and z2.shape[0] == G * K
    C = z2.shape[1]
    dev = z2.device
    P = _slots(G, C)
    zmax, zmin = _f32(G, C, device=dev), _f32(G, C, device=dev)
    kmax = torch.empty((G, C), device=dev, dtype=torch.uint8)
    kmin = torch.empty((G, C), device=dev, dtype=torch.uint8)
    part = _f32(P, 2, C, device=dev)
    call("s3d_sa_group_reduce", ptr(z2), G, K, C, ptr(zmax), ptr(zmin), ptr(kmax), ptr(kmin), ptr(part), P, stream())
    return zmax, zmin, kmax, kmin, part


def sa_pool_select(zmax, zmin, kmax, kmin, scale, shift):
    G, C = zmax.shape
    out, zsel = torch.empty_like(zmax), torch.empty_like(zmax)
    ksel = torch.empty_like(kmax)
    call("s3d_sa_pool_select", ptr(zmax), ptr(zmin), ptr(kmax), ptr(kmin), ptr(scale), ptr(shift), ptr(out), ptr(zsel),
         ptr(ksel), G, C, stream())
    return out, zsel, ksel


def sa_dz2_expand(z2, dout, zsel, ksel, scale, shift, mean, rstd, m1, m2, G, K):
    C = z2.shape[1]
    assert dout.dtype == torch.float32 and dout.is_contiguous() and dout.shape == (G, C)
    dz2 = torch.empty((G * K, C), device=z2.device, dtype=torch.bfloat16)
    call("s3d_sa_dz2_expand", ptr(z2), ptr(dout), ptr(zsel), ptr(ksel), ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
         ptr(m1), ptr(m2), ptr(dz2), G, K, C, _slots(G, C), stream())
    return dz2


def bn_rows_stats(z):
    _need_cuda(z)
    assert z.dtype == torch.float32 and z.dim() == 2 and z.is_contiguous()
    R, C = z.shape
    P = _slots(R, C)
    part = _f32(P, 2, C, device=z.device)
    call("s3d_bn_rows_stats", ptr(z), R, C, ptr(part), P, stream())
    return part


def bn_rows_bwd_stats(dout, z, scale, shift, mean, rstd):
    _need_cuda(dout, z)
    assert z.dtype == torch.float32 and z.dim() == 2 and z.is_contiguous()
    assert dout.dtype == torch.float32 and dout.is_contiguous() and dout.shape == z.shape
    R, C = z.shape
    P = _slots(R, C)
    part = _f32(P, 2, C, device=z.device)
    call("s3d_bn_rows_bwd_stats", ptr(dout), ptr(z), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), R, C, ptr(part), P,
         stream())
    return part


def bn_finalize_fwd(part, count, gamma, beta, eps, momentum, running_mean, running_var):
    P, _, C = part.shape
    dev = part.device
    mean, rstd, scale, shift = (_f32(C, device=dev) for _ in range(4))
    call("s3d_bn_finalize_fwd", ptr(part), P, C, float(count), ptr(gamma), ptr(beta), float(eps), float(momentum),
         ptr(running_mean), ptr(running_var), ptr(mean), ptr(rstd), ptr(scale), ptr(shift), stream())
    return mean, rstd, scale, shift


def bn_finalize_bwd(part, count, training):
    P, _, C = part.shape
    dev = part.device
    m1, m2, dgamma, dbeta = (_f32(C, device=dev) for _ in range(4))
    call("s3d_bn_finalize_bwd", ptr(part), P, C, float(count), int(training), ptr(m1), ptr(m2), ptr(dgamma), ptr(dbeta),
         0, stream())
    return m1, m2, dgamma, dbeta


def bn_relu_apply(z, scale, shift, want_f32=True, want_bf16=False):
    R, C = z.shape
    y32 = torch.empty_like(z) if want_f32 else None
    y16 = torch.empty(z.shape, device=z.device, dtype=torch.bfloat16) if want_bf16 else None
    call("s3d_bn_relu_apply", ptr(z), ptr(scale), ptr(shift), ptr(y32), ptr(y16), R, C, stream())
    return y32, y16


def bn_relu_bwd_apply(dout, z, scale, shift, mean, rstd, m1, m2):
    R, C = z.shape
    dz16 = torch.empty(z.shape, device=z.device, dtype=torch.bfloat16)
    call("s3d_bn_relu_bwd_apply", ptr(dout), ptr(z), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(m1), ptr(m2),
         ptr(dz16), R, C, stream())
    return dz16


def three_nn_interp_fwd(feats, idx, dist, addend=None):
    _need_cuda(feats, idx, dist)
    B, S, C = feats.shape
    N = idx.shape[1]
    assert feats.dtype == torch.float32 and feats.is_contiguous() and idx.is_contiguous() and dist.is_contiguous()
    assert idx.shape == (B, N, 3) and dist.shape == (B, N, 3)
    if addend is not None:
        assert addend.dtype == torch.float32 and addend.is_contiguous() and addend.shape == (B, N, C)
    out = _f32(B, N, C, device=feats.device)
    call("s3d_three_nn_interp_fwd", ptr(feats), ptr(idx), ptr(dist), ptr(addend), ptr(out), B, S, N, C, stream())
    return out


def three_nn_interp_bwd(dout, idx, dist, S):
    B, N, C = dout.shape
    assert dout.dtype == torch.float32 and dout.is_contiguous()
    dfeats = _f32(B, S, C, device=dout.device)
    call("s3d_three_nn_interp_bwd", ptr(dout), ptr(idx), ptr(dist), ptr(dfeats), B, S, N, C, stream())
    return dfeats


def split_bf16x3(x, weight_layout=False):
    """fp32 [R,K] -> bf16 [R,3K] hi/lo split ([hi|lo|hi] for activations, [hi|hi|lo] for weights)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    R, K = x.shape
    out = torch.empty((R, 3 * K), device=x.device, dtype=torch.bfloat16)
    call("s3d_split_bf16x3", ptr(x), ptr(out), R, K, x.stride(0), int(weight_layout), stream())
    return out


def binvox_expand(payload, offsets, V, out_dtype=torch.uint8, fix_coords=True):
    """payload uint8 [bytes] (value, count) pairs of B models back to back, offsets int64 [B+1] (even byte offsets).
    Returns (grid [B,1,V,V,V] of out_dtype, totals int64 [B] = voxels encoded per model)."""
    _need_cuda(payload, offsets)
    assert payload.dtype == torch.uint8 and payload.is_contiguous() and offsets.dtype == torch.long
    B = offsets.numel() - 1
    dev = payload.device
    run_offsets = offsets // 2
    run_end = torch.empty(max(payload.numel() // 2, 1), device=dev, dtype=torch.int32)
    totals = torch.empty(B, device=dev, dtype=torch.long)
    call("s3d_binvox_scan", ptr(payload), ptr(offsets), ptr(run_end), ptr(run_offsets), ptr(totals), B, stream())
    out = torch.empty((B, 1, V, V, V), device=dev, dtype=out_dtype)
    call("s3d_binvox_expand", ptr(payload), ptr(offsets), ptr(run_end), ptr(run_offsets), ptr(out),
         _VOXEL_DTYPES[out_dtype], B, V, int(fix_coords), stream())
    return out, totals
