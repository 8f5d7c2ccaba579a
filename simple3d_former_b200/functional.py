"""torch.autograd.Function wrappers that chain the C-ABI kernels into the fused pieces of the hot path.

Numerics contract: bf16 storage for GEMM/attention operands, fp32 accumulation, fp32 residual stream, fp32 LayerNorm
statistics and softmax, fp32 parameter gradients. PyTorch only allocates tensors and records the autograd graph here.
"""
from __future__ import annotations

import os

import torch

from . import _lib as L

_EPS_BLOCK = 1e-6


# ----------------------------------------------------------------------------------------------------------------
# bf16 weight shadows
# ----------------------------------------------------------------------------------------------------------------
def shadow(p: torch.Tensor) -> torch.Tensor:
    """bf16 copy of an fp32 parameter (2-D+ weights are viewed as [out, -1]).

    If a trainer manages a flat shadow buffer it sets `p._s3d_shadow` (kept fresh by the fused Adam kernel); otherwise
    the copy is cached per parameter and refreshed when the parameter's version counter or storage changes."""
    s = getattr(p, "_s3d_shadow", None)
    if s is not None:
        return s
    key = (p._version, p.data_ptr())
    cache = getattr(p, "_s3d_shadow_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    src = p.detach()
    if not src.is_contiguous():
        src = src.contiguous()
    s16 = L.cast_bf16(src.float() if src.dtype != torch.float32 else src)
    try:
        p._s3d_shadow_cache = (key, s16)
    except Exception:
        pass
    return s16


# ----------------------------------------------------------------------------------------------------------------
# gradient sinks: under DataParallelTrainer every parameter owns a view of the flat fp32 gradient buffer
# (`p._s3d_grad_sink`); weight / bias / LayerNorm gradients are then accumulated IN PLACE by the producing kernel
# (GEMM epilogue accumulate or split-K red.add, colsum accumulate, LN-bwd atomics) and backward() returns None for them:
# no temporary dW tensors, no autograd AccumulateGrad add kernels, no zero-fill kernels.
# ----------------------------------------------------------------------------------------------------------------
def _sink(p):
    return getattr(p, "_s3d_grad_sink", None) if p is not None else None


def _wgrad(w, a16, b16, alpha=1.0):
    """dW (+)= alpha * a16^T @ b16 (both operands consumed MN-major)."""
    sk = _sink(w)
    if sk is None:
        return L.gemm(a16, b16, a_mn=True, b_mn=True, alpha=alpha, out_dtype=torch.float32).view(w.shape)
    s2 = sk.view(sk.shape[0], -1)
    L.gemm(a16, b16, a_mn=True, b_mn=True, alpha=alpha, out=s2, residual=s2)
    w._s3d_owner.note_write(w)
    return None


def _bgrad(b, x16):
    """db (+)= column sums of x16."""
    if b is None:
        return None
    sk = _sink(b)
    if sk is None:
        return L.colsum(x16)
    L.colsum(x16, out=sk, accumulate=True)
    b._s3d_owner.note_write(b)
    return None


def _ln_bwd(dy, x, gamma, beta, mean, rstd, **kw):
    """LayerNorm backward; returns (dx, dx16, dgamma, dbeta) with None parameter gradients when they went to sinks."""
    sg, sb = _sink(gamma), _sink(beta)
    if sg is not None and sb is not None:
        dx, dx16, _, _ = L.layernorm_bwd(dy, x, gamma, mean, rstd, dgamma=sg, dbeta=sb, **kw)
        gamma._s3d_owner.note_write(gamma)
        beta._s3d_owner.note_write(beta)
        return dx, dx16, None, None
    return L.layernorm_bwd(dy, x, gamma, mean, rstd, **kw)


def _ln_bwd_colsum(dy, x, gamma, beta, mean, rstd, bias, **kw):
    """LayerNorm backward whose dx is the output gradient of a Linear layer with bias parameter `bias`: the kernel also
    accumulates the column sums of dx, i.e. that bias gradient, so no separate pass over dx is needed. `bias` = None asks
    for a fresh accumulator (the caller hands it to whoever owns the bias). Returns (dx, dx16, dgamma, dbeta, dbias)
    where dbias is None when it went straight into the bias parameter's gradient sink."""
    sg, sb = _sink(gamma), _sink(beta)
    sk = _sink(bias)
    args = dict(kw)
    if sk is not None:
        args["dxsum"] = sk
    else:
        args["want_dxsum"] = True
    if sg is not None and sb is not None:
        dx, dx16, _, _, dxsum = L.layernorm_bwd(dy, x, gamma, mean, rstd, dgamma=sg, dbeta=sb, **args)
        gamma._s3d_owner.note_write(gamma)
        beta._s3d_owner.note_write(beta)
        dg = db = None
    else:
        dx, dx16, dg, db, dxsum = L.layernorm_bwd(dy, x, gamma, mean, rstd, **args)
    if sk is not None:
        bias._s3d_owner.note_write(bias)
        dxsum = None
    return dx, dx16, dg, db, dxsum


def _w2d(w16: torch.Tensor) -> torch.Tensor:
    return w16.reshape(w16.shape[0], -1)


# ----------------------------------------------------------------------------------------------------------------
# Side stream for the parameter-gradient work of small problems. In a Block's backward only 7 of the 17 launches sit on
# the activation-gradient chain; the 4 weight-gradient GEMMs, 4 bias column sums (and their split-K memsets) depend on
# it but nothing depends on them. At the cfg2 sizes (T*D = 0.64 M) every kernel occupies 40-80 of the 148 SMs for
# 5-12 us, so the off-chain half runs concurrently on a second stream (fork after each producer, one join before the node
# returns: every tensor the side stream touches is alive until then). Large problems fill the machine with every
# kernel and keep the single-stream order.
# ----------------------------------------------------------------------------------------------------------------
_SIDE_STREAMS = {}
_OVERLAP_LIMIT = int(os.environ.get("S3D_OVERLAP_MAX_ELEMS", str(8 * 1024 * 1024)))  # token * channel elements


class _ParamGradStream:
    """fork(): work issued under `with ctx.fork():` runs on the side stream after everything enqueued so far on the
    current stream; join(): the current stream waits for all of it. A no-op object when overlap is off."""

    def __init__(self, device, enabled):
        self.side = None
        if enabled:
            key = (device.type, device.index)
            if key not in _SIDE_STREAMS:
                _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
            self.side = _SIDE_STREAMS[key]
            self.main = torch.cuda.current_stream(device)

    def fork(self):
        if self.side is None:
            return _NullCtx()
        self.side.wait_stream(self.main)
        return torch.cuda.stream(self.side)

    def join(self):
        if self.side is not None:
            self.main.wait_stream(self.side)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _attn_strides_timm(N, H, dh):
    E = H * dh
    return (N * 3 * E, dh, 3 * E), (N * E, dh, E)


# ----------------------------------------------------------------------------------------------------------------
# Linear: y = x W^T + b on the tensor cores (used by the stand-alone Attention / Mlp modules)
# ----------------------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        K = x.shape[-1]
        x16 = L.cast_bf16(x.reshape(-1, K).contiguous().float())
        w16 = _w2d(shadow(weight))
        y = L.gemm(x16, w16, bias=bias, out_dtype=torch.float32)
        ctx.save_for_backward(x16, weight)
        ctx.has_bias = bias is not None
        ctx.bias_ref = bias
        ctx.in_shape = x.shape
        return y.reshape(*x.shape[:-1], w16.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x16, weight = ctx.saved_tensors
        w16 = _w2d(shadow(weight))
        dy16 = L.cast_bf16(dy.reshape(-1, dy.shape[-1]).contiguous())
        dx = L.gemm(dy16, w16, b_mn=True, out_dtype=torch.float32).reshape(ctx.in_shape)
        dw = _wgrad(weight, dy16, x16)
        db = _bgrad(ctx.bias_ref, dy16) if ctx.has_bias else None
        return dx, dw, db


class MlpFn(torch.autograd.Function):
    """timm Mlp.forward: fc2(GELU_erf(fc1(x))) with bias+GELU and bias epilogues fused into the two GEMMs."""

    @staticmethod
    def forward(ctx, x, fc1_w, fc1_b, fc2_w, fc2_b):
        K = x.shape[-1]
        x16 = L.cast_bf16(x.reshape(-1, K).contiguous().float())
        w1, w2 = shadow(fc1_w), shadow(fc2_w)
        pre = torch.empty((x16.shape[0], w1.shape[0]), device=x.device, dtype=torch.bfloat16)
        a16 = L.gemm(x16, w1, bias=fc1_b, epilogue=L.EPI_GELU, aux_out=pre)
        y = L.gemm(a16, w2, bias=fc2_b, out_dtype=torch.float32)
        ctx.save_for_backward(x16, pre, a16, fc1_w, fc2_w)
        ctx.meta = (x.shape, fc1_b is not None, fc2_b is not None)
        return y.reshape(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x16, pre, a16, fc1_w, fc2_w = ctx.saved_tensors
        in_shape, has_b1, has_b2 = ctx.meta
        dy16 = L.cast_bf16(dy.reshape(-1, dy.shape[-1]).contiguous())
        dfc2_w = L.gemm(dy16, a16, a_mn=True, b_mn=True, out_dtype=torch.float32)
        dfc2_b = L.colsum(dy16) if has_b2 else None
        dpre = L.gemm(dy16, shadow(fc2_w), b_mn=True, epilogue=L.EPI_DGELU, aux_in=pre)
        dfc1_w = L.gemm(dpre, x16, a_mn=True, b_mn=True, out_dtype=torch.float32)
        dfc1_b = L.colsum(dpre) if has_b1 else None
        dx = L.gemm(dpre, shadow(fc1_w), b_mn=True, out_dtype=torch.float32).reshape(in_shape)
        return dx, dfc1_w, dfc1_b, dfc2_w, dfc2_b


# ----------------------------------------------------------------------------------------------------------------
# Attention core on [B, N, 3, H, dh] bf16 (timm layout)
# ----------------------------------------------------------------------------------------------------------------
def _attn_core_fwd(qkv16, B, N, H, dh, scale):
    E = H * dh
    out = torch.empty((B, N, E), device=qkv16.device, dtype=torch.bfloat16)
    lse = torch.empty((B, H, N), device=qkv16.device, dtype=torch.float32)
    qs, os_ = _attn_strides_timm(N, H, dh)
    base = qkv16.data_ptr()
    L.attn_fwd(base, base + 2 * E, base + 4 * E, out, lse, B, H, N, dh, qs, os_, scale)
    return out, lse


def _attn_core_bwd(qkv16, out16, dout16, lse, B, N, H, dh, scale):
    E = H * dh
    dqkv = torch.empty_like(qkv16)
    delta = torch.empty_like(lse)
    qs, os_ = _attn_strides_timm(N, H, dh)
    base, dbase = qkv16.data_ptr(), dqkv.data_ptr()
    L.attn_bwd(base, base + 2 * E, base + 4 * E, out16, dout16, lse, delta, dbase, dbase + 2 * E, dbase + 4 * E, B, H, N,
               dh, qs, os_, scale)
    return dqkv


class AttentionFn(torch.autograd.Function):
    """timm Attention.forward as one autograd node: qkv GEMM -> flash attention core -> proj GEMM."""

    @staticmethod
    def forward(ctx, x, qkv_w, qkv_b, proj_w, proj_b, num_heads, scale):
        B, N, C = x.shape
        dh = C // num_heads
        x16 = L.cast_bf16(x.reshape(B * N, C).contiguous())
        qkv = L.gemm(x16, shadow(qkv_w), bias=qkv_b)
        o16, lse = _attn_core_fwd(qkv, B, N, num_heads, dh, scale)
        y = L.gemm(o16.view(B * N, C), shadow(proj_w), bias=proj_b, out_dtype=torch.float32)
        ctx.save_for_backward(x16, qkv, o16, lse, qkv_w, proj_w)
        ctx.meta = (B, N, C, num_heads, dh, scale, qkv_b is not None, proj_b is not None)
        return y.view(B, N, C)

    @staticmethod
    def backward(ctx, dy):
        x16, qkv, o16, lse, qkv_w, proj_w = ctx.saved_tensors
        B, N, C, H, dh, scale, has_qb, has_pb = ctx.meta
        dy16 = L.cast_bf16(dy.reshape(B * N, C).contiguous())
        dproj_w = L.gemm(dy16, o16.view(B * N, C), a_mn=True, b_mn=True, out_dtype=torch.float32)
        dproj_b = L.colsum(dy16) if has_pb else None
        do16 = L.gemm(dy16, shadow(proj_w), b_mn=True)
        dqkv = _attn_core_bwd(qkv, o16, do16.view(B, N, C), lse, B, N, H, dh, scale)
        dqkv_w = L.gemm(dqkv, x16, a_mn=True, b_mn=True, out_dtype=torch.float32)
        dqkv_b = L.colsum(dqkv) if has_qb else None
        dx = L.gemm(dqkv, shadow(qkv_w), b_mn=True, out_dtype=torch.float32).view(B, N, C)
        return dx, dqkv_w, dqkv_b, dproj_w, dproj_b, None, None


# ----------------------------------------------------------------------------------------------------------------
# Whole encoder block (timm Block.forward, pre-norm): 7 kernels forward, 17 backward, one autograd node
# ----------------------------------------------------------------------------------------------------------------
class BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b, num_heads, scale,
                eps1, eps2):
        B, N, D = x.shape
        T = B * N
        dh = D // num_heads
        x = x.contiguous()
        x2 = x.view(T, D)
        h16, _, _, mean1, rstd1 = L.layernorm_fwd(x2, n1w, n1b, eps1)
        qkv = L.gemm(h16, shadow(qkv_w), bias=qkv_b)
        o16, lse = _attn_core_fwd(qkv, B, N, num_heads, dh, scale)
        x1 = L.gemm(o16.view(T, D), shadow(proj_w), bias=proj_b, residual=x2, out_dtype=torch.float32)
        g16, _, _, mean2, rstd2 = L.layernorm_fwd(x1, n2w, n2b, eps2)
        w1 = shadow(fc1_w)
        pre = torch.empty((T, w1.shape[0]), device=x.device, dtype=torch.bfloat16)
        a16 = L.gemm(g16, w1, bias=fc1_b, epilogue=L.EPI_GELU, aux_out=pre)
        y = L.gemm(a16, shadow(fc2_w), bias=fc2_b, residual=x1, out_dtype=torch.float32)
        ctx.save_for_backward(x2, mean1, rstd1, h16, qkv, o16, lse, x1, mean2, rstd2, g16, pre, a16, n1w, qkv_w, proj_w,
                              n2w, fc1_w, fc2_w)
        ctx.meta = (B, N, D, num_heads, dh, scale, qkv_b is not None, proj_b is not None, fc1_b is not None,
                    fc2_b is not None)
        ctx.refs = (n1b, qkv_b, proj_b, n2b, fc1_b, fc2_b)  # bias-type parameters (gradient sinks are looked up on them)
        return y.view(B, N, D)

    @staticmethod
    def backward(ctx, dy):
        (x2, mean1, rstd1, h16, qkv, o16, lse, x1, mean2, rstd2, g16, pre, a16, n1w, qkv_w, proj_w, n2w, fc1_w,
         fc2_w) = ctx.saved_tensors
        B, N, D, H, dh, scale, has_qb, has_pb, has_b1, has_b2 = ctx.meta
        T = B * N
        dy2 = dy.reshape(T, D).contiguous()
        # the downstream block's backward leaves a bf16 copy of this gradient on the tensor object (see below); autograd
        # hands the same object through when it did not have to accumulate, which saves one cast pass per block
        # The side products are only valid for the very tensor they were computed from: same storage AND same version
        # (autograd's input buffer may add another consumer's gradient IN PLACE into the tensor it was handed first,
        # which keeps the Python object and its data pointer but bumps the version counter).
        fresh = (getattr(dy, "_s3d_bf16_src", None) == dy.data_ptr() and getattr(dy, "_s3d_ver", None) == dy._version)
        dy16 = getattr(dy, "_s3d_bf16", None) if fresh else None
        if dy16 is None or dy16.shape != (T, D):
            dy16 = L.cast_bf16(dy2)
        n1b, qkv_b, proj_b, n2b, fc1_b, fc2_b = ctx.refs
        pg = _ParamGradStream(dy.device, T * D <= _OVERLAP_LIMIT)  # weight / bias gradients off the critical chain
        # MLP
        # bias gradient of fc2 = column sums of dy: the LayerNorm backward that PRODUCED dy (the downstream block's norm1)
        # has accumulated them on the way (see the end of this function); otherwise one pass over dy
        dy_colsum = getattr(dy, "_s3d_colsum", None) if fresh else None
        with pg.fork():
            dfc2_w = _wgrad(fc2_w, dy16, a16)
            if fc2_b is None:
                dfc2_b = None
            elif dy_colsum is None:
                dfc2_b = _bgrad(fc2_b, dy16)
            elif _sink(fc2_b) is not None:
                _sink(fc2_b).add_(dy_colsum)
                fc2_b._s3d_owner.note_write(fc2_b)
                dfc2_b = None
            else:
                dfc2_b = dy_colsum
        dpre = L.gemm(dy16, shadow(fc2_w), b_mn=True, epilogue=L.EPI_DGELU, aux_in=pre)
        with pg.fork():
            dfc1_w = _wgrad(fc1_w, dpre, g16)
            dfc1_b = _bgrad(fc1_b, dpre)
        dg = L.gemm(dpre, shadow(fc1_w), b_mn=True)
        # dx1 is the output gradient of attn.proj: its column sums (= the proj bias gradient) come out of this kernel
        if proj_b is not None:
            dx1, dx1_16, dn2w, dn2b, dproj_b = _ln_bwd_colsum(dg, x1, n2w, n2b, mean2, rstd2, proj_b, dres=dy2, want_bf16=True)
        else:
            dx1, dx1_16, dn2w, dn2b = _ln_bwd(dg, x1, n2w, n2b, mean2, rstd2, dres=dy2, want_bf16=True)
            dproj_b = None
        # attention
        with pg.fork():
            dproj_w = _wgrad(proj_w, dx1_16, o16.view(T, D))
        do16 = L.gemm(dx1_16, shadow(proj_w), b_mn=True)
        dqkv = _attn_core_bwd(qkv, o16, do16.view(B, N, D), lse, B, N, H, dh, scale)
        with pg.fork():
            dqkv_w = _wgrad(qkv_w, dqkv, h16)
            dqkv_b = _bgrad(qkv_b, dqkv)
        dh_ = L.gemm(dqkv, shadow(qkv_w), b_mn=True)
        # dx is the output gradient of the upstream block's mlp.fc2: leave its column sums on the tensor for that block
        dx, dx16, dn1w, dn1b, dx_colsum = _ln_bwd_colsum(dh_, x2, n1w, n1b, mean1, rstd1, None, dres=dx1, want_bf16=True)
        pg.join()
        dx = dx.view(B, N, D)
        dx._s3d_bf16 = dx16
        dx._s3d_bf16_src = dx.data_ptr()
        dx._s3d_colsum = dx_colsum
        dx._s3d_ver = dx._version
        return (dx, dn1w, dn1b, dqkv_w, dqkv_b, dproj_w, dproj_b, dn2w, dn2b, dfc1_w, dfc1_b, dfc2_w,
                dfc2_b, None, None, None, None)


# ----------------------------------------------------------------------------------------------------------------
# Thin fp32 layers (classification heads, the point models' stem): CUDA-core fp32 GEMM, no bf16 rounding
# ----------------------------------------------------------------------------------------------------------------
_ONES = {}


def _ones_row(n, device):
    key = (device.type, device.index)
    t = _ONES.get(key)
    if t is None or t.numel() < n:
        t = torch.ones(max(n, 1 << 18), device=device, dtype=torch.float32)
        _ONES[key] = t
    return t[:n].view(1, n)


def _wgrad_f32(w, dy2, x2):
    """dW (+)= dy2^T @ x2 in fp32 (into the parameter's gradient sink when it has one)."""
    sk = _sink(w)
    if sk is None:
        return L.sgemm(dy2.t(), x2).view(w.shape)
    L.sgemm(dy2.t(), x2, out=sk.view(sk.shape[0], -1), accumulate=True)
    w._s3d_owner.note_write(w)
    return None


def _bgrad_f32(b, dy2):
    """db (+)= column sums of dy2 (fp32) as a [1, R] x [R, N] product of the same kernel."""
    if b is None:
        return None
    ones = _ones_row(dy2.shape[0], dy2.device)
    sk = _sink(b)
    if sk is None:
        return L.sgemm(ones, dy2).view(b.shape)
    L.sgemm(ones, dy2, out=sk.view(1, -1), accumulate=True)
    b._s3d_owner.note_write(b)
    return None


class LinearF32Fn(torch.autograd.Function):
    """y = x W^T + b in fp32 (nn.Linear heads: vit_3d_2d_pretrain.py:366, models/3DViT/model.py:232)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        K = x.shape[-1]
        x2 = x.reshape(-1, K).float()
        y = L.sgemm(x2, weight.detach().t(), bias=bias.detach() if bias is not None else None)
        ctx.save_for_backward(x2, weight)
        ctx.bias_ref = bias
        ctx.in_shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).float()
        dx = L.sgemm(dy2, weight.detach()).view(ctx.in_shape) if ctx.needs_input_grad[0] else None
        dw = _wgrad_f32(weight, dy2, x2) if ctx.needs_input_grad[1] else None
        db = _bgrad_f32(ctx.bias_ref, dy2) if (ctx.bias_ref is not None and ctx.needs_input_grad[2]) else None
        return dx, dw, db


class PointStemFn(torch.autograd.Function):
    """f = fc1(x) + fc_pos_embed(xyz), each Linear -> ReLU -> Linear (models/3DViT/model.py:236-247, 310-311) as one
    node: four fp32 GEMM launches forward (the second branch accumulates onto the first), the ReLU masks applied inside
    the backward GEMMs. The inputs are data (no input gradient)."""

    @staticmethod
    def forward(ctx, x, xyz, w1a, b1a, w1b, b1b, w2a, b2a, w2b, b2b):
        B, N, _ = x.shape
        x2 = x.reshape(B * N, -1).float()
        p2 = xyz.reshape(B * N, -1).float()
        h1 = L.sgemm(x2, w1a.detach().t(), bias=b1a.detach(), relu=True)
        h2 = L.sgemm(p2, w2a.detach().t(), bias=b2a.detach(), relu=True)
        f = L.sgemm(h1, w1b.detach().t(), bias=b1b.detach())
        L.sgemm(h2, w2b.detach().t(), bias=b2b.detach(), out=f, accumulate=True)
        ctx.save_for_backward(x2, p2, h1, h2, w1a, w1b, w2a, w2b)
        ctx.refs = (b1a, b1b, b2a, b2b)
        return f.view(B, N, -1)

    @staticmethod
    def backward(ctx, df):
        x2, p2, h1, h2, w1a, w1b, w2a, w2b = ctx.saved_tensors
        b1a, b1b, b2a, b2b = ctx.refs
        df2 = df.reshape(h1.shape[0], -1).float()
        if not df2.is_contiguous():
            df2 = df2.contiguous()
        out = []
        for inp, h, wa, ba, wb, bb in ((x2, h1, w1a, b1a, w1b, b1b), (p2, h2, w2a, b2a, w2b, b2b)):
            dwb = _wgrad_f32(wb, df2, h)
            dbb = _bgrad_f32(bb, df2)
            dh = L.sgemm(df2, wb.detach(), gate=h)  # gradient through the ReLU: zero where the activation was clipped
            dwa = _wgrad_f32(wa, dh, inp)
            dba = _bgrad_f32(ba, dh)
            out += [dwa, dba, dwb, dbb]
        return (None, None, *out)


# ----------------------------------------------------------------------------------------------------------------
# LayerNorm as its own node (VisionTransformer.norm on the fp32 residual stream)
# ----------------------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous()
        _, y, _, mean, rstd = L.layernorm_fwd(x2, weight, bias, eps, want_bf16=False, want_f32=True)
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.bias_ref = bias
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        dy2 = dy.reshape(x2.shape).contiguous()
        dx, _, dg, db = _ln_bwd(dy2, x2, weight, ctx.bias_ref, mean, rstd)
        return dx.view(dy.shape), dg, db, None


# ----------------------------------------------------------------------------------------------------------------
# Voxel patchify: Conv3d(1 -> D, k = s = cell) (+ mean over z) as gather + tensor-core GEMM, token-major output
# ----------------------------------------------------------------------------------------------------------------
class VoxelPatchifyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, cell, patch, zmean):
        D = weight.shape[0]
        K = cell ** 3
        kpad = (K + 63) // 64 * 64
        if x.dtype not in (torch.float32, torch.uint8, torch.bool, torch.int32):
            x = x.float()
        P = L.voxel_patch_gather(x.contiguous(), cell, patch, kpad, zmean)  # uint8 / int32 occupancy is read as is
        w16 = _padded_conv_weight(weight, K, kpad)
        alpha = 1.0 / patch if zmean else 1.0
        tok = L.gemm(P, w16, bias=bias, alpha=alpha, out_dtype=torch.float32)  # [B*p*p(*p), D]
        ctx.save_for_backward(P, weight)
        ctx.meta = (K, kpad, alpha, bias is not None)
        return tok.view(x.shape[0], -1, D)

    @staticmethod
    def backward(ctx, dtok):
        P, weight = ctx.saved_tensors
        K, kpad, alpha, has_bias = ctx.meta
        D = weight.shape[0]
        d16 = L.cast_bf16(dtok.reshape(-1, D).contiguous())
        dw = L.gemm(d16, P, a_mn=True, b_mn=True, alpha=alpha, out_dtype=torch.float32)  # [D, kpad]
        dw = dw[:, :K].reshape(weight.shape)
        db = L.colsum(d16) if has_bias else None
        return None, dw, db, None, None, None


def _padded_conv_weight(weight, K, kpad):
    """bf16 [D, kpad] view of the Conv3d weight [D,1,c,c,c], zero padded along K (cached per parameter version)."""
    key = (weight._version, weight.data_ptr(), kpad)
    cache = getattr(weight, "_s3d_pad_cache", None)
    s = getattr(weight, "_s3d_shadow", None)
    if s is None and cache is not None and cache[0] == key:
        return cache[1]
    D = weight.shape[0]
    w16 = torch.zeros((D, kpad), device=weight.device, dtype=torch.bfloat16)
    src = s if s is not None else L.cast_bf16(weight.detach().contiguous())
    w16[:, :K] = src.reshape(D, K)
    if s is None:
        try:
            weight._s3d_pad_cache = (key, w16)
        except Exception:
            pass
    return w16


# ----------------------------------------------------------------------------------------------------------------
# group_embed: nn.TransformerEncoderLayer (post-norm, ReLU, sequence-first) with a flash attention core
# ----------------------------------------------------------------------------------------------------------------
class GroupEmbedFn(torch.autograd.Function):
    """nn.TransformerEncoderLayer (post-norm, ReLU, sequence-first) as one autograd node.

    Dropout (p = 0.1 by default in the reference, vit_3d_2d_pretrain.py:381) is active when `drop_p > 0` and a device
    seed is given (the module does that in train()): site 1 = attention probabilities (inside the flash kernels),
    site 2 = dropout1 on the attention block output, site 3 = dropout inside the FFN (after ReLU), site 4 = dropout2 on
    the FFN output. Masks are counter-based hashes of (seed, site, row, column), regenerated in backward."""

    @staticmethod
    def forward(ctx, x, in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1w, n1b, n2w, n2b, nhead, eps, drop_p=0.0,
                drop_seed=None):
        S, Nb, E = x.shape
        T = S * Nb
        dh = E // nhead
        scale = dh ** -0.5
        drop = drop_seed is not None and drop_p > 0.0
        seed = drop_seed.clone() if drop else None  # backward must see the value this forward used
        x2 = x.contiguous().view(T, E)
        x16 = L.cast_bf16(x2)
        qkv = L.gemm(x16, shadow(in_w), bias=in_b)  # [S*Nb, 3E], sequence-first
        o16 = torch.empty((T, E), device=x.device, dtype=torch.bfloat16)
        lse = torch.empty((Nb, nhead, S), device=x.device, dtype=torch.float32)
        qs = (3 * E, dh, Nb * 3 * E)
        os_ = (E, dh, Nb * E)
        base = qkv.data_ptr()
        L.attn_fwd(base, base + 2 * E, base + 4 * E, o16, lse, Nb, nhead, S, dh, qs, os_, scale, drop_seed=seed,
                   drop_site=1, drop_p=drop_p if drop else 0.0)
        if drop:
            attn_out = L.gemm(o16, shadow(out_w), bias=out_b, out_dtype=torch.float32)
            sa = L.dropout_add(attn_out, x2, seed, 2, drop_p)  # x + dropout1(attn(x))
        else:
            sa = L.gemm(o16, shadow(out_w), bias=out_b, residual=x2, out_dtype=torch.float32)  # x + attn(x)
        y1_16, y1, _, mean1, rstd1 = L.layernorm_fwd(sa, n1w, n1b, eps, want_f32=True)
        h16 = L.gemm(y1_16, shadow(l1_w), bias=l1_b, epilogue=L.EPI_RELU)
        if drop:
            L.dropout_bf16(h16, seed, 3, drop_p, inplace=True)
            ffn = L.gemm(h16, shadow(l2_w), bias=l2_b, out_dtype=torch.float32)
            f = L.dropout_add(ffn, y1, seed, 4, drop_p)
        else:
            f = L.gemm(h16, shadow(l2_w), bias=l2_b, residual=y1, out_dtype=torch.float32)
        _, y2, _, mean2, rstd2 = L.layernorm_fwd(f, n2w, n2b, eps, want_bf16=False, want_f32=True)
        ctx.save_for_backward(x16, qkv, o16, lse, sa, mean1, rstd1, y1_16, h16, f, mean2, rstd2, in_w, out_w, l1_w, l2_w,
                              n1w, n2w, seed)
        ctx.meta = (S, Nb, E, nhead, dh, scale, qs, os_, drop_p if drop else 0.0)
        ctx.refs = (in_b, out_b, l1_b, l2_b, n1b, n2b)
        return y2.view(S, Nb, E)

    @staticmethod
    def backward(ctx, dy):
        (x16, qkv, o16, lse, sa, mean1, rstd1, y1_16, h16, f, mean2, rstd2, in_w, out_w, l1_w, l2_w, n1w,
         n2w, seed) = ctx.saved_tensors
        S, Nb, E, nhead, dh, scale, qs, os_, drop_p = ctx.meta
        drop = drop_p > 0.0
        keep_scale = 1.0
        if drop:
            keep_scale = 1.0 / (1.0 - int(drop_p * 16384.0 + 0.5) / 16384.0)  # csrc/common.cuh::drop_keep_scale
        T = S * Nb
        dy2 = dy.reshape(T, E).contiguous()
        in_b, out_b, l1_b, l2_b, n1b, n2b = ctx.refs
        df, df16, dn2w, dn2b = _ln_bwd(dy2, f, n2w, n2b, mean2, rstd2, want_bf16=True)
        dffn16 = L.dropout_bf16(df16, seed, 4, drop_p) if drop else df16  # gradient of the (pre-dropout2) FFN output
        dl2_w = _wgrad(l2_w, dffn16, h16)
        dl2_b = _bgrad(l2_b, dffn16)
        # h16 holds the dropped, rescaled ReLU output: its zeros mask both the ReLU and the dropout; kept entries carry 1/(1-p)
        dh16 = L.gemm(dffn16, shadow(l2_w), b_mn=True, epilogue=L.EPI_DRELU, aux_in=h16, alpha=keep_scale)
        dl1_w = _wgrad(l1_w, dh16, y1_16)
        dl1_b = _bgrad(l1_b, dh16)
        dy1 = L.gemm(dh16, shadow(l1_w), b_mn=True, residual=df, out_dtype=torch.float32)  # + residual branch of y1
        dsa, dsa16, dn1w, dn1b = _ln_bwd(dy1, sa, n1w, n1b, mean1, rstd1, want_bf16=True)
        dattn16 = L.dropout_bf16(dsa16, seed, 2, drop_p) if drop else dsa16
        dout_w = _wgrad(out_w, dattn16, o16)
        dout_b = _bgrad(out_b, dattn16)
        do16 = L.gemm(dattn16, shadow(out_w), b_mn=True)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        base, dbase = qkv.data_ptr(), dqkv.data_ptr()
        L.attn_bwd(base, base + 2 * E, base + 4 * E, o16, do16, lse, delta, dbase, dbase + 2 * E, dbase + 4 * E, Nb, nhead,
                   S, dh, qs, os_, scale, drop_seed=seed if drop else None, drop_site=1, drop_p=drop_p)
        din_w = _wgrad(in_w, dqkv, x16)
        din_b = _bgrad(in_b, dqkv)
        dx = L.gemm(dqkv, shadow(in_w), b_mn=True, residual=dsa, out_dtype=torch.float32)
        return (dx.view(S, Nb, E), din_w, din_b, dout_w, dout_b, dl1_w, dl1_b, dl2_w, dl2_b, dn1w, dn1b, dn2w, dn2b, None,
                None, None, None)


# ----------------------------------------------------------------------------------------------------------------
# index_points with a scatter-add backward
# ----------------------------------------------------------------------------------------------------------------
class GatherRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        pts = points.contiguous().float()
        flat = idx.reshape(idx.shape[0], -1).contiguous()
        out = L.gather_rows(pts, flat)
        ctx.save_for_backward(flat)
        ctx.n = pts.shape[1]
        return out.view(*idx.shape, pts.shape[-1])

    @staticmethod
    def backward(ctx, dout):
        (flat,) = ctx.saved_tensors
        g = dout.reshape(flat.shape[0], flat.shape[1], -1).contiguous().float()
        return L.scatter_add_rows(g, flat, ctx.n), None


# ----------------------------------------------------------------------------------------------------------------
# BatchNorm helpers shared by the point-tokenizer layers
# ----------------------------------------------------------------------------------------------------------------
def _bn_eval_affine(gamma, beta, running_mean, running_var, eps):
    """eval(): BatchNorm is the affine map of its running statistics (a handful of [C]-sized ops)."""
    rstd = torch.rsqrt(running_var.float() + eps)
    scale = (gamma.detach() * rstd).contiguous()
    shift = (beta.detach() - running_mean * scale).contiguous()
    return running_mean.float().contiguous(), rstd.contiguous(), scale, shift


# ----------------------------------------------------------------------------------------------------------------
# Set abstraction: sample_and_group + [Conv2d 1x1 -> BatchNorm2d -> ReLU] x 2 + max over the K neighbours
# (data/pointnet_util.py:99-138, 220-244) without ever building the grouped tensor; see csrc/pointnet_fused.cu
# ----------------------------------------------------------------------------------------------------------------
class SetAbstractionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, xyz, cxyz, idx, w1, b1, g1, be1, rm1, rv1, w2, b2, g2, be2, rm2, rv2, training, eps1, mom1,
                eps2, mom2):
        B, N, Cf = f.shape
        S, K = idx.shape[1], idx.shape[2]
        C1, C2 = w1.shape[0], w2.shape[0]
        Cin = w1.numel() // C1
        assert Cin == Cf + 3 and w2.numel() == C2 * C1
        R, G = B * S * K, B * S
        w1_2d = w1.detach().reshape(C1, Cin)
        # per-point part of layer 1, [B*N, C1]: feature columns on the tensor cores with hi/lo-split bf16 operands
        # (fp32-grade: a BatchNorm follows, see s3d_split_bf16x3); the xyz columns stay fp32 in the gather passes
        f3 = L.split_bf16x3(f.detach().reshape(B * N, Cf).contiguous().float())
        w1f3 = L.split_bf16x3(w1_2d[:, 3:], weight_layout=True)
        uf = L.gemm(f3, w1f3, bias=b1, out_dtype=torch.float32)
        grp = L.SaGroup(uf, xyz.contiguous(), cxyz.contiguous(), idx.contiguous(), w1_2d)
        if training:
            mean1, rstd1, sc1, sh1 = L.bn_finalize_fwd(L.sa_group_fwd_stats(grp), R, g1, be1, eps1, mom1, rm1, rv1)
        else:
            mean1, rstd1, sc1, sh1 = _bn_eval_affine(g1, be1, rm1, rv1, eps1)
        a1 = L.sa_group_fwd_act(grp, sc1, sh1)  # bf16 [R, C1]
        w2_16 = shadow(w2).reshape(C2, C1)
        z2 = L.gemm(a1, w2_16, bias=b2, out_dtype=torch.float32)  # [R, C2]
        zmax, zmin, kmax, kmin, part2 = L.sa_group_reduce(z2, G, K)
        if training:
            mean2, rstd2, sc2, sh2 = L.bn_finalize_fwd(part2, R, g2, be2, eps2, mom2, rm2, rv2)
        else:
            mean2, rstd2, sc2, sh2 = _bn_eval_affine(g2, be2, rm2, rv2, eps2)
        out, zsel, ksel = L.sa_pool_select(zmax, zmin, kmax, kmin, sc2, sh2)
        ctx.save_for_backward(uf, grp.t[1], grp.t[2], grp.t[3], f3, w1f3, a1, z2, zsel, ksel, w1, w2, mean1, rstd1,
                              sc1, mean2, rstd2, sc2, sh2)
        ctx.meta = (B, N, S, K, Cf, C1, C2, training)
        ctx.refs = (b1, b2)
        return out.view(B, S, C2)

    @staticmethod
    def backward(ctx, dout):
        (uf, xyz, cxyz, idx, f3, w1f3, a1, z2, zsel, ksel, w1, w2, mean1, rstd1, sc1, mean2, rstd2, sc2,
         sh2) = ctx.saved_tensors
        B, N, S, K, Cf, C1, C2, training = ctx.meta
        f16, w1f16 = f3[:, :Cf], w1f3[:, :Cf]  # the hi parts (row stride 3*Cf)
        b1, b2 = ctx.refs
        R, G = B * S * K, B * S
        Cin = Cf + 3
        dout2 = dout.reshape(G, C2).contiguous().float()
        # layer 2 (pooled): BatchNorm backward statistics only involve the pooled rows
        m1, m2, dg2, dbe2 = L.bn_finalize_bwd(L.bn_rows_bwd_stats(dout2, zsel, sc2, sh2, mean2, rstd2), R, training)
        dz2 = L.sa_dz2_expand(z2, dout2, zsel, ksel, sc2, sh2, mean2, rstd2, m1, m2, G, K)  # bf16 [R, C2]
        dw2 = _wgrad(w2, dz2, a1)
        db2 = _bgrad(b2, dz2)
        da1 = L.gemm(dz2, shadow(w2).reshape(C2, C1), b_mn=True)  # bf16 [R, C1]
        # layer 1
        grp = L.SaGroup(uf, xyz, cxyz, idx, w1.detach().reshape(C1, Cin))
        n1, n2, dg1, dbe1 = L.bn_finalize_bwd(L.sa_group_bwd_stats(grp, mean1, rstd1, a1, da1), R, training)
        duf, dwx_part = L.sa_group_bwd_scatter(grp, sc1, mean1, rstd1, n1, n2, a1, da1)
        duf16 = L.cast_bf16(duf)
        df = None
        if ctx.needs_input_grad[0]:
            df = L.gemm(duf16, w1f16, b_mn=True, out_dtype=torch.float32).view(B, N, Cf)
        dw1f = L.gemm(duf16, f16, a_mn=True, b_mn=True, out_dtype=torch.float32)  # [C1, Cf]
        db1 = L.colsum(duf16) if b1 is not None else None
        dw1 = torch.cat([dwx_part.sum(0).t(), dw1f], dim=1).view(w1.shape)
        return (df, None, None, None, dw1, db1, dg1, dbe1, None, None, dw2, db2, dg2, dbe2, None, None, None, None,
                None, None, None)


# ----------------------------------------------------------------------------------------------------------------
# Linear -> BatchNorm1d -> ReLU on [rows, C] (TransitionUp.fc1 / fc2, models/3DViT/model.py:47-60)
# ----------------------------------------------------------------------------------------------------------------
class LinearBnReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, g, be, rm, rv, training, eps, mom):
        K = x.shape[-1]
        x3 = L.split_bf16x3(x.detach().reshape(-1, K).contiguous().float())  # hi/lo split: a BatchNorm follows
        z = L.gemm(x3, L.split_bf16x3(w.detach(), weight_layout=True), bias=b, out_dtype=torch.float32)
        R = z.shape[0]
        if training:
            mean, rstd, sc, sh = L.bn_finalize_fwd(L.bn_rows_stats(z), R, g, be, eps, mom, rm, rv)
        else:
            mean, rstd, sc, sh = _bn_eval_affine(g, be, rm, rv, eps)
        y, _ = L.bn_relu_apply(z, sc, sh)
        ctx.save_for_backward(x3, z, w, mean, rstd, sc, sh)
        ctx.meta = (x.shape, training)
        ctx.bias_ref = b
        return y.view(*x.shape[:-1], z.shape[1])

    @staticmethod
    def backward(ctx, dy):
        x3, z, w, mean, rstd, sc, sh = ctx.saved_tensors
        in_shape, training = ctx.meta
        x16 = x3[:, :in_shape[-1]]
        R = z.shape[0]
        dy2 = dy.reshape(z.shape).contiguous().float()
        m1, m2, dg, dbe = L.bn_finalize_bwd(L.bn_rows_bwd_stats(dy2, z, sc, sh, mean, rstd), R, training)
        dz16 = L.bn_relu_bwd_apply(dy2, z, sc, sh, mean, rstd, m1, m2)
        dw = _wgrad(w, dz16, x16)
        db = _bgrad(ctx.bias_ref, dz16)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = L.gemm(dz16, _w2d(shadow(w)), b_mn=True, out_dtype=torch.float32).view(in_shape)
        return dx, dw, db, dg, dbe, None, None, None, None, None


# ----------------------------------------------------------------------------------------------------------------
# 3-NN inverse-distance interpolation (pointnet_util.py:401-408) + the residual add of TransitionUp
# ----------------------------------------------------------------------------------------------------------------
class ThreeNNInterpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, idx, dist, addend):
        feats = feats.contiguous().float()
        out = L.three_nn_interp_fwd(feats, idx, dist, None if addend is None else addend.contiguous().float())
        ctx.save_for_backward(idx, dist)
        ctx.S = feats.shape[1]
        ctx.has_addend = addend is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        idx, dist = ctx.saved_tensors
        dout = dout.contiguous().float()
        dfeats = L.three_nn_interp_bwd(dout, idx, dist, ctx.S) if ctx.needs_input_grad[0] else None
        return dfeats, None, None, (dout if ctx.has_addend else None)
