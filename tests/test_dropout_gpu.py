"""GPU: dropout of the group_embed layer (nn.TransformerEncoderLayer(d_model, nhead=4, dim_feedforward=d_model), default
p = 0.1, reference vit_3d_2d_pretrain.py:381, active in train()).

torch's dropout masks cannot be reproduced bit for bit by any other implementation (they depend on its Philox stream and
kernel launch geometry), so parity is checked the other way round: the product's counter-based mask is recomputed on
the host from (seed, site, row, column) -- tests restate the hash of csrc/common.cuh in numpy -- and fed to a plain fp32
torch restatement of the layer; outputs and gradients must then agree within the bf16 tolerance (2e-2 / 5e-2 relative
to the tensor norm), and the mask statistics must match p."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

M32 = np.uint64(0xFFFFFFFF)


def _mix(x):
    x = (x * np.uint64(0x2C1B3C6D)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x297A2D39)) & M32
    x ^= x >> np.uint64(15)
    return x


def _site_seed(seed, site):
    return _mix(np.uint64((int(seed) ^ ((site * 0x632BE5AB) & 0xFFFFFFFF) ^ 0xA511E9B3) & 0xFFFFFFFF))


def _fold(x, m):
    w = x * np.uint64(m)  # x, m < 2^32: the product fits uint64
    return (w & M32) ^ (w >> np.uint64(32))


def keep_mask(seed, site, rows, cols, p):
    """bool [len(rows), len(cols)]: the mask of csrc/common.cuh::drop_keep (row key -> two multiply-fold rounds per
    column pair -> 14-bit draw per element, kept when >= round(p * 16384))."""
    thresh = int(p * 16384.0 + 0.5)
    rows = np.asarray(rows, dtype=np.uint64)[:, None]
    cols = np.asarray(cols, dtype=np.uint64)[None, :]
    rowkey = _mix(_site_seed(seed, site) ^ ((rows * np.uint64(0x9E3779B1)) & M32))
    y = (rowkey + (cols >> np.uint64(1)) * np.uint64(0x85EBCA77)) & M32
    z = _fold(_fold(y, 0xD6E8FEB9), 0xCA6B1B35) & np.uint64(0x3FFF3FFF)
    bits = np.where((cols & np.uint64(1)) != 0, z >> np.uint64(16), z & np.uint64(0xFFFF))
    return bits >= thresh


def keep_scale(p):
    return 1.0 / (1.0 - int(p * 16384.0 + 0.5) / 16384.0)


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


def test_elementwise_dropout_kernels_match_host_mask():
    from simple3d_former_b200 import _lib as L
    dev = _dev()
    g = torch.Generator().manual_seed(0)
    rows, cols, p = 777, 256, 0.1
    seed = torch.tensor([123456789], dtype=torch.int32, device=dev)
    x = torch.randn(rows, cols, generator=g)
    res = torch.randn(rows, cols, generator=g)
    m = torch.from_numpy(keep_mask(123456789, 4, np.arange(rows), np.arange(cols), p))
    got = L.dropout_add(x.to(dev), res.to(dev), seed, 4, p).cpu()
    want = res + torch.where(m, x * keep_scale(p), torch.zeros(()))
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
    m3 = torch.from_numpy(keep_mask(123456789, 3, np.arange(rows), np.arange(cols), p))
    xb = x.bfloat16()
    got16 = L.dropout_bf16(xb.to(dev), seed, 3, p).cpu()
    want16 = torch.where(m3, xb.float() * keep_scale(p), torch.zeros(())).bfloat16()
    assert torch.equal(got16, want16)
    assert abs(float(m.float().mean()) - 0.9) < 5e-3 and abs(float(m3.float().mean()) - 0.9) < 5e-3
    assert float((m != m3).float().mean()) > 0.1  # sites draw independent masks
    # a different seed gives a different mask
    seed2 = torch.tensor([123456790], dtype=torch.int32, device=dev)
    assert not torch.equal(L.dropout_bf16(xb.to(dev), seed2, 3, p).cpu(), got16)


@pytest.mark.parametrize("shape", [(2, 2, 200, 64), (1, 4, 700, 64), (3, 4, 333, 192)])
def test_attention_probability_dropout_matches_masked_reference(shape):
    """tcgen05 flash kernels, timm layout [B, N, 3, H, dh]; mask row = (b*H + h)*N + query, column = key."""
    from simple3d_former_b200 import _lib as L
    B, H, N, dh = shape
    dev = _dev()
    E = H * dh
    p, seed_val = 0.1, 424242
    g = torch.Generator().manual_seed(1)
    qkv = (torch.randn(B, N, 3, H, dh, generator=g) * 0.7).bfloat16().to(dev)
    dout = torch.randn(B, N, E, generator=g).bfloat16().to(dev)
    seed = torch.tensor([seed_val], dtype=torch.int32, device=dev)
    scale = dh ** -0.5
    out = torch.empty(B, N, E, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev, dtype=torch.float32)
    qs, os_ = (N * 3 * E, dh, 3 * E), (N * E, dh, E)
    base = qkv.data_ptr()
    L.attn_fwd(base, base + 2 * E, base + 4 * E, out, lse, B, H, N, dh, qs, os_, scale, drop_seed=seed, drop_site=1, drop_p=p)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    db = dqkv.data_ptr()
    L.attn_bwd(base, base + 2 * E, base + 4 * E, out, dout, lse, delta, db, db + 2 * E, db + 4 * E, B, H, N, dh, qs, os_, scale,
               drop_seed=seed, drop_site=1, drop_p=p)
    torch.cuda.synchronize()
    # fp32 reference with the same mask
    q, k, v = (qkv[:, :, i].float().permute(0, 2, 1, 3).clone().requires_grad_(True) for i in range(3))  # [B,H,N,dh]
    mask = torch.from_numpy(keep_mask(seed_val, 1, np.arange(B * H * N), np.arange(N), p)).view(B, H, N, N).to(dev)
    P = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1)
    O = (P * mask * keep_scale(p)) @ v
    want = O.permute(0, 2, 1, 3).reshape(B, N, E)
    want.backward(dout.float())
    assert _rel(out, want) < 2e-2, _rel(out, want)
    assert torch.allclose(lse, torch.logsumexp((q @ k.transpose(-1, -2)) * scale, dim=-1), atol=2e-2, rtol=1e-3)
    for i, t in enumerate((q, k, v)):
        got = dqkv[:, :, i].float().permute(0, 2, 1, 3)
        assert _rel(got, t.grad) < 5e-2, (i, _rel(got, t.grad))
    # with dropout off the same call reproduces plain attention (and differs from the dropped output)
    out0 = torch.empty_like(out)
    L.attn_fwd(base, base + 2 * E, base + 4 * E, out0, lse, B, H, N, dh, qs, os_, scale)
    assert _rel(out0, (P @ v).permute(0, 2, 1, 3).reshape(B, N, E)) < 2e-2
    assert _rel(out0, out) > 0.1


def _layer_reference(layer, x, seed_val, p):
    """nn.TransformerEncoderLayer.forward (post-norm, ReLU, sequence-first) in fp32 torch ops with the product's masks."""
    S, Nb, E = x.shape
    H = layer.nhead
    dh = E // H
    a = layer.self_attn
    dev = x.device

    def mask2d(site):
        return torch.from_numpy(keep_mask(seed_val, site, np.arange(S * Nb), np.arange(E), p)).view(S, Nb, E).to(dev)

    qkv = F.linear(x, a.in_proj_weight, a.in_proj_bias)  # [S, Nb, 3E]
    q, k, v = (qkv[..., i * E:(i + 1) * E].reshape(S, Nb, H, dh).permute(1, 2, 0, 3) for i in range(3))  # [Nb,H,S,dh]
    P = torch.softmax((q @ k.transpose(-1, -2)) * dh ** -0.5, dim=-1)
    m1 = torch.from_numpy(keep_mask(seed_val, 1, np.arange(Nb * H * S), np.arange(S), p)).view(Nb, H, S, S).to(dev)
    o = ((P * m1 * keep_scale(p)) @ v).permute(2, 0, 1, 3).reshape(S, Nb, E)
    sa = x + F.linear(o, a.out_proj.weight, a.out_proj.bias) * mask2d(2) * keep_scale(p)
    y1 = F.layer_norm(sa, (E,), layer.norm1.weight, layer.norm1.bias, layer.norm1.eps)
    h = F.relu(F.linear(y1, layer.linear1.weight, layer.linear1.bias)) * mask2d(3) * keep_scale(p)
    f = y1 + F.linear(h, layer.linear2.weight, layer.linear2.bias) * mask2d(4) * keep_scale(p)
    return F.layer_norm(f, (E,), layer.norm2.weight, layer.norm2.bias, layer.norm2.eps)


@pytest.mark.parametrize("dims", [(160, 3, 256), (392, 5, 768)])
def test_group_embed_layer_dropout_matches_masked_reference(dims):
    import copy

    from simple3d_former_b200.models import GroupEmbedLayer
    S, Nb, E = dims
    dev = _dev()
    torch.manual_seed(5)
    layer = GroupEmbedLayer(d_model=E, nhead=4, dim_feedforward=E).to(dev).train()
    assert layer.dropout_p == 0.1
    ref = copy.deepcopy(layer)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(S, Nb, E, generator=g).to(dev)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    w = torch.randn(S, Nb, E, generator=g).to(dev)
    seed_val = int(layer._drop_seed) + 1  # the module bumps its device seed before every training forward
    ya = layer(xa)
    assert int(layer._drop_seed) == seed_val
    yb = _layer_reference(ref, xb, seed_val, 0.1)
    assert _rel(ya, yb) < 2e-2, _rel(ya, yb)
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    torch.cuda.synchronize()
    assert _rel(xa.grad, xb.grad) < 5e-2, _rel(xa.grad, xb.grad)
    for (n, pa), (_, pb) in zip(layer.named_parameters(), ref.named_parameters()):
        assert _rel(pa.grad, pb.grad) < 5e-2, (n, _rel(pa.grad, pb.grad))
    # a second training forward draws new masks; eval() is deterministic and equals the p = 0 arithmetic of torch's layer
    with torch.no_grad():
        y2 = layer(x)
        assert _rel(y2, ya) > 0.05
        layer.eval()
        e1, e2 = layer(x), layer(x)
        assert torch.equal(e1, e2)
        tl = torch.nn.TransformerEncoderLayer(d_model=E, nhead=4, dim_feedforward=E).to(dev).eval()
        tl.load_state_dict(layer.state_dict())
        assert _rel(e1, tl(x)) < 2e-2
