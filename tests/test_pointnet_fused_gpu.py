"""GPU parity of the fused set-abstraction / feature-propagation path (csrc/pointnet_fused.cu, through the C ABI) against
the same layers evaluated with plain PyTorch fp32 ops exactly as the reference writes them
(data/pointnet_util.py:99-138, 220-244, 381-420; models/3DViT/model.py:33-72).

Tolerances: the fused path feeds bf16 operands to the tensor cores (fp32 accumulate, fp32 BatchNorm statistics), so
outputs / gradients are compared relative to the tensor norm (2e-2 / 5e-2); running statistics within 1e-2."""
import copy

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


def _randomize_bn(mod, gen):
    for m in mod.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            with torch.no_grad():
                # both signs of gamma: a negative scale turns the max over neighbours into a min of the pre-activation
                m.weight.copy_(torch.randn(m.weight.shape, generator=gen) * 0.8)
                m.bias.copy_(torch.randn(m.bias.shape, generator=gen) * 0.3)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.2)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) + 0.5)


def _r16(t):
    """bf16 rounding with a straight-through gradient: places the reference's operands on the same bf16 grid as the
    tensor-core operands of the fused path, so ReLU masks / arg-max choices agree and gradients can be compared tightly
    (an fp32-vs-bf16 comparison of gradients is dominated by a handful of flipped masks, not by arithmetic error)."""
    return t + (t.bfloat16().float() - t).detach()


def _bn(x, bn, training):
    return F.batch_norm(x, bn.running_mean.clone(), bn.running_var.clone(), bn.weight, bn.bias, training, bn.momentum,
                        bn.eps)


def _sa_reference(sa, xyz, pts, training):
    """data/pointnet_util.py:99-138 + :220-244 in fp32 torch ops (bf16-rounded GEMM operands, see _r16)."""
    from simple3d_former_b200 import pointnet_util as P
    B = xyz.shape[0]
    fps_idx = P.farthest_point_sample(xyz, sa.npoint, sa.fps_start)
    bi = torch.arange(B, device=xyz.device)
    new_xyz = xyz[bi[:, None], fps_idx]
    idx = P.knn_point(sa.nsample, xyz, new_xyz)
    (c1, c2), (n1, n2) = sa.mlp_convs, sa.mlp_bns
    W1 = c1.weight.flatten(1)
    gx = xyz[bi[:, None, None], idx] - new_xyz[:, :, None]
    z1 = gx @ W1[:, :3].t() + pts[bi[:, None, None], idx] @ W1[:, 3:].t() + c1.bias  # [B,S,K,C1]; split GEMM = fp32-grade
    a1 = _r16(F.relu(_bn(z1.permute(0, 3, 2, 1), n1, training)))  # [B,C1,K,S]
    z2 = F.conv2d(a1, _r16(c2.weight), c2.bias)
    y2 = F.relu(_bn(z2, n2, training))
    return torch.max(y2, 2)[0].transpose(1, 2)


def _sa_unfused(sa, xyz, pts):
    """The module as the reference writes it (data/pointnet_util.py:220-244): plain fp32 PyTorch ops through the module's
    own Conv2d / BatchNorm2d children (so `sa`'s running statistics are updated like the reference's would be)."""
    from simple3d_former_b200 import pointnet_util as P
    B = xyz.shape[0]
    bi = torch.arange(B, device=xyz.device)
    fps_idx = P.farthest_point_sample(xyz, sa.npoint, sa.fps_start)
    new_xyz = xyz[bi[:, None], fps_idx]
    idx = P.knn_point(sa.nsample, xyz, new_xyz)
    x = torch.cat([xyz[bi[:, None, None], idx] - new_xyz[:, :, None], pts[bi[:, None, None], idx]], dim=-1)
    x = x.permute(0, 3, 2, 1)
    for conv, bn in zip(sa.mlp_convs, sa.mlp_bns):
        x = F.relu(bn(conv(x)))
    return new_xyz, torch.max(x, 2)[0].transpose(1, 2)


def _tu_unfused(tu, xyz1, p1, xyz2, p2):
    """models/3DViT/model.py:47-72 through the module's own Sequential children (plain fp32 PyTorch ops)."""
    from simple3d_former_b200 import pointnet_util as P
    f1, f2 = tu.fc1(p1), tu.fc2(p2)
    B = xyz2.shape[0]
    dd, ii = P.square_distance(xyz2, xyz1).sort(dim=-1)
    dd, ii = dd[:, :, :3], ii[:, :, :3]
    rec = 1.0 / (dd + 1e-8)
    wgt = rec / rec.sum(dim=2, keepdim=True)
    bi = torch.arange(B, device=xyz2.device)
    return (f1[bi[:, None, None], ii] * wgt[..., None]).sum(dim=2) + f2


def _tu_reference(tu, xyz1, p1, xyz2, p2, training):
    """models/3DViT/model.py:33-72 + pointnet_util.py:381-420 in fp32 torch ops (bf16-rounded GEMM operands)."""
    from simple3d_former_b200 import pointnet_util as P

    def fc(seq, x):
        lin, bn = seq[0], seq[2]
        z = x @ lin.weight.t() + lin.bias  # hi/lo-split GEMM in the product = fp32-grade
        return F.relu(_bn(z.transpose(1, 2), bn, training).transpose(1, 2))

    f1, f2 = fc(tu.fc1, p1), fc(tu.fc2, p2)
    B, N, _ = xyz2.shape
    dd, ii = P.square_distance(xyz2, xyz1).sort(dim=-1)
    dd, ii = dd[:, :, :3], ii[:, :, :3]
    rec = 1.0 / (dd + 1e-8)
    wgt = rec / rec.sum(dim=2, keepdim=True)
    bi = torch.arange(B, device=xyz2.device)
    return (f1[bi[:, None, None], ii] * wgt[..., None]).sum(dim=2) + f2


def test_gemm_short_and_ragged_k():
    """K not a multiple of the 64-wide TMA box (48, 96) and N below a tile (48, 96): shapes of the point-tokenizer GEMMs."""
    from simple3d_former_b200 import _lib as L
    dev = _dev()
    g = torch.Generator().manual_seed(0)
    for M, N, K in [(1000, 96, 48), (4096, 96, 96), (777, 48, 96), (2048, 192, 96)]:
        a = torch.randn(M, K, generator=g).to(dev).bfloat16()
        b = torch.randn(N, K, generator=g).to(dev).bfloat16()
        bias = torch.randn(N, generator=g).to(dev)
        want = a.float() @ b.float().t() + bias
        got = L.gemm(a, b, bias=bias, out_dtype=torch.float32)
        assert _rel(got, want) < 1e-5, (M, N, K, _rel(got, want))
        # activation-gradient form (B consumed MN-major) and weight-gradient form (both MN-major, split-K)
        d = torch.randn(M, N, generator=g).to(dev).bfloat16()
        assert _rel(L.gemm(d, b, b_mn=True, out_dtype=torch.float32), d.float() @ b.float()) < 1e-5
        assert _rel(L.gemm(d, a, a_mn=True, b_mn=True, out_dtype=torch.float32), d.float().t() @ a.float()) < 1e-4


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("dims", [(2, 128, 64, 16, 16, 32), (3, 256, 256, 16, 48, 96), (2, 300, 77, 8, 96, 192)])
def test_set_abstraction_fused_matches_torch(dims, training):
    from simple3d_former_b200.pointnet_util import PointNetSetAbstraction
    B, N, S, K, Cf, C = dims
    dev = _dev()
    torch.manual_seed(11)  # layer initialisation does not depend on which tests ran before
    g = torch.Generator().manual_seed(1)
    sa = PointNetSetAbstraction(S, 0, K, Cf + 3, [C, C], False, knn=True)
    _randomize_bn(sa, g)
    sa = sa.to(dev).train(training)
    sa.fps_start = torch.zeros(B, dtype=torch.long, device=dev)
    ref = copy.deepcopy(sa)
    xyz = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    pts = torch.randn(B, N, Cf, generator=g).to(dev)
    pa, pb = pts.clone().requires_grad_(True), pts.clone().requires_grad_(True)
    w = torch.randn(B, S, C, generator=g).to(dev)
    xa, ya = sa(xyz, pa)
    with torch.no_grad():  # the layer as the reference writes it, in fp32 PyTorch ops (also updates ref's buffers)
        xb, yb32 = _sa_unfused(ref, xyz, pts)
    assert torch.equal(xa, xb)
    assert ya.shape == yb32.shape == (B, S, C)
    assert _rel(ya, yb32) < 2e-2, _rel(ya, yb32)
    yb = _sa_reference(ref, xyz, pb, training)
    assert _rel(ya, yb) < 5e-3, _rel(ya, yb)
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    torch.cuda.synchronize()
    assert _rel(pa.grad, pb.grad) < 5e-2, _rel(pa.grad, pb.grad)
    na, nb = dict(sa.named_parameters()), dict(ref.named_parameters())
    wscale = max(float(nb[k].grad.norm()) for k in nb if nb[k].grad is not None and "weight" in k)
    for k, p in nb.items():
        if p.grad is None:
            assert na[k].grad is None or float(na[k].grad.abs().max()) == 0.0, k
            continue
        if training and k.startswith("mlp_convs") and k.endswith("bias"):
            # a bias in front of a training-mode BatchNorm has an exactly-zero gradient: both sides are rounding noise
            assert float(na[k].grad.norm()) < 2e-2 * wscale, k
            continue
        assert _rel(na[k].grad, p.grad) < 5e-2, (k, _rel(na[k].grad, p.grad))
    for (ka, ba), (kb, bb) in zip(sa.named_buffers(), ref.named_buffers()):
        if "num_batches" in ka:
            assert int(ba) == int(bb), ka
        else:
            assert _rel(ba, bb) < 1e-2, (ka, _rel(ba, bb))
    if not training:  # bitwise reproducible forward (the training-mode statistics pass is covered below)
        with torch.no_grad():
            assert torch.equal(sa(xyz, pts)[1], sa(xyz, pts)[1])


def test_set_abstraction_forward_reproducible_in_training():
    from simple3d_former_b200.pointnet_util import PointNetSetAbstraction
    dev = _dev()
    torch.manual_seed(13)
    g = torch.Generator().manual_seed(2)
    sa = PointNetSetAbstraction(128, 0, 16, 48 + 3, [96, 96], False, knn=True).to(dev).train()
    sa.fps_start = torch.zeros(4, dtype=torch.long, device=dev)
    xyz = (torch.rand(4, 512, 3, generator=g) * 2 - 1).to(dev)
    pts = torch.randn(4, 512, 48, generator=g).to(dev)
    with torch.no_grad():
        outs = [sa(xyz, pts)[1].clone() for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("training", [True, False])
def test_transition_up_fused_matches_torch(training):
    from simple3d_former_b200.models import TransitionUp
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    B, S, N, d1, d2, do = 3, 64, 256, 64, 32, 32
    torch.manual_seed(12)
    tu = TransitionUp(d1, d2, do)
    _randomize_bn(tu, g)
    tu = tu.to(dev).train(training)
    ref = copy.deepcopy(tu)
    xyz2 = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    xyz1 = xyz2[:, :S].contiguous()
    p1 = torch.randn(B, S, d1, generator=g).to(dev)
    p2 = torch.randn(B, N, d2, generator=g).to(dev)
    a1, a2 = p1.clone().requires_grad_(True), p2.clone().requires_grad_(True)
    b1, b2 = p1.clone().requires_grad_(True), p2.clone().requires_grad_(True)
    w = torch.randn(B, N, do, generator=g).to(dev)
    ya = tu(xyz1, a1, xyz2, a2)
    with torch.no_grad():  # the layer as the reference writes it, in fp32 PyTorch ops (also updates ref's buffers)
        yb32 = _tu_unfused(ref, xyz1, p1, xyz2, p2)
    assert ya.shape == yb32.shape == (B, N, do)
    assert _rel(ya, yb32) < 2e-2, _rel(ya, yb32)
    yb = _tu_reference(ref, xyz1, b1, xyz2, b2, training)
    assert _rel(ya, yb) < 5e-3, _rel(ya, yb)
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    torch.cuda.synchronize()
    assert _rel(a1.grad, b1.grad) < 5e-2 and _rel(a2.grad, b2.grad) < 5e-2
    na, nb = dict(tu.named_parameters()), dict(ref.named_parameters())
    wscale = max(float(p.grad.norm()) for p in nb.values() if p.grad is not None)
    for k, p in nb.items():
        if p.grad is None:
            continue
        if training and k.endswith("0.bias"):  # Linear bias in front of a training-mode BatchNorm: zero gradient
            assert float(na[k].grad.norm()) < 2e-2 * wscale, k
            continue
        assert _rel(na[k].grad, p.grad) < 5e-2, (k, _rel(na[k].grad, p.grad))
    for (ka, ba), (kb, bb) in zip(tu.named_buffers(), ref.named_buffers()):
        if "num_batches" in ka:
            assert int(ba) == int(bb), ka
        else:
            assert _rel(ba, bb) < 1e-2, (ka, _rel(ba, bb))


def test_three_nn_interpolation_exact_against_torch():
    """fp32 kernel: same arithmetic as the reference's sort-based 3-NN interpolation (tolerance 1e-6 relative)."""
    from simple3d_former_b200 import functional as Fn
    from simple3d_former_b200 import pointnet_util as P
    dev = _dev()
    g = torch.Generator().manual_seed(4)
    B, S, N, C = 2, 50, 333, 40
    src = (torch.rand(B, S, 3, generator=g) * 2 - 1).to(dev)
    qry = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    feats = torch.randn(B, S, C, generator=g).to(dev).requires_grad_(True)
    idx, dist = P.knn_point(3, src, qry, return_dist=True)
    out = Fn.ThreeNNInterpFn.apply(feats, idx, dist, None)
    d = P.square_distance(qry, src)
    dd, ii = d.sort(dim=-1)
    dd, ii = dd[:, :, :3], ii[:, :, :3]
    rec = 1.0 / (dd + 1e-8)
    wgt = rec / rec.sum(dim=2, keepdim=True)
    f2 = feats.detach().clone().requires_grad_(True)
    gathered = torch.gather(f2[:, None].expand(B, N, S, C), 2, ii[..., None].expand(B, N, 3, C))
    want = (gathered * wgt[..., None]).sum(dim=2)
    assert torch.equal(idx, ii)
    assert _rel(out, want) < 1e-6
    out.sum().backward()
    want.sum().backward()
    assert _rel(feats.grad, f2.grad) < 1e-5
