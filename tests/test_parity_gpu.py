"""GPU parity: the sm_100a product path (through the C ABI) against the reference's golden outputs and the oracle.
Tolerances are north_star's: logits / loss within 1e-2 (bf16 operands, fp32 accumulate); integer indices bit-exact."""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import s3d_oracle as O

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-2  # north_star: logits / loss within 1e-2 for bf16 operands -- an ABSOLUTE bound


def _logit_tol(ref_logits):
    """1e-2 absolute (north_star); fixtures whose logits are small get the TIGHTER bound of 2 % of max|logit| (an absolute
    1e-2 on |logit| < 0.1 would pass a visibly wrong kernel; bf16 operand rounding over 12 blocks measures 0.3-1 % of the
    logit scale on every fixture, tools/parity_report.py)."""
    return min(LOGIT_TOL, 2e-2 * float(ref_logits.abs().max()))


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _voxel_sd(fix):
    sd = O.init_voxel_state_dict(fix["backbone"], fix["cell"], fix["patch"], fix["n_classes"], fix["pos"], seed=fix["weight_seed"])
    g = torch.Generator().manual_seed(fix["embed_seed"])
    for k in ("voxel_pos_embed", "group_pos_embed", "group_cls_token"):
        if k in sd:
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.02
    assert abs(O.state_dict_checksum(sd) - fix["sd_checksum"]) <= 1e-6 * abs(fix["sd_checksum"])
    return sd


def _build_voxel(fix):
    from simple3d_former_b200.embed_layer_3d_modality import VoxelEmbed, VoxelEmbed_no_average
    from simple3d_former_b200.models import Feature3D_ViT2D_V2
    D = O.BACKBONES[fix["backbone"]]["embed_dim"]
    emb = (VoxelEmbed if fix["average"] else VoxelEmbed_no_average)(fix["V"], fix["cell"], fix["patch"], embed_dim=D)
    m = Feature3D_ViT2D_V2(embed_layer=emb, n_classes=fix["n_classes"], transformer_backbone=fix["backbone"],
                           pretrained=False, pos_embedding=fix["pos"])
    res = m.load_state_dict(_voxel_sd(fix), strict=False)
    assert not res.unexpected_keys
    assert all(k.startswith(("pos_embed", "patch_embed.", "head.")) for k in res.missing_keys), res.missing_keys
    return m.to(_dev())


def _check_grads(model, ref_grads, rel=3e-2, skip=(), head_rel=None):
    """Gradient fingerprints: L2 norm within `rel`, and the first 16 elements within `head_rel` (default 4*rel) of their
    own L2 norm (bf16 operand rounding is relative to the tensor's scale, not to each tiny element)."""
    head_rel = 4 * rel if head_rel is None else head_rel
    named = dict(model.named_parameters())
    for k, ref in ref_grads.items():
        if k in skip:
            continue
        g = named[k].grad
        assert g is not None, k
        g = g.detach().float().cpu()
        assert abs(float(g.norm()) - ref["norm"]) <= rel * ref["norm"] + 1e-7, (k, float(g.norm()), ref["norm"])
        head = g.flatten()[:16]
        scale = max(float(ref["head"].norm()), ref["norm"] * (16 / max(g.numel(), 16)) ** 0.5)
        assert float((head - ref["head"]).norm()) <= head_rel * scale + 1e-7, (k, head, ref["head"])


@pytest.mark.parametrize("name", ["cfg1_deit_small_voxel30", "cfg3_small_deit_base_group36", "cfg3_deit_base_group128",
                                  "cfg3_deit_base_group128_b3"])
def test_voxel_model_matches_reference(golden, name):
    """cfg3_deit_base_group128_b3 has S = 3*196 = 588 >= 512 group_embed tokens: the model-level comparison runs the
    tcgen05 flash-attention forward AND backward kernels (the other group_embed fixtures stay below the switch-over)."""
    fix = golden(name)
    # the golden logits come from the reference in eval(): the only train/eval difference of the voxel models is the
    # group_embed layer's dropout (p = 0.1), which is stochastic in the reference too (tests/test_dropout_gpu.py covers it)
    model = _build_voxel(fix).eval()
    x, y = O.synthetic_voxels(fix["B"], fix["V"], seed=fix["input_seed"], n_classes=fix["n_classes"])
    logits = model(x.to(_dev()))
    loss = F.cross_entropy(logits, y.to(_dev()))
    loss.backward()
    torch.cuda.synchronize()
    err = (logits.detach().cpu() - fix["logits"]).abs().max().item()
    assert err <= _logit_tol(fix["logits"]), f"logits differ from the reference by {err}"
    assert abs(float(loss) - fix["loss"]) <= LOGIT_TOL
    _check_grads(model, fix["grads"])


def _point_cfg(fix):
    model = types.SimpleNamespace(nblocks=4, nneighbor=16, transformer_backbone=fix["backbone"], pretrained=False,
                                  head="Linear", transformer_dim=512)
    return types.SimpleNamespace(num_point=fix["N"], num_class=fix["n_classes"], input_dim=fix["input_dim"], model=model)


@pytest.mark.parametrize("name", ["cfg4_point_cls_tiny1024", "cfg5_point_seg_tiny2048", "cfg4_point_cls_tiny1024_sharp",
                                  "cfg5_point_seg_tiny2048_sharp"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_point_model_matches_reference(golden, name, mode):
    """The *_sharp fixtures scale the qkv and head weights (oracle.sharpen_point_state_dict) so that |logit| ~ 0.15-1.7
    and attention is far from uniform; at the reference's init (the other two) |logit| < 0.08."""
    from simple3d_former_b200.models import PointTransformerCls, PointTransformerSeg
    fix = golden(name)
    model = (PointTransformerSeg if fix["seg"] else PointTransformerCls)(_point_cfg(fix))
    sd = O.init_point_state_dict(fix["backbone"], fix["input_dim"], fix["n_classes"], seed=fix["weight_seed"])
    if fix.get("sharp"):
        sd = O.sharpen_point_state_dict(sd, head_gain=fix["head_gain"])
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    model = model.to(_dev()).train(mode == "train")
    model.set_fps_starts([s.to(_dev()) for s in fix["fps_starts"]])
    x, y = O.synthetic_points(fix["B"], fix["N"], extra=fix["input_dim"] - 3, seed=fix["input_seed"], n_classes=fix["n_classes"])
    if fix["seg"]:
        y = torch.randint(0, fix["n_classes"], (fix["B"], fix["N"]), generator=torch.Generator().manual_seed(fix["label_seed"]))
    logits = model(x.to(_dev()))
    loss = F.cross_entropy(logits.reshape(-1, fix["n_classes"]), y.reshape(-1).to(_dev()))
    loss.backward()
    torch.cuda.synchronize()
    err = (logits.detach().cpu() - fix[mode]["logits"]).abs().max().item()
    assert err <= _logit_tol(fix[mode]["logits"]), f"logits differ from the reference by {err}"
    assert abs(float(loss) - fix[mode]["loss"]) <= LOGIT_TOL
    # At the reference's init the point models have near-uniform attention, so the softmax gradient P*(dP - sum(P*dP))
    # is a difference of nearly equal numbers and bf16 operand rounding is amplified (up to 5.7 % on the L2 norm of
    # blocks.0.attn.qkv.weight.grad): those two fixtures keep 8 %; the sharpened fixtures are held to 3 %.
    # Both point models discard the cls token's output row (models/3DViT/model.py:322, :519: `x = x[:, 1:]`), so its
    # gradient only arrives through the other tokens' attention to it: ~1e-6 and dominated by rounding noise -> skipped.
    # Element-level fingerprints (first 16 entries) of the point models pass through training-mode BatchNorms over a batch
    # of 2 clouds, which amplify operand rounding of single entries far more than the tensor norm: 20 % of the scale.
    _check_grads(model, fix[mode]["grads"], rel=3e-2 if fix.get("sharp") else 8e-2, skip=("cls_token",),
                 head_rel=0.2 if fix.get("sharp") else None)


def test_forward_images_matches_reference(golden):
    """Feature3D_ViT2D_V2.forward_images (vit_3d_2d_pretrain.py:435-451): the 2-D DeiT
    path through the SAME fused blocks at N = 197, against the reference's golden logits / loss / gradients."""
    from simple3d_former_b200.embed_layer_3d_modality import VoxelEmbed
    from simple3d_former_b200.models import Feature3D_ViT2D_V2
    fix = golden("cfg1_forward_images")
    sd = O.init_voxel_state_dict(fix["backbone"], 6, 5, 40, "default", seed=fix["weight_seed"])
    sd.update(O.init_image_branch_state_dict(fix["backbone"], seed=fix["image_seed"]))
    assert abs(O.state_dict_checksum(sd) - fix["sd_checksum"]) <= 1e-6 * abs(fix["sd_checksum"])
    model = Feature3D_ViT2D_V2(embed_layer=VoxelEmbed(30, 6, 5, embed_dim=384), n_classes=40,
                               transformer_backbone=fix["backbone"], pretrained=False, pos_embedding="default")
    model.load_state_dict(sd, strict=True)
    model = model.to(_dev()).eval()
    x = torch.randn(fix["B"], 3, 224, 224, generator=torch.Generator().manual_seed(fix["input_seed"]))
    y = torch.randint(0, 1000, (fix["B"],), generator=torch.Generator().manual_seed(fix["label_seed"]))
    logits = model.forward_images(x.to(_dev()))
    loss = F.cross_entropy(logits, y.to(_dev()))
    loss.backward()
    torch.cuda.synchronize()
    assert logits.shape == (fix["B"], 1000)
    assert (logits.detach().cpu() - fix["logits"]).abs().max().item() <= LOGIT_TOL
    assert abs(float(loss) - fix["loss"]) <= LOGIT_TOL
    _check_grads(model, fix["grads"])


def test_sgd_momentum_step_matches_torch():
    """s3d_sgd_momentum_step -- the optimizer of the point configurations (train_cls.py:91, train_partseg.py:95:
    torch.optim.SGD(lr=0.01, momentum=0.9)) -- against torch.optim.SGD over several steps, with and without weight decay
    and gradient averaging; the bf16 shadow must be the rounded updated parameter."""
    from simple3d_former_b200 import _lib as L
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    n = 100_003  # not a multiple of the vector width
    for wd, scale in ((0.0, 1.0), (1e-4, 0.25)):
        p0 = torch.randn(n, generator=g)
        p_ref = p0.clone().to(dev).requires_grad_(True)
        opt = torch.optim.SGD([p_ref], lr=0.01, momentum=0.9, weight_decay=wd)
        p = p0.clone().to(dev)
        buf = torch.zeros(n, device=dev)
        shadow = torch.zeros(n, device=dev, dtype=torch.bfloat16)
        step_t = torch.zeros(1, device=dev, dtype=torch.int32)
        for step in range(1, 5):
            grad = torch.randn(n, generator=g).to(dev)
            p_ref.grad = grad * scale
            opt.step()
            step_t += 1
            L.sgd_momentum_step(p, grad.clone(), buf, shadow, 0.01, 0.9, wd, step, grad_scale=scale, step_tensor=step_t)
            torch.cuda.synchronize()
            assert (p - p_ref.detach()).abs().max().item() <= 2e-6, (wd, step)
            assert torch.equal(shadow, p.bfloat16()), (wd, step)
        assert (buf - opt.state[p_ref]["momentum_buffer"]).abs().max().item() <= 2e-6


def test_trainer_sgd_matches_torch_on_point_model(golden):
    """DataParallelTrainer(optimizer='sgd') on the cfg4 model: one step == torch.optim.SGD(momentum=0.9) on autograd's
    gradients of the same model (dead parameters excluded as in bench.py)."""
    from simple3d_former_b200.dp import DataParallelTrainer
    from simple3d_former_b200.models import PointTransformerCls
    fix = golden("cfg4_point_cls_tiny1024_sharp")
    sd = O.sharpen_point_state_dict(O.init_point_state_dict(fix["backbone"], fix["input_dim"], fix["n_classes"],
                                                            seed=fix["weight_seed"]), head_gain=fix["head_gain"])
    x, y = O.synthetic_points(fix["B"], fix["N"], extra=fix["input_dim"] - 3, seed=fix["input_seed"], n_classes=fix["n_classes"])
    x, y = x.to(_dev()), y.to(_dev())

    def build():
        m = PointTransformerCls(_point_cfg(fix))
        m.load_state_dict(sd, strict=False)
        m = m.to(_dev()).train()
        m.set_fps_starts([s.to(_dev()) for s in fix["fps_starts"]])
        return m

    ref_model = build()
    dead = set(ref_model.unused_parameter_names())
    F.cross_entropy(ref_model(x), y).backward()
    live = [p for n, p in ref_model.named_parameters() if n not in dead]
    assert all(p.grad is not None for p in live)
    torch.optim.SGD(live, lr=0.01, momentum=0.9).step()
    model = build()
    trainer = DataParallelTrainer(model, lr=0.01, exclude=model.unused_parameter_names(), optimizer="sgd", momentum=0.9)
    trainer.step(x, y, F.cross_entropy)
    torch.cuda.synchronize()
    ref_named = dict(ref_model.named_parameters())
    init = {k: v.to(_dev()) for k, v in sd.items()}
    for n, p in model.named_parameters():
        if n in dead:
            continue
        # the UPDATE (p_after - p_before) of both sides, compared in L2: the two runs are separate executions whose fp32
        # atomics (scatter-adds of the point path, split-K) sum in different orders, and training-mode BatchNorms over a
        # batch of 2 clouds amplify that for single entries -- element-wise bounds on one step are flaky, the norm is not
        upd, upd_ref = p.detach() - init[n], ref_named[n].detach() - init[n]
        scale = float(upd_ref.norm())
        if scale < 1e-3 * 0.01 * upd_ref.numel() ** 0.5:
            # gradient RMS below 1e-3: biases in front of a training-mode BatchNorm (conv / Linear biases of the set
            # abstraction, TransitionUp and the input MLPs) have an exactly-zero true gradient -- both sides are noise
            assert float(upd.norm()) < 1e-3 * 0.01 * upd_ref.numel() ** 0.5, n
            continue
        assert float((upd - upd_ref).norm()) <= 5e-2 * scale + 1e-7, (n, float((upd - upd_ref).norm()), scale)


def test_block_gradient_side_products_with_two_consumers():
    """A Block's backward leaves a bf16 copy and the column sums of its input gradient on the tensor for the upstream
    Block. When that activation has a SECOND consumer, autograd adds the other gradient (possibly in place, same Python
    object and data pointer): the side products must then be ignored. By linearity the parameter gradients of
    loss1 + loss2 equal the sum of the gradients of two single-consumer graphs."""
    from simple3d_former_b200.vision_transformer import Block
    torch.manual_seed(4)
    dev = _dev()
    B, N, D, H = 4, 50, 192, 3
    blk1 = Block(D, H, mlp_ratio=4.0, qkv_bias=True).to(dev)
    blk2 = Block(D, H, mlp_ratio=4.0, qkv_bias=True).to(dev)
    x = torch.randn(B, N, D, device=dev)
    w = torch.randn(B, N, D, device=dev) * 30.0  # the side branch dominates: a stale side product cannot hide
    params = list(blk1.parameters())

    def grads(loss_fn):
        for p in params:
            p.grad = None
        h = blk1(x)
        loss_fn(h).backward()
        return [p.grad.detach().clone() for p in params]

    g_chain = grads(lambda h: blk2(h).sum())
    g_side = grads(lambda h: (h * w).sum())
    for order in (lambda h: blk2(h).sum() + (h * w).sum(), lambda h: (h * w).sum() + blk2(h).sum()):
        g_both = grads(order)
        for (n, _), a, b, c in zip(blk1.named_parameters(), g_both, g_chain, g_side):
            want = b + c
            assert float((a - want).norm()) <= 2e-2 * float(want.norm()) + 1e-6, (n, float((a - want).norm()), float(want.norm()))
