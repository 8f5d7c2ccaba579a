#!/usr/bin/env python
"""Prints the margins of the model-level parity tests (max |logit - reference|, loss error, worst gradient-norm error)
for every golden fixture, so tolerances in tests/test_parity_gpu.py can be judged against what is actually measured.
GPU box: python tools/parity_report.py > gpurun_out/parity_report.txt"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import s3d_oracle as O  # noqa: E402
from simple3d_former_b200.embed_layer_3d_modality import VoxelEmbed, VoxelEmbed_no_average  # noqa: E402
from simple3d_former_b200.models import Feature3D_ViT2D_V2, PointTransformerCls, PointTransformerSeg  # noqa: E402

dev = torch.device("cuda:0")
GOLD = os.path.join(ROOT, "tests", "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), map_location="cpu", weights_only=False)


def grads_report(model, ref):
    named = dict(model.named_parameters())
    worst = (0.0, "")
    for k, r in ref.items():
        g = named[k].grad
        if g is None:
            continue
        e = abs(float(g.float().norm()) - r["norm"]) / (r["norm"] + 1e-12)
        if e > worst[0]:
            worst = (e, k)
    return worst


for name in ("cfg1_deit_small_voxel30", "cfg3_small_deit_base_group36", "cfg3_deit_base_group128", "cfg3_deit_base_group128_b3"):
    fix = load(name)
    sd = O.init_voxel_state_dict(fix["backbone"], fix["cell"], fix["patch"], fix["n_classes"], fix["pos"], seed=fix["weight_seed"])
    g = torch.Generator().manual_seed(fix["embed_seed"])
    for k in ("voxel_pos_embed", "group_pos_embed", "group_cls_token"):
        if k in sd:
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.02
    D = O.BACKBONES[fix["backbone"]]["embed_dim"]
    emb = (VoxelEmbed if fix["average"] else VoxelEmbed_no_average)(fix["V"], fix["cell"], fix["patch"], embed_dim=D)
    m = Feature3D_ViT2D_V2(embed_layer=emb, n_classes=fix["n_classes"], transformer_backbone=fix["backbone"], pretrained=False,
                           pos_embedding=fix["pos"])
    m.load_state_dict(sd, strict=False)
    m = m.to(dev).eval()
    x, y = O.synthetic_voxels(fix["B"], fix["V"], seed=fix["input_seed"], n_classes=fix["n_classes"])
    logits = m(x.to(dev))
    loss = F.cross_entropy(logits, y.to(dev))
    loss.backward()
    err = (logits.detach().cpu() - fix["logits"]).abs().max().item()
    w = grads_report(m, fix["grads"])
    print(f"{name}: max|logit|={fix['logits'].abs().max():.3f} err={err:.2e} loss_err={abs(float(loss) - fix['loss']):.2e} "
          f"worst grad-norm err={w[0]:.3%} ({w[1]})", flush=True)

for name in ("cfg4_point_cls_tiny1024", "cfg5_point_seg_tiny2048", "cfg4_point_cls_tiny1024_sharp", "cfg5_point_seg_tiny2048_sharp"):
    fix = load(name)
    mc = types.SimpleNamespace(nblocks=4, nneighbor=16, transformer_backbone=fix["backbone"], pretrained=False, head="Linear",
                               transformer_dim=512)
    for mode in ("eval", "train"):
        pc = types.SimpleNamespace(num_point=fix["N"], num_class=fix["n_classes"], input_dim=fix["input_dim"], model=mc)
        model = (PointTransformerSeg if fix["seg"] else PointTransformerCls)(pc)
        sd = O.init_point_state_dict(fix["backbone"], fix["input_dim"], fix["n_classes"], seed=fix["weight_seed"])
        if fix.get("sharp"):
            sd = O.sharpen_point_state_dict(sd, head_gain=fix["head_gain"])
        model.load_state_dict(sd, strict=False)
        model = model.to(dev).train(mode == "train")
        model.set_fps_starts([s.to(dev) for s in fix["fps_starts"]])
        x, y = O.synthetic_points(fix["B"], fix["N"], extra=fix["input_dim"] - 3, seed=fix["input_seed"], n_classes=fix["n_classes"])
        if fix["seg"]:
            y = torch.randint(0, fix["n_classes"], (fix["B"], fix["N"]), generator=torch.Generator().manual_seed(fix["label_seed"]))
        logits = model(x.to(dev))
        loss = F.cross_entropy(logits.reshape(-1, fix["n_classes"]), y.reshape(-1).to(dev))
        loss.backward()
        ref = fix[mode]
        err = (logits.detach().cpu() - ref["logits"]).abs().max().item()
        w = grads_report(model, {k: v for k, v in ref["grads"].items() if not (fix["seg"] and k == "cls_token")})
        print(f"{name} {mode}: max|logit|={ref['logits'].abs().max():.3f} err={err:.2e} "
              f"loss_err={abs(float(loss) - ref['loss']):.2e} worst grad-norm err={w[0]:.3%} ({w[1]})", flush=True)
