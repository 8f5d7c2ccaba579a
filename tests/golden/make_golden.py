"""Generates tests/golden/*.pt by running the UNMODIFIED reference modules (through oracle/reference_harness.py) on
seeded synthetic inputs. Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Weights are NOT stored (88-360 MB): both sides regenerate them from a seed with oracle/s3d_oracle.init_*_state_dict and
the fixture carries a checksum. Stored: the reference's logits / loss / gradient fingerprints / neighbour indices.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_harness as H  # noqa: E402
import s3d_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def load_into(model, sd):
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    used = set(sd)
    for k in res.missing_keys:  # only dead / 2-D-image parameters may be absent
        assert (k.startswith(("patch_embed.", "head.", "pos_embed")) or "last_pos_embed" in k
                or "num_batches_tracked" in k), k
    return used


def grad_fingerprint(model, names):
    fp = {}
    for n, p in model.named_parameters():
        if n in names and p.grad is not None:
            g = p.grad.detach()
            fp[n] = dict(norm=float(g.norm()), head=g.flatten()[:16].clone(), sum=float(g.double().sum()))
    return fp


VOX_GRAD_KEYS = ["blocks.0.attn.qkv.weight", "blocks.0.attn.qkv.bias", "blocks.0.norm1.weight", "blocks.11.mlp.fc2.weight",
                 "blocks.11.mlp.fc1.bias", "blocks.5.attn.proj.weight", "voxel_embed.proj.conv3d_1.weight",
                 "voxel_embed.proj.conv3d_1.bias", "voxel_head.weight", "cls_token", "voxel_pos_embed", "norm.weight",
                 "group_embed.self_attn.in_proj_weight", "group_embed.linear1.weight", "group_embed.norm2.weight",
                 "group_pos_embed", "group_cls_token"]


def voxel_case(ref, name, backbone, V, cell, patch, pos, B, n_classes, average):
    D = O.BACKBONES[backbone]["embed_dim"]
    emb = (ref.embed.VoxelEmbed if average else ref.embed.VoxelEmbed_no_average)(V, cell, patch, embed_dim=D)
    model = ref.vit.Feature3D_ViT2D_V2(embed_layer=emb, n_classes=n_classes, transformer_backbone=backbone,
                                       pretrained=False, pos_embedding=pos).eval()
    sd = O.init_voxel_state_dict(backbone, cell, patch, n_classes, pos, seed=9)
    # exercise the (zero-initialised in the reference) positional / cls embeddings with non-trivial values
    g = torch.Generator().manual_seed(10)
    for k in ("voxel_pos_embed", "group_pos_embed", "group_cls_token"):
        if k in sd:
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.02
    load_into(model, sd)
    x, y = O.synthetic_voxels(B, V, seed=9, n_classes=n_classes)
    logits = model(x)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    fix = dict(kind="voxel", backbone=backbone, V=V, cell=cell, patch=patch, pos=pos, B=B, n_classes=n_classes,
               average=average, weight_seed=9, embed_seed=10, input_seed=9, sd_checksum=O.state_dict_checksum(sd),
               logits=logits.detach().clone(), loss=float(loss), grads=grad_fingerprint(model, VOX_GRAD_KEYS))
    torch.save(fix, os.path.join(OUT, name + ".pt"))
    print(name, "loss", float(loss), "logits", tuple(logits.shape))


PT_GRAD_KEYS = ["blocks.0.attn.qkv.weight", "blocks.11.mlp.fc2.weight", "fc1.0.weight", "fc_pos_embed.2.weight",
                "transition_downs.0.sa.mlp_convs.0.weight", "transition_downs.1.sa.mlp_convs.1.weight",
                "transition_downs.1.sa.mlp_bns.1.weight", "transition_ups.0.fc1.0.weight", "transition_ups.1.fc2.0.weight",
                "head.weight", "cls_token", "norm.weight"]


def point_case(ref, name, seg, backbone, N, input_dim, n_classes, B):
    cfg = H.point_cfg(N, n_classes, input_dim, backbone=backbone)
    cls = ref.point.PointTransformerSeg if seg else ref.point.PointTransformerCls
    model = cls(cfg)
    sd = O.init_point_state_dict(backbone, input_dim, n_classes, seed=9)
    load_into(model, sd)
    x, y = O.synthetic_points(B, N, extra=input_dim - 3, seed=9, n_classes=n_classes)
    if seg:
        gy = torch.Generator().manual_seed(11)
        y = torch.randint(0, n_classes, (B, N), generator=gy)
    out = {}
    for mode in ("eval", "train"):
        model.train(mode == "train")
        model.zero_grad()
        model.load_state_dict(sd, strict=False)  # reset BN running stats
        torch.manual_seed(1234)  # fixes the torch.randint FPS start points (pointnet_util.py:65)
        logits = model(x)
        loss = F.cross_entropy(logits.reshape(-1, n_classes), y.reshape(-1))
        loss.backward()
        out[mode] = dict(logits=logits.detach().clone(), loss=float(loss), grads=grad_fingerprint(model, PT_GRAD_KEYS))
        print(name, mode, "loss", float(loss), tuple(logits.shape))
    torch.manual_seed(1234)
    starts = [torch.randint(0, N, (B,)), torch.randint(0, N, (B,))]  # same draws, same order as the two FPS calls
    fix = dict(kind="point", seg=seg, backbone=backbone, N=N, input_dim=input_dim, n_classes=n_classes, B=B,
               weight_seed=9, input_seed=9, label_seed=11, fps_starts=starts, sd_checksum=O.state_dict_checksum(sd), **out)
    torch.save(fix, os.path.join(OUT, name + ".pt"))


def pointops_case(ref):
    pu = ref.pointnet_util
    g = torch.Generator().manual_seed(21)
    fix = dict(kind="pointops", seed=21, cases=[])
    for (B, N, S, K) in [(2, 1024, 1024, 16), (2, 1024, 256, 16), (1, 2048, 512, 16), (3, 257, 100, 3), (2, 64, 64, 16)]:
        xyz = torch.rand(B, N, 3, generator=g) * 2 - 1
        q = xyz[:, torch.randperm(N, generator=g)[:S]].contiguous()
        d = pu.square_distance(q, xyz)
        ds, order = d.sort(dim=-1, stable=True)
        # the reference's default argsort is unstable on ties; fixtures are tie-free up to K+1, checked here
        assert bool((ds[:, :, 1:K + 1] > ds[:, :, :K]).all())
        knn = d.argsort()[:, :, :K].clone()
        assert torch.equal(knn, order[:, :, :K])
        ball = pu.query_ball_point(0.2, 16, xyz, q)
        torch.manual_seed(300 + N + S)
        fps = pu.farthest_point_sample(xyz, S)
        torch.manual_seed(300 + N + S)
        start = torch.randint(0, N, (B,))
        pts = torch.randn(B, N, 7, generator=g)
        gathered = pu.index_points(pts, knn)
        fix["cases"].append(dict(B=B, N=N, S=S, K=K, xyz=xyz.clone(), query=q.clone(), knn=knn.clone(),
                                 knn_dist=ds[:, :, :K].clone(), ball=ball.clone(), radius=0.2, nsample=16, fps=fps.clone(),
                                 fps_start=start.clone(), gather_checksum=float(gathered.double().sum())))
        print("pointops", B, N, S, K)
    # duplicated points: tie contract = ascending (distance, index); reference argsort is implementation-defined here
    xyz = torch.rand(1, 128, 3, generator=g)
    xyz[:, 64:] = xyz[:, :64]
    fix["tie_case"] = dict(xyz=xyz, K=16, knn_stable=pu.square_distance(xyz, xyz).sort(dim=-1, stable=True)[1][:, :, :16].clone())
    torch.save(fix, os.path.join(OUT, "pointops.pt"))


def main():
    assert H.available(), "reference tree not found"
    ref = H.load()
    torch.set_num_threads(os.cpu_count())
    pointops_case(ref)
    voxel_case(ref, "cfg1_deit_small_voxel30", "deit_small_patch16_224", 30, 6, 5, "default", 8, 40, True)
    voxel_case(ref, "cfg3_small_deit_base_group36", "deit_base_patch16_224", 36, 9, 4, "group_embed", 3, 55, False)
    voxel_case(ref, "cfg3_deit_base_group128", "deit_base_patch16_224", 128, 9, 14, "group_embed", 2, 55, False)
    point_case(ref, "cfg4_point_cls_tiny1024", False, "deit_tiny_patch16_224", 1024, 6, 40, 2)
    point_case(ref, "cfg5_point_seg_tiny2048", True, "deit_tiny_patch16_224", 2048, 22, 50, 2)


if __name__ == "__main__":
    main()
