"""CPU oracle of the Simple3D-Former encoder hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this module; the
product (simple3d_former_b200/) never does. It is a *functional* restatement: every function takes the reference's own
state-dict (same key names as the reference modules) and evaluates the reference arithmetic with plain fp32 torch/numpy
ops on CPU. Each function cites the reference file:line (tree: VITA-Group/Simple3D-Former @ a6f74c8) or, for the encoder
blocks, the timm==0.3.2 symbol (third-party, absent from the tree, pinned by requirements.txt:6).

Pinning: the reference ships no tests or golden vectors ("parity unpinned" upstream). This oracle is pinned
against the reference's own modules executed in the build container (oracle/reference_harness.py): tests/golden/*.pt hold
seeded inputs, reference state-dicts and reference outputs produced by tests/golden/make_golden.py, and
tests/test_oracle.py checks this file against them on CPU.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

BACKBONES = {  # models/vit_3d_2d_pretrain.py:279-325, models/3DViT/model.py:148-196 (note: base = 3 heads)
    "deit_tiny_patch16_224": dict(embed_dim=192, depth=12, num_heads=3),
    "deit_small_patch16_224": dict(embed_dim=384, depth=12, num_heads=6),
    "deit_base_patch16_224": dict(embed_dim=768, depth=12, num_heads=3),
}


# ----------------------------------------------------------------------------------------------------------------
# timm-0.3.2 encoder (Attention / Mlp / Block), functional
# ----------------------------------------------------------------------------------------------------------------
def attention(sd, pre, x, num_heads):
    """timm 0.3.2 Attention.forward."""
    B, N, C = x.shape
    dh = C // num_heads
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(B, N, 3, num_heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = torch.softmax((q @ k.transpose(-2, -1)) * dh ** -0.5, dim=-1)
    y = (a @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def mlp(sd, pre, x):
    """timm 0.3.2 Mlp.forward (exact erf GELU, dropout p=0)."""
    h = F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"]))
    return F.linear(h, sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def block(sd, pre, x, num_heads, eps=1e-6):
    """timm 0.3.2 Block.forward with norm_layer = LayerNorm(eps=1e-6) (vit_3d_2d_pretrain.py:287)."""
    D = x.shape[-1]
    x = x + attention(sd, pre + "attn.", F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps),
                      num_heads)
    x = x + mlp(sd, pre + "mlp.", F.layer_norm(x, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps))
    return x


def encoder(sd, x, depth, num_heads):
    for i in range(depth):
        x = block(sd, f"blocks.{i}.", x, num_heads)
    D = x.shape[-1]
    return F.layer_norm(x, (D,), sd["norm.weight"], sd["norm.bias"], 1e-6)


# ----------------------------------------------------------------------------------------------------------------
# voxel tokenizers + Feature3D_ViT2D_V2
# ----------------------------------------------------------------------------------------------------------------
def voxel_embed(sd, x, cell, average, pre="voxel_embed."):
    """VoxelEmbed.forward (embed_layer_3d_modality.py:33-40) / VoxelEmbed_no_average.forward (:63-70)."""
    y = F.conv3d(x, sd[pre + "proj.conv3d_1.weight"], sd[pre + "proj.conv3d_1.bias"], stride=cell)
    return y.mean(dim=4) if average else y


def group_embed_layer(sd, x, nhead=4, eps=1e-5, pre="group_embed."):
    """nn.TransformerEncoderLayer(d_model=D, nhead=4, dim_feedforward=D), post-norm, ReLU, batch_first=False
    (vit_3d_2d_pretrain.py:381, called at :479): x is [S, Nb, E] and attention runs over dim 0. Dropout (p=0.1) is an
    identity here: parity is defined in eval mode (group_embed semantics)."""
    S, Nb, E = x.shape
    dh = E // nhead
    qkv = F.linear(x, sd[pre + "self_attn.in_proj_weight"], sd[pre + "self_attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)

    def heads(t):  # [S, Nb, E] -> [Nb*nhead, S, dh]
        return t.reshape(S, Nb * nhead, dh).transpose(0, 1)

    q, k, v = heads(q), heads(k), heads(v)
    a = torch.softmax((q * dh ** -0.5) @ k.transpose(-2, -1), dim=-1)
    o = (a @ v).transpose(0, 1).reshape(S, Nb, E)
    o = F.linear(o, sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"])
    x = F.layer_norm(x + o, (E,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
    h = F.linear(F.relu(F.linear(x, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"])), sd[pre + "linear2.weight"],
                 sd[pre + "linear2.bias"])
    return F.layer_norm(x + h, (E,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)


def voxel_vit_features(sd, x, backbone, cell, patch, pos_embedding="default"):
    """Feature3D_ViT2D_V2.forward_features (vit_3d_2d_pretrain.py:453-496), paths 'default' and 'group_embed'."""
    cfg = BACKBONES[backbone]
    depth, H = cfg["depth"], cfg["num_heads"]
    B = x.shape[0]
    if pos_embedding in (None, "default"):
        t = voxel_embed(sd, x, cell, average=True).flatten(2).transpose(1, 2)  # [B, p*p, D]
        t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1) + sd["voxel_pos_embed"]
        return encoder(sd, t, depth, H)[:, 0]
    if pos_embedding == "group_embed":
        t = voxel_embed(sd, x, cell, average=False)  # [B, D, px, py, pz]
        D = t.shape[1]
        t = t.permute(0, 2, 3, 4, 1).reshape(B * patch * patch, patch, D)  # '(b px py) pz c'
        t = torch.cat((sd["group_cls_token"].expand(t.shape[0], -1, -1), t), dim=1) + sd["group_pos_embed"]
        t = group_embed_layer(sd, t)  # sequence-first: attention across (b px py)
        t = encoder(sd, t, depth, H)[:, 0]  # [(b px py), D]
        t = t.reshape(B, patch * patch, D)
        t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1) + sd["voxel_pos_embed"]
        return encoder(sd, t, depth, H)[:, 0]  # same blocks applied a second time (:493-495)
    raise ValueError("Unknown positional embedding scheme!")


def voxel_vit_logits(sd, x, backbone, cell, patch, pos_embedding="default"):
    """Feature3D_ViT2D_V2.forward (vit_3d_2d_pretrain.py:523-526)."""
    f = voxel_vit_features(sd, x, backbone, cell, patch, pos_embedding)
    return F.linear(f, sd["voxel_head.weight"], sd["voxel_head.bias"])


def voxel_vit_forward_images(sd, x, backbone, patch=16):
    """Feature3D_ViT2D_V2.forward_images (vit_3d_2d_pretrain.py:435-451): the 2-D DeiT path through the SAME blocks --
    timm PatchEmbed (Conv2d k = s = 16, flatten(2).transpose(1, 2)), cls token, + pos_embed, blocks, norm, head(x[:, 0])."""
    cfg = BACKBONES[backbone]
    B = x.shape[0]
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1) + sd["pos_embed"]
    f = encoder(sd, t, cfg["depth"], cfg["num_heads"])[:, 0]
    return F.linear(f, sd["head.weight"], sd["head.bias"])


# ----------------------------------------------------------------------------------------------------------------
# point grouping (data/pointnet_util.py) -- integer work in numpy, ordering contract = ascending (distance, index)
# ----------------------------------------------------------------------------------------------------------------
def square_distance_np(src, dst):
    """square_distance (pointnet_util.py:22-36): fp32 ((dx*dx)+(dy*dy))+(dz*dz), src [B,S,3], dst [B,N,3] -> [B,S,N]."""
    src = np.asarray(src, dtype=np.float32)
    dst = np.asarray(dst, dtype=np.float32)
    d = src[:, :, None, :] - dst[:, None, :, :]
    d = d * d
    return (d[..., 0] + d[..., 1]) + d[..., 2]


def knn_np(xyz, query, K):
    """square_distance + argsort()[:, :, :K] (pointnet_util.py:119-120, models/3DViT/model.py:23-24); stable order."""
    out = []
    for b in range(xyz.shape[0]):  # per cloud to bound memory
        d = square_distance_np(query[b:b + 1], xyz[b:b + 1])[0]
        out.append(np.argsort(d, axis=-1, kind="stable")[:, :K])
    return np.stack(out).astype(np.int64)


def ball_query_np(radius, nsample, xyz, query):
    """query_ball_point (pointnet_util.py:76-96)."""
    B, N, _ = xyz.shape
    S = query.shape[1]
    r2 = np.float32(radius ** 2)
    out = np.empty((B, S, nsample), dtype=np.int64)
    for b in range(B):
        d = square_distance_np(query[b:b + 1], xyz[b:b + 1])[0]
        gi = np.broadcast_to(np.arange(N, dtype=np.int64), (S, N)).copy()
        gi[d > r2] = N
        gi = np.sort(gi, axis=-1)[:, :nsample]
        first = np.repeat(gi[:, :1], nsample, axis=1)
        mask = gi == N
        gi[mask] = first[mask]
        out[b] = gi
    return out


def fps_np(xyz, npoint, start):
    """farthest_point_sample (pointnet_util.py:53-73) with the torch.randint start (:65) passed in explicitly."""
    xyz = np.asarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    cent = np.zeros((B, npoint), dtype=np.int64)
    dist = np.full((B, N), 1e10, dtype=np.float32)
    far = np.asarray(start, dtype=np.int64).copy()
    bi = np.arange(B)
    for i in range(npoint):
        cent[:, i] = far
        c = xyz[bi, far][:, None, :]
        d = xyz - c
        d = d * d
        d = (d[..., 0] + d[..., 1]) + d[..., 2]
        dist = np.minimum(dist, d)
        far = np.argmax(dist, axis=-1)  # first maximum, like torch.max(dim) on CPU
    return cent


def index_points(points, idx):
    """index_points (pointnet_util.py:39-50)."""
    raw = idx.shape
    flat = idx.reshape(raw[0], -1)
    res = torch.gather(points, 1, flat[..., None].expand(-1, -1, points.size(-1)))
    return res.reshape(*raw, -1)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


# ----------------------------------------------------------------------------------------------------------------
# PointTransformerCls / PointTransformerSeg (models/3DViT/model.py)
# ----------------------------------------------------------------------------------------------------------------
def _bn(sd, pre, x, training):
    return F.batch_norm(x, sd[pre + "running_mean"].clone(), sd[pre + "running_var"].clone(), sd[pre + "weight"],
                        sd[pre + "bias"], training=training, momentum=0.1, eps=1e-5)


def set_abstraction(sd, pre, xyz, points, npoint, nsample, fps_start, training):
    """PointNetSetAbstraction.forward with knn=True (pointnet_util.py:220-244 -> sample_and_group :99-138).
    The second kNN at :233-235 is dead code (its result is unused) and is not evaluated."""
    B, N, _ = xyz.shape
    fps_idx = _t(fps_np(xyz.detach().numpy(), npoint, fps_start))
    new_xyz = index_points(xyz, fps_idx)
    idx = _t(knn_np(xyz.detach().numpy(), new_xyz.detach().numpy(), nsample))
    grouped = torch.cat([index_points(xyz, idx) - new_xyz.view(B, npoint, 1, 3), index_points(points, idx)], dim=-1)
    h = grouped.permute(0, 3, 2, 1)  # [B, 3+D, K, S]
    for i in range(2):
        h = F.conv2d(h, sd[f"{pre}mlp_convs.{i}.weight"], sd[f"{pre}mlp_convs.{i}.bias"])
        h = F.relu(_bn(sd, f"{pre}mlp_bns.{i}.", h, training))
    return new_xyz, h.max(dim=2)[0].transpose(1, 2)


def three_nn_interpolate(xyz_dst, xyz_src, feats_src):
    """PointNetFeaturePropagation.forward, mlp=[] (pointnet_util.py:381-420): inverse-distance 3-NN interpolation."""
    d = (xyz_dst[:, :, None] - xyz_src[:, None]) ** 2
    d = d.sum(-1)
    dist, idx = d.sort(dim=-1, stable=True)
    dist, idx = dist[:, :, :3], idx[:, :, :3]
    w = 1.0 / (dist + 1e-8)
    w = w / w.sum(dim=2, keepdim=True)
    B, N = xyz_dst.shape[:2]
    return (index_points(feats_src, idx) * w.view(B, N, 3, 1)).sum(dim=2)


def transition_up(sd, pre, xyz1, points1, xyz2, points2, training):
    """TransitionUp.forward (models/3DViT/model.py:42-72)."""
    def fc(p, x):
        h = F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"]).transpose(1, 2)
        return F.relu(_bn(sd, p + "2.", h, training).transpose(1, 2))

    f1 = fc(pre + "fc1.", points1)
    f2 = fc(pre + "fc2.", points2)
    return three_nn_interpolate(xyz2, xyz1, f1) + f2


def _mlp2(sd, pre, x):
    return F.linear(F.relu(F.linear(x, sd[pre + "0.weight"], sd[pre + "0.bias"])), sd[pre + "2.weight"], sd[pre + "2.bias"])


def point_vit_features(sd, x, backbone, num_point, nneighbor, fps_starts, training=False, seg=False):
    """PointTransformerCls.forward_features (models/3DViT/model.py:297-327) / PointTransformerSeg (:494-525)."""
    cfg = BACKBONES[backbone]
    xyz = x[..., :3]
    f = _mlp2(sd, "fc1.", x) + _mlp2(sd, "fc_pos_embed.", xyz)
    xyz0, p0 = set_abstraction(sd, "transition_downs.0.sa.", xyz, f, num_point, nneighbor, fps_starts[0], training)
    xyz1, p1 = set_abstraction(sd, "transition_downs.1.sa.", xyz0, p0, num_point // 4, nneighbor, fps_starts[1], training)
    t = torch.cat((sd["cls_token"].expand(x.shape[0], -1, -1), p1), dim=1)  # no positional embedding is added (:310-319)
    t = encoder(sd, t, cfg["depth"], cfg["num_heads"])[:, 1:]
    t = transition_up(sd, "transition_ups.0.", xyz1, t, xyz0, p0, training)
    t = transition_up(sd, "transition_ups.1.", xyz0, t, xyz, f, training)
    return t if seg else t.mean(1)


def point_vit_logits(sd, x, backbone, num_point, nneighbor, fps_starts, training=False, seg=False):
    f = point_vit_features(sd, x, backbone, num_point, nneighbor, fps_starts, training, seg)
    return F.linear(f, sd["head.weight"], sd["head.bias"])


# ----------------------------------------------------------------------------------------------------------------
# synthetic inputs of SURVEY.md section 8(d) (shared by golden generation, tests and bench)
# ----------------------------------------------------------------------------------------------------------------
def synthetic_voxels(B, V, seed=9, n_classes=40):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(B, 1, V, V, V, generator=g) < 0.1).float()
    y = torch.randint(0, n_classes, (B,), generator=g)
    return x, y


def synthetic_points(B, N, extra=3, seed=9, n_classes=40):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(B, N, 3, generator=g) * 2 - 1
    xyz = xyz / xyz.norm(dim=-1).max(dim=1, keepdim=True)[0][..., None].clamp_min(1e-6)  # inside the unit ball
    feats = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1)
    x = torch.cat([xyz, feats], dim=-1)
    if extra > 3:
        onehot = F.one_hot(torch.randint(0, extra - 3, (B,), generator=g), extra - 3).float()
        x = torch.cat([x, onehot[:, None, :].expand(-1, N, -1)], dim=-1)
    y = torch.randint(0, n_classes, (B,), generator=g)
    return x, y


def init_voxel_state_dict(backbone, cell, patch, n_classes, pos_embedding, seed=9):
    """Random-init weights with the reference's key names and init distributions (timm _init_weights: trunc_normal
    std .02 for Linear, LayerNorm (1, 0); Conv3d keeps torch's default init)."""
    g = torch.Generator().manual_seed(seed)
    cfg = BACKBONES[backbone]
    D, depth = cfg["embed_dim"], cfg["depth"]

    def tn(*shape):
        return (torch.randn(*shape, generator=g) * 0.02).clamp_(-2, 2)

    sd = {"cls_token": tn(1, 1, D), "norm.weight": torch.ones(D), "norm.bias": torch.zeros(D)}
    for i in range(depth):
        p = f"blocks.{i}."
        sd.update({p + "norm1.weight": torch.ones(D), p + "norm1.bias": torch.zeros(D),
                   p + "attn.qkv.weight": tn(3 * D, D), p + "attn.qkv.bias": torch.zeros(3 * D),
                   p + "attn.proj.weight": tn(D, D), p + "attn.proj.bias": torch.zeros(D),
                   p + "norm2.weight": torch.ones(D), p + "norm2.bias": torch.zeros(D),
                   p + "mlp.fc1.weight": tn(4 * D, D), p + "mlp.fc1.bias": torch.zeros(4 * D),
                   p + "mlp.fc2.weight": tn(D, 4 * D), p + "mlp.fc2.bias": torch.zeros(D)})
    k = cell ** 3
    bound = 1.0 / math.sqrt(k)
    sd["voxel_embed.proj.conv3d_1.weight"] = (torch.rand(D, 1, cell, cell, cell, generator=g) * 2 - 1) * bound
    sd["voxel_embed.proj.conv3d_1.bias"] = (torch.rand(D, generator=g) * 2 - 1) * bound
    sd["voxel_head.weight"] = tn(n_classes, D)
    sd["voxel_head.bias"] = torch.zeros(n_classes)
    sd["voxel_pos_embed"] = torch.zeros(1, patch * patch + 1, D)  # stays zero in the reference (Appendix B.2)
    if pos_embedding == "group_embed":
        sd["group_pos_embed"] = torch.zeros(1, patch + 1, D)
        sd["group_cls_token"] = torch.zeros(1, 1, D)
        b = math.sqrt(6.0 / (D + 3 * D))  # xavier_uniform_ on in_proj_weight
        sd["group_embed.self_attn.in_proj_weight"] = (torch.rand(3 * D, D, generator=g) * 2 - 1) * b
        sd["group_embed.self_attn.in_proj_bias"] = torch.zeros(3 * D)
        lb = 1.0 / math.sqrt(D)
        for name, shape in (("self_attn.out_proj", (D, D)), ("linear1", (D, D)), ("linear2", (D, D))):
            sd[f"group_embed.{name}.weight"] = (torch.rand(*shape, generator=g) * 2 - 1) * lb
            sd[f"group_embed.{name}.bias"] = (torch.rand(shape[0], generator=g) * 2 - 1) * lb
        sd["group_embed.self_attn.out_proj.bias"] = torch.zeros(D)
        for n in ("norm1", "norm2"):
            sd[f"group_embed.{n}.weight"] = torch.ones(D)
            sd[f"group_embed.{n}.bias"] = torch.zeros(D)
    return sd


def init_image_branch_state_dict(backbone, seed=19, img=224, patch=16, n_classes=1000):
    """The 2-D image branch of Feature3D_ViT2D_V2 (pos_embed, patch_embed.proj, head: timm VisionTransformer members,
    Appendix A) -- a separate generator so the voxel fixtures' weight stream is unchanged."""
    g = torch.Generator().manual_seed(seed)
    D = BACKBONES[backbone]["embed_dim"]
    k = 3 * patch * patch
    return {"pos_embed": (torch.randn(1, (img // patch) ** 2 + 1, D, generator=g) * 0.02).clamp_(-2, 2),
            "patch_embed.proj.weight": (torch.rand(D, 3, patch, patch, generator=g) * 2 - 1) / math.sqrt(k),
            "patch_embed.proj.bias": (torch.rand(D, generator=g) * 2 - 1) / math.sqrt(k),
            "head.weight": (torch.randn(n_classes, D, generator=g) * 0.02).clamp_(-2, 2),
            "head.bias": torch.zeros(n_classes)}


def sharpen_point_state_dict(sd, qkv_gain=4.0, head_gain=12.0):
    """Fixture variant with non-trivial logits and attention: at the reference's init the point models have |logit| < 0.08
    and near-uniform attention (trunc_normal std .02 everywhere), which makes an absolute 1e-2 tolerance toothless.
    Scaling the qkv weights (scores x gain^2) and the head puts |logit| ~ 1 and spreads the softmax."""
    out = dict(sd)
    for k in sd:
        if k.endswith("attn.qkv.weight"):
            out[k] = sd[k] * qkv_gain
    out["head.weight"] = sd["head.weight"] * head_gain
    return out


def init_point_state_dict(backbone, input_dim, n_classes, seed=9):
    """Random-init weights for PointTransformerCls/Seg with the reference's key names (models/3DViT/model.py:199-263).
    Only keys used by forward() are produced (PointEmbed, pos_embed and every last_pos_embed are dead, Appendix B.10)."""
    g = torch.Generator().manual_seed(seed)
    cfg = BACKBONES[backbone]
    D, depth = cfg["embed_dim"], cfg["depth"]

    def tn(*shape):
        return (torch.randn(*shape, generator=g) * 0.02).clamp_(-2, 2)

    def uni(shape, fan_in):
        return (torch.rand(*shape, generator=g) * 2 - 1) / math.sqrt(fan_in)

    sd = {"cls_token": tn(1, 1, D), "norm.weight": torch.ones(D), "norm.bias": torch.zeros(D)}
    for i in range(depth):
        p = f"blocks.{i}."
        sd.update({p + "norm1.weight": torch.ones(D), p + "norm1.bias": torch.zeros(D),
                   p + "attn.qkv.weight": tn(3 * D, D), p + "attn.qkv.bias": torch.zeros(3 * D),
                   p + "attn.proj.weight": tn(D, D), p + "attn.proj.bias": torch.zeros(D),
                   p + "norm2.weight": torch.ones(D), p + "norm2.bias": torch.zeros(D),
                   p + "mlp.fc1.weight": tn(4 * D, D), p + "mlp.fc1.bias": torch.zeros(4 * D),
                   p + "mlp.fc2.weight": tn(D, 4 * D), p + "mlp.fc2.bias": torch.zeros(D)})
    q = D // 4
    for name, din in (("fc1", input_dim), ("fc_pos_embed", 3)):
        sd[f"{name}.0.weight"] = uni((q, din), din)
        sd[f"{name}.0.bias"] = uni((q,), din)
        sd[f"{name}.2.weight"] = uni((q, q), q)
        sd[f"{name}.2.bias"] = uni((q,), q)

    def bn(pre, c):
        sd[pre + "weight"] = torch.rand(c, generator=g) * 0.5 + 0.75
        sd[pre + "bias"] = (torch.rand(c, generator=g) - 0.5) * 0.2
        sd[pre + "running_mean"] = (torch.rand(c, generator=g) - 0.5) * 0.1
        sd[pre + "running_var"] = torch.rand(c, generator=g) * 0.5 + 0.75

    for i in range(2):  # TransitionDown(k, nneighbor, [channel//2 + 3, channel, channel])
        ch = q * 2 ** (i + 1)
        last = ch // 2 + 3
        for j in range(2):
            sd[f"transition_downs.{i}.sa.mlp_convs.{j}.weight"] = uni((ch, last, 1, 1), last)
            sd[f"transition_downs.{i}.sa.mlp_convs.{j}.bias"] = uni((ch,), last)
            bn(f"transition_downs.{i}.sa.mlp_bns.{j}.", ch)
            last = ch
    for n, i in enumerate(reversed(range(2))):  # TransitionUp(channel * 2, channel, channel)
        ch = q * 2 ** i
        for fc, din in (("fc1", ch * 2), ("fc2", ch)):
            sd[f"transition_ups.{n}.{fc}.0.weight"] = uni((ch, din), din)
            sd[f"transition_ups.{n}.{fc}.0.bias"] = uni((ch,), din)
            bn(f"transition_ups.{n}.{fc}.2.", ch)
    sd["head.weight"] = tn(n_classes, q)
    sd["head.bias"] = torch.zeros(n_classes)
    return sd


def state_dict_checksum(sd):
    """Order-independent fingerprint used to prove both sides regenerated identical weights from the seed."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float(v.sum()) + float((v * v).sum()) * 1e-3 + float(v.flatten()[:: max(1, v.numel() // 7)].sum()) * 1e-2
    return tot
