"""Model wrappers of the reference ("callers" of the hot path) rebuilt on the fused
modules: same constructor arguments, attributes and state-dict keys as

  * models/vit_3d_2d_pretrain.py::Feature3D_ViT2D_V2  (:275-526)   -- voxel classification, 'default' / 'group_embed'
  * models/3DViT/model.py::PointTransformerCls (:144-337) / ::PointTransformerSeg (:341-535)

so reference checkpoints load unchanged and train_cls_voxel.py / train_cls.py / train_partseg.py can construct them
in place of the originals. Reference quirks are kept on purpose: deit_base uses 3 heads,
`voxel_pos_embed` / `group_pos_embed` / `group_cls_token` start at zero, the 12 blocks run twice in group-embed mode,
no positional embedding is added to point tokens, PointEmbed / pos_embed / last_pos_embed are constructed but unused.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from . import functional as Fn
from .pointnet_util import PointNetFeaturePropagation, PointNetSetAbstraction, knn_point
from .vision_transformer import FusedLayerNorm, VisionTransformer, _cfg, trunc_normal_

_norm = partial(FusedLayerNorm, eps=1e-6)

BACKBONES = {
    'deit_tiny_patch16_224': dict(patch_size=16, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4, qkv_bias=True, norm_layer=_norm),
    'deit_small_patch16_224': dict(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True, norm_layer=_norm),
    'deit_base_patch16_224': dict(patch_size=16, embed_dim=768, depth=12, num_heads=3, mlp_ratio=4, qkv_bias=True, norm_layer=_norm),
    'deit_base_distilled_patch16_224': dict(patch_size=16, embed_dim=768, depth=12, num_heads=3, mlp_ratio=4, qkv_bias=True, norm_layer=_norm),
    'vit_base_patch16_224_21k': dict(patch_size=16, embed_dim=768, depth=12, num_heads=3, mlp_ratio=4, qkv_bias=True, norm_layer=_norm),
}
PRETRAINED_URLS = {
    'deit_tiny_patch16_224': "https://dl.fbaipublicfiles.com/deit/deit_tiny_patch16_224-a1311bcf.pth",
    'deit_small_patch16_224': "https://dl.fbaipublicfiles.com/deit/deit_small_patch16_224-cd65a155.pth",
    'deit_base_patch16_224': "https://dl.fbaipublicfiles.com/deit/deit_base_patch16_224-b5f2ef4d.pth",
    'deit_base_distilled_patch16_224': "https://dl.fbaipublicfiles.com/deit/deit_base_distilled_patch16_224-df68dfff.pth",
    # a LOCAL file in the reference (jax -> PyTorch conversion of ViT-B/16 21k), vit_3d_2d_pretrain.py:331, models/3DViT/model.py:196
    'vit_base_patch16_224_21k': "./3rd_party/ViT-PyTorch/jax_to_pytorch/weights/B_16.pth",
}
# options the reference's scripts expose that are NOT on the hot path named by north_star (INTEGRATION.md lists them)
UNSUPPORTED_POS_EMBEDDINGS = ("no_embed", "weight_sharing")


class _SelfAttnParams(nn.Module):
    """Parameter container with nn.MultiheadAttention's names (in_proj_weight, in_proj_bias, out_proj.{weight,bias})."""

    def __init__(self, embed_dim, num_heads):
        super().__init__()
        self.embed_dim, self.num_heads, self.batch_first = embed_dim, num_heads, False
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.)


class GroupEmbedLayer(nn.Module):
    """nn.TransformerEncoderLayer(d_model, nhead, dim_feedforward) as used for `group_embed`
    (vit_3d_2d_pretrain.py:381): post-norm, ReLU, sequence-first [S, Nb, E] -- attention runs over S = B*px*py.
    State-dict keys match torch's layer. Dropout (p = 0.1 as in the reference) is active in train() at the four sites of
    torch's layer (attention probabilities, dropout1, FFN dropout, dropout2) with counter-based masks keyed by a
    device-resident seed that advances every training forward (so CUDA-graph replays draw fresh masks); eval() and
    p = 0 run the deterministic arithmetic that parity is defined on."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, layer_norm_eps=1e-5):
        super().__init__()
        self.self_attn = _SelfAttnParams(d_model, nhead)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.dropout_p = dropout
        self.nhead = nhead
        # dropout seed lives on the device (not in the state dict): bumped in place before every training forward
        # (a fixed start value: drawing it from torch's generator would shift the reference's weight-init stream)
        self.register_buffer("_drop_seed", torch.full((1,), 20210914, dtype=torch.int32), persistent=False)

    def forward(self, src):
        a = self.self_attn
        drop = self.training and self.dropout_p > 0.0
        if drop:
            with torch.no_grad():
                self._drop_seed += 1
        return Fn.GroupEmbedFn.apply(src, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias,
                                     self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                                     self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias, self.nhead,
                                     self.norm1.eps, self.dropout_p if drop else 0.0, self._drop_seed if drop else None)


def remap_vit21k_keys(ckpt, depth=12):
    """Key layout of the jax->PyTorch ViT-B/16 21k checkpoint -> timm 0.3.2 names (what the reference's `fit_dict`,
    vit_3d_2d_pretrain.py:16-36, does): drop the `transformer.` prefix, `pwff` is `mlp`, and the separate
    `attn.proj_{q,k,v}` Linear layers are stacked row-wise into timm's single `attn.qkv` ([q | k | v], Appendix A)."""
    out = {}
    for k, v in ckpt.items():
        k = k.replace("pwff", "mlp")
        out[k[len("transformer."):] if k.startswith("transformer.") else k] = v
    for i in range(depth):
        for kind in ("weight", "bias"):
            parts = [out.pop(f"blocks.{i}.attn.proj_{c}.{kind}") for c in "qkv"]
            out[f"blocks.{i}.attn.qkv.{kind}"] = torch.cat(parts, dim=0)
    return out


def _load_pretrained(model, url, distilled=False):
    if url is None:
        raise ValueError(f"no pretrained weights are known for backbone {model.transformer_backbone!r}")
    if "21k" in model.transformer_backbone:  # local file, as in the reference (vit_3d_2d_pretrain.py:400-403)
        source = remap_vit21k_keys(torch.load(url, map_location="cpu"))
    else:
        source = torch.hub.load_state_dict_from_url(url=url, map_location="cpu", check_hash=True)["model"]
    pretrained = {k: v for k, v in source.items() if k in model.state_dict()}
    if distilled:
        pretrained['pos_embed'] = pretrained['pos_embed'][:, 1:, :]
    model.load_state_dict(pretrained, strict=False)


class Feature3D_ViT2D_V2(VisionTransformer):
    def __init__(self, n_classes=10, embed_layer=None, data_shape=None, transformer_backbone='deit_base_patch16_224',
                 pretrained=True, pos_embedding=None, **kwargs):
        if transformer_backbone not in BACKBONES:
            raise ValueError("Unknown transformer backbone name!")
        self.transformer_backbone = transformer_backbone
        self.pretrained = pretrained
        super().__init__(**BACKBONES[transformer_backbone])
        self.default_cfg = _cfg()
        self.url = PRETRAINED_URLS.get(transformer_backbone)
        self.dist_token = None
        self.n_classes = n_classes
        if pretrained:
            _load_pretrained(self, self.url, 'distilled' in transformer_backbone)
            self.freeze_image_branch()
        self.voxel_embed = embed_layer
        if kwargs.get('head') == 'AMSoftmax':
            raise NotImplementedError("head='AMSoftmax' (vit_3d_2d_pretrain.py:39-74) is outside the hot path of this "
                                      "implementation; use the default Linear head")
        self.voxel_head = nn.Linear(self.embed_dim, self.n_classes)
        self.pos_embed_type = pos_embedding
        if pos_embedding is None or pos_embedding == "default":
            self.voxel_pos_embed = nn.Parameter(torch.zeros(1, self.voxel_embed.num_patches + 1, self.embed_dim))
            trunc_normal_(self.pos_embed, std=.02)  # sic: the reference initialises pos_embed, not voxel_pos_embed
        elif pos_embedding == "group_embed":
            self.voxel_pos_embed = nn.Parameter(torch.zeros(1, self.voxel_embed.patch_size ** 2 + 1, self.embed_dim))
            trunc_normal_(self.pos_embed, std=.02)
            self.group_embed = GroupEmbedLayer(d_model=self.embed_dim, dim_feedforward=self.embed_dim, nhead=4)
            self.group_pos_embed = nn.Parameter(torch.zeros(1, self.voxel_embed.patch_size + 1, self.embed_dim))
            self.group_cls_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
        elif pos_embedding in UNSUPPORTED_POS_EMBEDDINGS:
            raise NotImplementedError(f"pos_embedding={pos_embedding!r} is a reference option outside the hot path of "
                                      "this implementation (supported: None / 'default', 'group_embed')")
        else:
            raise ValueError("Unknown positional embedding scheme!")  # the reference's message (vit_3d_2d_pretrain.py:389)

    def freeze_image_branch(self):
        """head / pos_embed / patch_embed only serve forward_images; frozen as on the reference's pretrained path
        (vit_3d_2d_pretrain.py:428-432) so data-parallel training sees no unused trainable parameters."""
        self.head.weight.requires_grad = False
        self.head.bias.requires_grad = False
        self.pos_embed.requires_grad = False
        for p in self.patch_embed.parameters():
            p.requires_grad = False

    def forward_images(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1)
        x = self.pos_drop(x + self.pos_embed)
        for blk in self.blocks:
            x = blk(x)
        return Fn.LinearF32Fn.apply(self.norm(x[:, 0]), self.head.weight, self.head.bias)

    def _tokens(self, x):
        emb = self.voxel_embed
        if hasattr(emb, "forward_tokens"):
            return emb.forward_tokens(x)  # [B, T, D] token-major straight from the GEMM epilogue
        y = emb(x)  # foreign embed layer: channel-first like the reference
        return y.flatten(2).transpose(1, 2)

    def forward_features(self, x):
        B = x.shape[0]
        if self.pos_embed_type in (None, "default"):
            t = self._tokens(x)
            t = torch.cat((self.cls_token.expand(B, -1, -1), t), dim=1)
            t = self.pos_drop(t + self.voxel_pos_embed)
            for blk in self.blocks:
                t = blk(t)
            return self.norm(t[:, 0])  # LayerNorm is per token: identical to the reference's norm(t)[:, 0] (:469-470)
        p = self.voxel_embed.patch_size
        D = self.embed_dim
        t = self._tokens(x).reshape(B * p * p, p, D)  # '(b px py) pz c'
        t = torch.cat((self.group_cls_token.expand(t.shape[0], -1, -1), t), dim=1)
        t = self.pos_drop(t + self.group_pos_embed)
        t = self.group_embed(t)
        for blk in self.blocks:
            t = blk(t)
        t = self.norm(t[:, 0]).reshape(B, p * p, D)  # = norm(t)[:, 0] (:484-485) without normalising the 14 discarded rows
        t = torch.cat((self.cls_token.expand(B, -1, -1), t), dim=1)
        t = self.pos_drop(t + self.voxel_pos_embed)
        for blk in self.blocks:
            t = blk(t)
        return self.norm(t[:, 0])

    def forward(self, x):
        return Fn.LinearF32Fn.apply(self.forward_features(x), self.voxel_head.weight, self.voxel_head.bias)


# ----------------------------------------------------------------------------------------------------------------
# point models
# ----------------------------------------------------------------------------------------------------------------
class TransitionDown(nn.Module):
    def __init__(self, k, nneighbor, channels):
        super().__init__()
        self.sa = PointNetSetAbstraction(k, 0, nneighbor, channels[0], channels[1:], group_all=False, knn=True)

    def forward(self, xyz, points):
        return self.sa(xyz, points)


class _SwapAxes(nn.Module):
    def forward(self, x):
        return x.transpose(1, 2)


class TransitionUp(nn.Module):
    def __init__(self, dim1, dim2, dim_out):
        super().__init__()
        self.fc1 = nn.Sequential(nn.Linear(dim1, dim_out), _SwapAxes(), nn.BatchNorm1d(dim_out), _SwapAxes(), nn.ReLU())
        self.fc2 = nn.Sequential(nn.Linear(dim2, dim_out), _SwapAxes(), nn.BatchNorm1d(dim_out), _SwapAxes(), nn.ReLU())
        self.fp = PointNetFeaturePropagation(-1, [])

    @staticmethod
    def _fc(seq, x):
        """Linear -> BatchNorm1d -> ReLU of the reference's Sequential as one fused node (GEMM + two HBM passes)."""
        lin, bn = seq[0], seq[2]
        y = Fn.LinearBnReluFn.apply(x, lin.weight, lin.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                    seq.training, bn.eps, bn.momentum)
        if seq.training:
            with torch.no_grad():
                bn.num_batches_tracked += 1
        return y

    def forward(self, xyz1, points1, xyz2, points2):
        """xyz1 [B,S,3] / points1 [B,S,dim1]: coarse level; xyz2 [B,N,3] / points2 [B,N,dim2]: fine level.
        There is no PyTorch-op path: shapes the kernels do not cover raise (every width the reference's backbones
        produce -- multiples of 8 channels, more than one coarse point -- is covered)."""
        if not points1.is_cuda:
            raise RuntimeError("simple3d_former_b200 ops need CUDA tensors (no CPU fallback)")
        if xyz1.shape[1] < 3 or self.fc1[0].out_features % 8 or points1.shape[-1] % 8 or points2.shape[-1] % 8:
            raise NotImplementedError(
                "TransitionUp: the sm_100a path needs >= 3 coarse points and channel counts that are multiples of 8 "
                f"(got S={xyz1.shape[1]}, dims {points1.shape[-1]}/{points2.shape[-1]}/{self.fc1[0].out_features})")
        feats1 = self._fc(self.fc1, points1)
        feats2 = self._fc(self.fc2, points2)
        idx, dist = knn_point(3, xyz1.detach().contiguous(), xyz2.detach().contiguous(), return_dist=True)
        return Fn.ThreeNNInterpFn.apply(feats1, idx, dist, feats2)


class _LocalOpParams(nn.Module):
    """Parameter container with the names / shapes of the reference's Local_op (models/3DViT/model.py:75-94)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Conv1d(in_channels, out_channels, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(out_channels, out_channels, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm1d(out_channels)
        self.bn2 = nn.BatchNorm1d(out_channels)


class PointEmbed(nn.Module):
    """The reference's PointEmbed (models/3DViT/model.py:96-121) is built in __init__ (:227) and NEVER called in
    forward(), but its weights are part of every reference checkpoint
    (`patch_embed.{conv1,conv2,bn1,bn2,gather_local_0,gather_local_1}.*`) and train_cls.py:75 / train_partseg.py:80 load
    them with strict=True. This keeps exactly those parameters and buffers (same names, shapes and init order) so
    state dicts round-trip in both directions; there is no forward. The parameters never receive gradients and are
    listed by `unused_parameter_names()` so the data-parallel trainer leaves them out of its buckets."""

    def __init__(self, cfg):
        super().__init__()
        self.conv1 = nn.Conv1d(cfg.input_dim, 64, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(64, 64, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(64)
        self.gather_local_0 = _LocalOpParams(128, cfg.embed_dim // 4)
        self.gather_local_1 = _LocalOpParams(256, cfg.embed_dim // 4)

    def forward(self, x):
        raise NotImplementedError("PointEmbed is dead code in the reference's forward (models/3DViT/model.py:310-319); "
                                  "only its parameters are kept for checkpoint compatibility")


class _PointTransformerBase(VisionTransformer):
    _seg = False

    def __init__(self, cfg):
        npoints, nneighbor, n_c, d_points = cfg.num_point, cfg.model.nneighbor, cfg.num_class, cfg.input_dim
        self.transformer_backbone = cfg.model.transformer_backbone
        self.pretrained = cfg.model.pretrained
        if self.transformer_backbone not in BACKBONES:
            raise ValueError("Unknown transformer backbone name!")
        super().__init__(**BACKBONES[self.transformer_backbone])
        self.default_cfg = _cfg()
        self.url = PRETRAINED_URLS.get(self.transformer_backbone)
        self.dist_token = None
        self.n_classes = n_c
        cfg.embed_dim = self.embed_dim
        if self.pretrained:
            _load_pretrained(self, self.url)
        self.patch_embed = PointEmbed(cfg)
        if cfg.model.head == 'AMSoftmax':
            raise NotImplementedError("AMSoftmax head is outside the hot path")
        q = self.embed_dim // 4
        self.head = nn.Linear(q, self.n_classes)
        self.pos_embed_type = 'default'
        self.transition_downs = nn.ModuleList()
        for i in range(2):
            ch = q * 2 ** (i + 1)
            self.transition_downs.append(TransitionDown(npoints // 4 ** i, nneighbor, [ch // 2 + 3, ch, ch]))
        self.transition_ups = nn.ModuleList()
        for i in reversed(range(2)):
            ch = q * 2 ** i
            self.transition_ups.append(TransitionUp(ch * 2, ch, ch))
        self.fc1 = nn.Sequential(nn.Linear(d_points, q), nn.ReLU(), nn.Linear(q, q))
        self.fc_pos_embed = nn.Sequential(nn.Linear(3, q), nn.ReLU(), nn.Linear(q, q))

    def set_fps_starts(self, starts):
        """Explicit FPS start indices for the two transition-down stages (None restores random starts)."""
        for td, s in zip(self.transition_downs, starts or (None, None)):
            td.sa.fps_start = s

    def unused_parameter_names(self):
        """Parameters that never receive a gradient (kept only for checkpoint compatibility)."""
        return [n for n, _ in self.named_parameters()
                if n == "pos_embed" or "last_pos_embed" in n or n.startswith("patch_embed.")]

    def forward_features(self, x):
        xyz = x[..., :3].contiguous()
        f = self.pos_drop(Fn.PointStemFn.apply(x, xyz, self.fc1[0].weight, self.fc1[0].bias, self.fc1[2].weight,
                                               self.fc1[2].bias, self.fc_pos_embed[0].weight, self.fc_pos_embed[0].bias,
                                               self.fc_pos_embed[2].weight, self.fc_pos_embed[2].bias))
        xyz_0, points_0 = self.transition_downs[0](xyz, f)
        xyz_1, points_1 = self.transition_downs[1](xyz_0, points_0)
        t = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), points_1), dim=1)
        for blk in self.blocks:
            t = blk(t)
        t = self.norm(t)[:, 1:]
        t = self.transition_ups[0](xyz_1, t, xyz_0, points_0)
        t = self.transition_ups[1](xyz_0, t, xyz, f)
        return t if self._seg else t.mean(1)

    def forward(self, x):
        return Fn.LinearF32Fn.apply(self.forward_features(x), self.head.weight, self.head.bias)


class PointTransformerCls(_PointTransformerBase):
    _seg = False


class PointTransformerSeg(_PointTransformerBase):
    _seg = True
