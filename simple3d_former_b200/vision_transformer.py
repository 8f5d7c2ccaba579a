"""Drop-in, timm==0.3.2-compatible encoder modules backed by the sm_100a kernels.

Mirrors the module surface the reference relies on (timm 0.3.2 `timm/models/vision_transformer.py`,
imported by the reference at models/vit_3d_2d_pretrain.py:8-10, models/3DViT/model.py:6-8): constructor signatures,
attribute names (`blocks[i].attn.qkv`, `.num_heads`, `.scale`, `.norm1`, `.mlp.fc1`, ...) and state-dict keys are
identical, so DeiT checkpoints load and the reference's `forward_features()` loops (`for blk in self.blocks: x = blk(x)`)
run unchanged. Parameters stay fp32 `nn.Parameter`s; bf16 shadows feed the tensor cores.

There is no CPU path: calling these modules on CPU tensors raises.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from . import functional as Fn

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def _cfg(url='', **kwargs):
    out = {'url': url, 'num_classes': 1000, 'input_size': (3, 224, 224), 'pool_size': None, 'crop_pct': .9,
           'interpolation': 'bicubic', 'mean': IMAGENET_DEFAULT_MEAN, 'std': IMAGENET_DEFAULT_STD,
           'first_conv': 'patch_embed.proj', 'classifier': 'head'}
    out.update(kwargs)
    return out


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    """Stochastic depth per sample (identity at rate 0, which is all the reference uses)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.drop_prob or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x / keep * mask


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        if not isinstance(self.act, nn.GELU):
            raise NotImplementedError("only the exact-erf GELU of timm 0.3.2 is implemented in the fused MLP")

    def forward(self, x):
        if self.drop.p > 0 and self.training:
            raise NotImplementedError("Mlp dropout > 0 is not used by the reference and not implemented")
        return Fn.MlpFn.apply(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        if self.training and (self.attn_drop.p > 0 or self.proj_drop.p > 0):
            raise NotImplementedError("attention dropout > 0 is not used by the reference and not implemented")
        return Fn.AttentionFn.apply(x, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias, self.num_heads,
                                    self.scale)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def _fusable(self):
        if not (isinstance(self.norm1, nn.LayerNorm) and isinstance(self.norm2, nn.LayerNorm)):
            return False
        if self.training and (self.attn.attn_drop.p > 0 or self.attn.proj_drop.p > 0 or self.mlp.drop.p > 0):
            # every reference backbone uses rate 0; never train silently without the dropout
            raise NotImplementedError("Block with attn_drop / proj_drop / mlp drop > 0 is not implemented")
        if isinstance(self.drop_path, DropPath) and self.training and self.drop_path.drop_prob:
            return False
        # forward hooks on sub-modules (e.g. the reference's attention visualiser hooks blocks[i].attn) need the
        # module-by-module path so the hooks fire
        for m in (self.attn, self.mlp, self.norm1, self.norm2, self.attn.qkv):
            if m._forward_hooks or m._forward_pre_hooks:
                return False
        return True

    def forward(self, x):
        if self._fusable():
            a, m = self.attn, self.mlp
            return Fn.BlockFn.apply(x, self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight,
                                    a.proj.bias, self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias,
                                    m.fc2.weight, m.fc2.bias, a.num_heads, a.scale, self.norm1.eps, self.norm2.eps)
        x = x + self.drop_path(self.attn(self.norm1(x)))
        x = x + self.drop_path(self.mlp(self.norm2(x)))
        return x


class FusedLayerNorm(nn.LayerNorm):
    """nn.LayerNorm with the one-pass sm_100a kernel on CUDA inputs (same parameters / state-dict keys)."""

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float32 and self.elementwise_affine and len(self.normalized_shape) == 1:
            return Fn.LayerNormFn.apply(x, self.weight, self.bias, self.eps)
        return super().forward(x)


class PatchEmbed(nn.Module):
    """2-D image to patch embedding (only reached through forward_images, the LwF side path)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., hybrid_backbone=None, norm_layer=nn.LayerNorm):
        super().__init__()
        if hybrid_backbone is not None:
            raise NotImplementedError("hybrid CNN backbones are outside the hot path")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        rates = torch.linspace(0, drop_path_rate, depth).tolist()
        self.blocks = nn.ModuleList(
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=rates[i], norm_layer=norm_layer)
            for i in range(depth))
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def forward_features(self, x):
        B = x.shape[0]
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        x = self.pos_drop(x + self.pos_embed)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)[:, 0]

    def forward(self, x):
        f = self.forward_features(x)
        if isinstance(self.head, nn.Linear):  # fp32 head on our own CUDA-core GEMM (no cuBLAS on the path)
            return Fn.LinearF32Fn.apply(f, self.head.weight, self.head.bias)
        return self.head(f)


# norm_layer used by every backbone of the reference (vit_3d_2d_pretrain.py:287): LayerNorm(eps=1e-6), fused kernel
default_norm_layer = partial(FusedLayerNorm, eps=1e-6)
