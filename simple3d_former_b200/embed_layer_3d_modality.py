"""Voxel tokenizers with the reference's module surface (models/embed_layer_3d_modality.py:10-70).

`VoxelEmbed` = Conv3d(1 -> D, k = s = cell) followed by a mean over the last spatial axis; `VoxelEmbed_no_average` = the
conv only. Same constructor, attributes (`voxel_size`, `cell_size`, `patch_size`, `num_patches`, `embed_dim`) and
state-dict key (`proj.conv3d_1.{weight,bias}`). The conv is evaluated as patch-gather + tcgen05 GEMM and produced
token-major; `forward()` returns the reference's channel-first shape as a zero-copy view, `forward_tokens()` returns
[B, tokens, D] directly (what the model wrappers use, killing the flatten/transpose/rearrange copies at
vit_3d_2d_pretrain.py:457,474).
"""
from collections import OrderedDict

import torch
from torch import nn

from . import functional as Fn


class _VoxelEmbedBase(nn.Module):
    _zmean = True

    def __init__(self, voxel_size=128, cell_size=16, patch_size=8, in_chans=1, embed_dim=768):
        super().__init__()
        if in_chans != 1:
            raise NotImplementedError("the reference only ever uses in_chans=1 occupancy grids")
        self.voxel_size = (voxel_size, voxel_size, voxel_size)
        self.cell_size = (cell_size, cell_size, cell_size)
        self.patch_size = patch_size
        self.num_patches = patch_size ** 2 if self._zmean else patch_size ** 3
        self.embed_dim = embed_dim
        self.proj = torch.nn.Sequential(OrderedDict([
            ('conv3d_1', torch.nn.Conv3d(in_channels=in_chans, out_channels=embed_dim, kernel_size=cell_size,
                                         stride=cell_size)),
        ]))
        # the conv output grid (floor division: trailing voxels are ignored, e.g. 128 -> 14 * 9 = 126)
        self._grid = (voxel_size - cell_size) // cell_size + 1

    def _check(self, x):
        B, C, H, W, V = x.shape
        assert H == self.voxel_size[0] and W == self.voxel_size[1] and V == self.voxel_size[2], \
            f"Input voxel size ({H}*{W}*{V}) doesn't match model ({self.voxel_size[0]}*{self.voxel_size[1]}*{self.voxel_size[2]})."

    def forward_tokens(self, x):
        """[B,1,V,V,V] -> [B, p*p(*p), D] fp32, token order (px, py[, pz])."""
        self._check(x)
        conv = self.proj.conv3d_1
        return Fn.VoxelPatchifyFn.apply(x, conv.weight, conv.bias, self.cell_size[0], self._grid, self._zmean)


class VoxelEmbed(_VoxelEmbedBase):
    """Voxel to patch embedding, mean over the z patches (embed_layer_3d_modality.py:10-40)."""
    _zmean = True

    def forward(self, x):
        t = self.forward_tokens(x)
        g = self._grid
        return t.view(x.shape[0], g, g, self.embed_dim).permute(0, 3, 1, 2)  # [B, D, p, p]


class VoxelEmbed_no_average(_VoxelEmbedBase):
    """Voxel to patch embedding without the z mean (embed_layer_3d_modality.py:42-70)."""
    _zmean = False

    def forward(self, x):
        t = self.forward_tokens(x)
        g = self._grid
        return t.view(x.shape[0], g, g, g, self.embed_dim).permute(0, 4, 1, 2, 3)  # [B, D, p, p, p]
