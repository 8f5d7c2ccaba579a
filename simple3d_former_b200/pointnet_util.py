"""Point-grouping utilities with the reference's function surface (data/pointnet_util.py), backed by sm_100a kernels.

`square_distance + argsort()[:, :, :K]` (:119-120, :233-234) becomes `knn_point` (no [B,S,N] matrix, no full sort),
`query_ball_point` (:76-96), `farthest_point_sample` (:53-73) and `index_points` (:39-50) are single kernels; integer
outputs are bit-exact with the reference on tie-free inputs (tie contract: ascending (distance, index)).
`PointNetSetAbstraction` with a two-layer MLP (the only shape the 3DViT models build) runs as
`functional.SetAbstractionFn`: the grouped tensor [B,S,K,3+C] is never materialised, layer 1 is evaluated per point on
the tensor cores and gathered, BatchNorm statistics / ReLU / the max over neighbours are fused passes
(csrc/pointnet_fused.cu). Configurations without a kernel (other layer counts, group_all, widths that are not multiples
of 8) raise NotImplementedError -- there is no PyTorch-op path. The reference's dead second kNN (:233-235) and its torch.cuda.empty_cache() stalls (:115-127) are
dropped.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import functional as Fn


def pc_normalize(pc):
    centroid = np.mean(pc, axis=0)
    pc = pc - centroid
    return pc / np.max(np.sqrt(np.sum(pc ** 2, axis=1)))


def square_distance(src, dst):
    """[B,N,3] x [B,M,3] -> [B,N,M]. Kept for API compatibility only; nothing on the hot path materialises it."""
    return torch.sum((src[:, :, None] - dst[:, None]) ** 2, dim=-1)


def index_points(points, idx):
    """points [B,N,C], idx [B,S(,K)] -> [B,S(,K),C]"""
    return Fn.GatherRowsFn.apply(points, idx)


def farthest_point_sample(xyz, npoint, start=None):
    """xyz [B,N,3] -> centroid indices [B,npoint] (int64). `start` makes the reference's torch.randint draw explicit."""
    B, N, _ = xyz.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.long, device=xyz.device)
    return L.fps(xyz.detach(), npoint, start.to(xyz.device))


def knn_point(nsample, xyz, new_xyz, return_dist=False):
    """Indices of the `nsample` nearest points of xyz [B,N,3] for every query new_xyz [B,S,3] -> int64 [B,S,nsample]."""
    return L.knn(xyz.detach(), new_xyz.detach(), nsample, want_dist=return_dist)


def query_ball_point(radius, nsample, xyz, new_xyz):
    r2 = float(np.float32(radius ** 2))
    return L.ball_query(r2, nsample, xyz.detach(), new_xyz.detach())


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, knn=False, fps_start=None):
    B, N, C = xyz.shape
    fps_idx = farthest_point_sample(xyz, npoint, fps_start)
    new_xyz = index_points(xyz, fps_idx)
    idx = knn_point(nsample, xyz, new_xyz) if knn else query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, npoint, 1, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz_norm, index_points(points, idx)], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device)
    grouped_xyz = xyz.view(B, 1, N, C)
    new_points = torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1) if points is not None else grouped_xyz
    return new_xyz, new_points


class PointNetSetAbstraction(nn.Module):
    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all, knn=False):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.knn, self.group_all = npoint, radius, nsample, knn, group_all
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        self.pos_embeds = nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last = out_channel
        # constructed (never used in forward) by the reference; kept so checkpoints load with strict=True
        self.last_pos_embed = nn.Sequential(nn.Linear(3, last), nn.ReLU(), nn.Linear(last, last))
        self.fps_start = None  # optional explicit FPS start indices (parity tests)

    def _unsupported(self, xyz, points):
        """Reason why this configuration has no sm_100a path (None when it has). The 3DViT models only ever build the
        two-layer, kNN / ball-query grouped form with 8-aligned widths (models/3DViT/model.py:33-44)."""
        if not xyz.is_cuda:
            return "CUDA tensors required (no CPU fallback)"
        if self.group_all:
            return "group_all=True is not used by the 3DViT models and has no kernel"
        if points is None:
            return "points=None (xyz-only grouping) has no kernel"
        if len(self.mlp_convs) != 2:
            return f"{len(self.mlp_convs)}-layer MLP (the fused set abstraction implements the two-layer form)"
        c1, c2 = self.mlp_convs[0].out_channels, self.mlp_convs[1].out_channels
        if not all(bn.track_running_stats and bn.affine and bn.momentum is not None for bn in self.mlp_bns):
            return "BatchNorm2d without running statistics / affine parameters / momentum"
        if points.shape[-1] % 8 or c1 % 8 or c2 % 8:
            return f"channel counts must be multiples of 8 (got {points.shape[-1]}, {c1}, {c2})"
        if self.nsample > 32:
            return f"nsample {self.nsample} > 32"
        return None

    def forward(self, xyz, points):
        """xyz [B,N,3], points [B,N,C] -> (new_xyz [B,S,3], new_points [B,S,C']). No PyTorch-op path: unsupported
        configurations raise instead of dispatching elsewhere."""
        why = self._unsupported(xyz, points)
        if why is not None:
            raise NotImplementedError("PointNetSetAbstraction: " + why)
        xyz = xyz.detach().contiguous().float()
        fps_idx = farthest_point_sample(xyz, self.npoint, self.fps_start)
        new_xyz = L.gather_rows(xyz, fps_idx)
        idx = knn_point(self.nsample, xyz, new_xyz) if self.knn else query_ball_point(self.radius, self.nsample, xyz,
                                                                                      new_xyz)
        (c1, c2), (n1, n2) = self.mlp_convs, self.mlp_bns
        out = Fn.SetAbstractionFn.apply(points, xyz, new_xyz, idx, c1.weight, c1.bias, n1.weight, n1.bias,
                                        n1.running_mean, n1.running_var, c2.weight, c2.bias, n2.weight, n2.bias,
                                        n2.running_mean, n2.running_var, self.training, n1.eps, n1.momentum, n2.eps,
                                        n2.momentum)
        if self.training:
            with torch.no_grad():
                n1.num_batches_tracked += 1
                n2.num_batches_tracked += 1
        return new_xyz, out


class PointNetFeaturePropagation(nn.Module):
    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last = out_channel

    def forward(self, xyz1, xyz2, points1, points2):
        """xyz1 [B,C,N], xyz2 [B,C,S], points1 [B,D,N] or None, points2 [B,D,S] -> [B,D',N] (3-NN interpolation).
        The 3DViT models build this with an empty MLP (models/3DViT/model.py:62) -- the only form implemented."""
        if len(self.mlp_convs):
            raise NotImplementedError("PointNetFeaturePropagation with a conv MLP is not used by the 3DViT models and "
                                      "has no sm_100a path")
        xyz1 = xyz1.permute(0, 2, 1).contiguous()
        xyz2 = xyz2.permute(0, 2, 1).contiguous()
        points2 = points2.permute(0, 2, 1).contiguous()
        B, N, _ = xyz1.shape
        S = xyz2.shape[1]
        if S < 3:
            raise NotImplementedError("PointNetFeaturePropagation needs >= 3 source points")
        idx, dists = knn_point(3, xyz2, xyz1, return_dist=True)  # queries = xyz1
        interpolated = Fn.ThreeNNInterpFn.apply(points2, idx, dists, None)
        if points1 is not None:
            interpolated = torch.cat([points1.permute(0, 2, 1), interpolated], dim=-1)
        return interpolated.permute(0, 2, 1)
