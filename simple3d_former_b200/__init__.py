"""simple3d_former_b200 -- B200-native (sm_100a) encoder hot path of VITA-Group/Simple3D-Former.

Drop-in, timm-0.3.2-compatible modules (`Attention`, `Mlp`, `Block`, `VisionTransformer`), the voxel tokenizers
(`VoxelEmbed`, `VoxelEmbed_no_average`), the point-grouping functions of `data/pointnet_util.py`, and a data-parallel
training step, all backed by hand-written CUDA kernels reached through the C ABI in include/s3d_b200.h.
"""
__version__ = "0.1.0"
