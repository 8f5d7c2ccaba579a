from simple3d_former_b200.vision_transformer import DropPath, to_2tuple, trunc_normal_  # noqa: F401
