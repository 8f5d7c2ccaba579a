from simple3d_former_b200.vision_transformer import (Attention, Block, DropPath, Mlp, PatchEmbed,  # noqa: F401
                                                     VisionTransformer, _cfg)
