from .vision_transformer import VisionTransformer, Block, Attention, Mlp, PatchEmbed, _cfg  # noqa: F401
