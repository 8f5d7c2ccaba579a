"""`timm` look-alike whose VisionTransformer / Block / Attention / Mlp ARE the fused sm_100a modules.

Put `simple3d_former_b200/timm_compat` first on sys.path and the reference's unmodified model files
(`from timm.models.vision_transformer import VisionTransformer, _cfg`, models/vit_3d_2d_pretrain.py:8-10) build on the
B200 kernels without a single edit. Only the symbols the reference imports are provided.
"""
__version__ = "0.3.2+s3d_b200"
