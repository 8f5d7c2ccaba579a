"""Minimal stand-in for timm==0.3.2 (the version pinned by the reference, requirements.txt:6).

TEST INFRASTRUCTURE ONLY. timm is not installed in this image and there is no network, so the symbols the reference
imports (models/vit_3d_2d_pretrain.py:8-10, models/3DViT/model.py:6-8, models/DeIT.py:10-12) are restated here from the
published timm-0.3.2 semantics. With this directory on sys.path the reference's own model files
import and run unchanged on CPU; oracle/s3d_oracle.py uses the same Block/Attention/Mlp as the encoder oracle.
"""
__version__ = "0.3.2"
