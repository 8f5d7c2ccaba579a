"""Loads the UNMODIFIED reference (VITA-Group/Simple3D-Former) model files.

TEST INFRASTRUCTURE ONLY. The reference sources are looked up, in order, at $S3D_REFERENCE_ROOT, /root/reference (the
read-only tree of the build container) and oracle/_ref (byte-for-byte copies of the hot-path files made by
oracle/make_ref.py; git-ignored, travels to the GPU box with the snapshot). Nothing in `-m gpu` tests, smoke() or
bench.py reads /root/reference at run time: on the GPU box only oracle/_ref exists.

What stands in for the reference's third-party / dataset imports:
  * `timm` -> either oracle/timm_shim (timm==0.3.2 restated from its published semantics, plain PyTorch: the CPU
    oracle / baseline configuration), or simple3d_former_b200/timm_compat (the product's drop-in: the reference's own
    class bodies then run on the sm_100a kernels -- the zero-edit integration route of INTEGRATION.md);
  * `pc_util`, `plyfile`, `h5py` -> empty modules (imported by the reference's data/__init__.py:7-12, never used on the path);
  * `data.pointnet_util` -> optionally the product's module (zero-edit route for the point models).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(_HERE, "timm_shim")
COMPAT = os.path.join(os.path.dirname(_HERE), "simple3d_former_b200", "timm_compat")
REF_COPY = os.path.join(_HERE, "_ref")


def reference_root() -> str | None:
    for cand in (os.environ.get("S3D_REFERENCE_ROOT"), "/root/reference", REF_COPY):
        if cand and os.path.isfile(os.path.join(cand, "models", "vit_3d_2d_pretrain.py")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


_cache = {}
_PURGE = ("timm", "models", "data")


def _purge():
    for name in list(sys.modules):
        if name.split(".")[0] in _PURGE:
            del sys.modules[name]
    for p in (SHIM, COMPAT):
        while p in sys.path:
            sys.path.remove(p)


def load(timm: str = "shim", product_pointnet_util: bool = False, root: str | None = None):
    """Returns a namespace with the reference modules: vit (models.vit_3d_2d_pretrain), embed
    (models.embed_layer_3d_modality), point (models.3DViT.model), pointnet_util (data.pointnet_util).

    timm="shim": plain-PyTorch timm 0.3.2 restatement (oracle). timm="compat": the product's timm look-alike, so the
    reference's VisionTransformer subclasses are built from the fused modules. product_pointnet_util=True binds
    `data.pointnet_util` to the product's module before the reference's point model file imports it."""
    key = (timm, product_pointnet_util, root)
    if key in _cache:
        return _cache[key]
    root = root or reference_root()
    if root is None:
        raise RuntimeError("reference sources not found (neither /root/reference nor oracle/_ref)")
    for name in ("pc_util", "plyfile", "h5py"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    _purge()
    sys.path.insert(0, {"shim": SHIM, "compat": COMPAT}[timm])
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(1, root)
    import warnings
    mods = {}
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if product_pointnet_util:
                import simple3d_former_b200.pointnet_util as ours
                pkg = types.ModuleType("data")
                pkg.__path__ = []
                pkg.pointnet_util = ours
                sys.modules["data"] = pkg
                sys.modules["data.pointnet_util"] = ours
            mods["vit"] = importlib.import_module("models.vit_3d_2d_pretrain")
            mods["embed"] = importlib.import_module("models.embed_layer_3d_modality")
            mods["pointnet_util"] = importlib.import_module("data.pointnet_util")
            mods["point"] = importlib.import_module("models.3DViT.model")
            mods["timm_file"] = sys.modules["timm"].__file__
    finally:
        # leave no `timm` / `models` / `data` entries behind: the next load() may bind them differently
        _purge()
        if root in sys.path:
            sys.path.remove(root)
    _cache[key] = types.SimpleNamespace(**mods)
    return _cache[key]


def point_cfg(num_point, num_class, input_dim, backbone="deit_tiny_patch16_224", nneighbor=16):
    """The attribute bag the reference's Hydra config provides to PointTransformerCls/Seg (models/3DViT/model.py:199-202)."""
    model = types.SimpleNamespace(nblocks=4, nneighbor=nneighbor, transformer_backbone=backbone, pretrained=False,
                                  head="Linear", transformer_dim=512)
    return types.SimpleNamespace(num_point=num_point, num_class=num_class, input_dim=input_dim, model=model)
