"""In-place conversion of a reference model (built on timm 0.3.2 modules) to the fused sm_100a modules.

`convert(model)` swaps every timm-style `Block` (anything exposing norm1 / attn.qkv / attn.proj / norm2 / mlp.fc1 / mlp.fc2)
for `simple3d_former_b200.Block`, every `VoxelEmbed*` for the GEMM tokenizer and `nn.TransformerEncoderLayer` (group_embed)
for `GroupEmbedLayer`, re-using the SAME nn.Parameter objects, so optimizers, checkpoints and DDP wrappers built on the
original model keep working. This is the integration path for an unmodified reference checkout.
"""
import torch.nn as nn

from . import embed_layer_3d_modality as E
from .models import GroupEmbedLayer
from .vision_transformer import Block


def _is_timm_block(m):
    return all(hasattr(m, a) for a in ("norm1", "attn", "norm2", "mlp")) and hasattr(m.attn, "qkv") and \
        hasattr(m.attn, "proj") and hasattr(m.mlp, "fc1") and hasattr(m.mlp, "fc2") and not isinstance(m, Block)


def _p(m, name):
    d = getattr(m, name, None)
    return float(getattr(d, "p", 0.0) or 0.0)


def _convert_block(b):
    dim = b.attn.qkv.in_features
    # the fused Block implements the configuration every reference backbone uses: all dropout rates 0, exact-erf GELU
    # (SURVEY.md Appendix A); anything else must fail here, not train silently without its dropout
    rates = {"attn.attn_drop": _p(b.attn, "attn_drop"), "attn.proj_drop": _p(b.attn, "proj_drop"), "mlp.drop": _p(b.mlp, "drop")}
    if any(rates.values()):
        raise NotImplementedError(f"convert(): timm Block with dropout {rates} has no fused equivalent")
    if not isinstance(getattr(b.mlp, "act", nn.GELU()), nn.GELU):
        raise NotImplementedError("convert(): only the exact-erf GELU Mlp of timm 0.3.2 is implemented")
    nb = Block(dim, b.attn.num_heads, mlp_ratio=b.mlp.fc1.out_features / dim, qkv_bias=b.attn.qkv.bias is not None)
    nb.attn.scale = b.attn.scale
    for name in ("norm1", "norm2"):
        setattr(nb, name, getattr(b, name))
    nb.attn.qkv, nb.attn.proj = b.attn.qkv, b.attn.proj
    nb.mlp.fc1, nb.mlp.fc2 = b.mlp.fc1, b.mlp.fc2
    nb.drop_path = b.drop_path
    nb.train(b.training)
    return nb


def _convert_embed(m):
    cls = {"VoxelEmbed": E.VoxelEmbed, "VoxelEmbed_no_average": E.VoxelEmbed_no_average}.get(type(m).__name__)
    if cls is None or isinstance(m, E._VoxelEmbedBase):
        return None
    conv = m.proj.conv3d_1
    n = cls(m.voxel_size[0], m.cell_size[0], m.patch_size, conv.in_channels, m.embed_dim)
    n.proj = m.proj
    return n


def _convert_encoder_layer(m):
    sa = m.self_attn
    act = getattr(m, "activation", None)
    relu = act is nn.functional.relu or isinstance(act, nn.ReLU) or getattr(act, "__name__", "") == "relu"
    if not relu or getattr(m, "norm_first", False) or getattr(sa, "batch_first", False):
        raise NotImplementedError("convert(): only the post-norm, ReLU, sequence-first nn.TransformerEncoderLayer of "
                                  "vit_3d_2d_pretrain.py:381 has a fused equivalent")
    ps = {_p(m, "dropout"), _p(m, "dropout1"), _p(m, "dropout2"), float(sa.dropout)}
    if len(ps) != 1:
        raise NotImplementedError(f"convert(): the four dropout sites of the encoder layer must share one rate, got {ps}")
    n = GroupEmbedLayer(sa.embed_dim, sa.num_heads, m.linear1.out_features, dropout=ps.pop(), layer_norm_eps=m.norm1.eps)
    n.train(m.training)
    n.self_attn.in_proj_weight, n.self_attn.in_proj_bias = sa.in_proj_weight, sa.in_proj_bias
    n.self_attn.out_proj = sa.out_proj
    n.linear1, n.linear2, n.norm1, n.norm2 = m.linear1, m.linear2, m.norm1, m.norm2
    return n


def convert(model: nn.Module) -> nn.Module:
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            new = None
            if _is_timm_block(child):
                new = _convert_block(child)
            elif isinstance(child, nn.TransformerEncoderLayer):
                new = _convert_encoder_layer(child)
            elif type(child).__name__ in ("VoxelEmbed", "VoxelEmbed_no_average"):
                new = _convert_embed(child)
            if new is not None:
                if isinstance(parent, nn.ModuleList) and name.isdigit():
                    parent[int(name)] = new
                else:
                    setattr(parent, name, new)
    return model
