"""CPU: the drop-in modules expose the reference's surface -- constructor arguments, attributes and state-dict keys /
shapes -- so reference checkpoints load and the reference scripts can construct them unchanged.
Checked against torch's own layer always, and against the UNMODIFIED reference modules when the tree is present
(build container; the GPU box skips that part)."""
import types

import pytest
import torch


def _shapes(sd):
    return {k: tuple(v.shape) for k, v in sd.items()}


def test_group_embed_layer_matches_torch_encoder_layer_keys():
    from simple3d_former_b200.models import GroupEmbedLayer
    ours = GroupEmbedLayer(d_model=64, nhead=4, dim_feedforward=64)
    ref = torch.nn.TransformerEncoderLayer(d_model=64, nhead=4, dim_feedforward=64)  # vit_3d_2d_pretrain.py:381
    assert _shapes(ours.state_dict()) == _shapes(ref.state_dict())
    assert "_drop_seed" not in ours.state_dict() and ours.dropout_p == 0.1 and ours.self_attn.batch_first is False
    ours.load_state_dict(ref.state_dict())  # strict


def test_voxel_and_point_models_match_reference_state_dicts():
    import reference_harness as H
    if not H.available():
        pytest.skip("reference tree not present (GPU box)")
    ref = H.load()
    from simple3d_former_b200.embed_layer_3d_modality import VoxelEmbed, VoxelEmbed_no_average
    from simple3d_former_b200.models import Feature3D_ViT2D_V2, PointTransformerCls, PointTransformerSeg
    for pos, emb_ours, emb_ref in (("default", VoxelEmbed(30, 6, 5, embed_dim=384), ref.embed.VoxelEmbed(30, 6, 5, embed_dim=384)),
                                   ("group_embed", VoxelEmbed_no_average(36, 9, 4, embed_dim=384),
                                    ref.embed.VoxelEmbed_no_average(36, 9, 4, embed_dim=384))):
        ours = Feature3D_ViT2D_V2(embed_layer=emb_ours, n_classes=40, transformer_backbone="deit_small_patch16_224",
                                  pretrained=False, pos_embedding=pos)
        theirs = ref.vit.Feature3D_ViT2D_V2(embed_layer=emb_ref, n_classes=40, transformer_backbone="deit_small_patch16_224",
                                            pretrained=False, pos_embedding=pos)
        assert _shapes(ours.state_dict()) == _shapes(theirs.state_dict()), pos
        for attr in ("embed_dim", "n_classes", "pos_embed_type", "transformer_backbone"):
            assert getattr(ours, attr) == getattr(theirs, attr), attr
        assert ours.blocks[0].attn.num_heads == theirs.blocks[0].attn.num_heads
        assert ours.blocks[0].attn.scale == theirs.blocks[0].attn.scale
        ours.load_state_dict(theirs.state_dict())  # strict: reference checkpoints load
    for cls_ours, cls_ref, n_c, d_in in ((PointTransformerCls, ref.point.PointTransformerCls, 40, 6),
                                         (PointTransformerSeg, ref.point.PointTransformerSeg, 50, 22)):
        def cfg():
            m = types.SimpleNamespace(nblocks=4, nneighbor=16, transformer_backbone="deit_tiny_patch16_224", pretrained=False,
                                      head="Linear", transformer_dim=512)
            return types.SimpleNamespace(num_point=1024, num_class=n_c, input_dim=d_in, model=m)
        ours, theirs = cls_ours(cfg()), cls_ref(cfg())
        # strict in both directions, as train_cls.py:75 / train_partseg.py:80 load checkpoints: the reference's dead
        # PointEmbed (models/3DViT/model.py:96-121, 227) is kept as a parameter-only module under the same keys
        assert _shapes(ours.state_dict()) == _shapes(theirs.state_dict()), cls_ours.__name__
        assert any(k.startswith("patch_embed.gather_local_1.") for k in ours.state_dict())
        ours.load_state_dict(theirs.state_dict(), strict=True)
        theirs.load_state_dict(ours.state_dict(), strict=True)
        dead = set(ours.unused_parameter_names())
        assert {n for n, _ in ours.named_parameters() if n.startswith("patch_embed.")} <= dead


def test_point_embed_keys_without_reference_tree():
    """The checkpoint keys of the reference's PointEmbed, spelled out (models/3DViT/model.py:75-121) so the contract is
    also checked where the reference tree is absent (GPU box)."""
    from simple3d_former_b200.models import PointTransformerCls
    m = types.SimpleNamespace(nblocks=4, nneighbor=16, transformer_backbone="deit_tiny_patch16_224", pretrained=False,
                              head="Linear", transformer_dim=512)
    sd = _shapes(PointTransformerCls(types.SimpleNamespace(num_point=1024, num_class=40, input_dim=6, model=m)).state_dict())
    want = {"patch_embed.conv1.weight": (64, 6, 1), "patch_embed.conv2.weight": (64, 64, 1),
            "patch_embed.bn1.weight": (64,), "patch_embed.bn2.running_var": (64,),
            "patch_embed.gather_local_0.conv1.weight": (48, 128, 1), "patch_embed.gather_local_0.conv2.weight": (48, 48, 1),
            "patch_embed.gather_local_0.bn2.num_batches_tracked": (),
            "patch_embed.gather_local_1.conv1.weight": (48, 256, 1), "patch_embed.gather_local_1.bn1.bias": (48,)}
    for k, shp in want.items():
        assert sd.get(k) == shp, (k, sd.get(k))
    assert sum(k.startswith("patch_embed.") for k in sd) == 36


def test_vit21k_key_remap_matches_reference_fit_dict():
    """models.remap_vit21k_keys vs the reference's fit_dict (vit_3d_2d_pretrain.py:16-36) on a synthetic checkpoint with
    the jax->PyTorch key layout; without the reference tree the expected keys are checked directly."""
    from simple3d_former_b200.models import remap_vit21k_keys
    g = torch.Generator().manual_seed(0)
    ck = {"cls_token": torch.randn(1, 1, 8, generator=g), "transformer.norm.weight": torch.randn(8, generator=g)}
    for i in range(12):
        for c in "qkv":
            ck[f"transformer.blocks.{i}.attn.proj_{c}.weight"] = torch.randn(8, 8, generator=g)
            ck[f"transformer.blocks.{i}.attn.proj_{c}.bias"] = torch.randn(8, generator=g)
        ck[f"transformer.blocks.{i}.pwff.fc1.weight"] = torch.randn(32, 8, generator=g)
    got = remap_vit21k_keys(ck)
    assert got["blocks.3.attn.qkv.weight"].shape == (24, 8) and "blocks.3.mlp.fc1.weight" in got and "norm.weight" in got
    assert torch.equal(got["blocks.3.attn.qkv.bias"][8:16], ck["transformer.blocks.3.attn.proj_k.bias"])
    import reference_harness as H
    if H.available():
        want = H.load().vit.fit_dict(dict(ck))
        assert set(want) == set(got) and all(torch.equal(want[k], got[k]) for k in want)


def test_unsupported_reference_options_raise_at_construction():
    from simple3d_former_b200.embed_layer_3d_modality import VoxelEmbed
    from simple3d_former_b200.models import Feature3D_ViT2D_V2
    kw = dict(embed_layer=VoxelEmbed(30, 6, 5, embed_dim=192), n_classes=40, transformer_backbone="deit_tiny_patch16_224",
              pretrained=False)
    for pos in ("no_embed", "weight_sharing"):
        with pytest.raises(NotImplementedError):
            Feature3D_ViT2D_V2(pos_embedding=pos, **kw)
    with pytest.raises(ValueError, match="Unknown positional embedding scheme!"):
        Feature3D_ViT2D_V2(pos_embedding="bogus", **kw)
    with pytest.raises(NotImplementedError):
        Feature3D_ViT2D_V2(pos_embedding="default", head="AMSoftmax", **kw)
    with pytest.raises(ValueError, match="Unknown transformer backbone name!"):
        Feature3D_ViT2D_V2(pos_embedding="default", **{**kw, "transformer_backbone": "resnet50"})
