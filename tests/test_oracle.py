"""CPU: the oracle restatement reproduces the outputs the UNMODIFIED reference produced (tests/golden/*.pt)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import s3d_oracle as O


def _voxel_sd(fix):
    sd = O.init_voxel_state_dict(fix["backbone"], fix["cell"], fix["patch"], fix["n_classes"], fix["pos"], seed=fix["weight_seed"])
    g = torch.Generator().manual_seed(fix["embed_seed"])
    for k in ("voxel_pos_embed", "group_pos_embed", "group_cls_token"):
        if k in sd:
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.02
    assert abs(O.state_dict_checksum(sd) - fix["sd_checksum"]) <= 1e-6 * abs(fix["sd_checksum"])
    return sd


@pytest.mark.parametrize("name", ["cfg1_deit_small_voxel30", "cfg3_small_deit_base_group36", "cfg3_deit_base_group128",
                                  "cfg3_deit_base_group128_b3"])
def test_voxel_models_match_reference(golden, name):
    fix = golden(name)
    sd = {k: v.requires_grad_(True) for k, v in _voxel_sd(fix).items()}
    x, y = O.synthetic_voxels(fix["B"], fix["V"], seed=fix["input_seed"], n_classes=fix["n_classes"])
    logits = O.voxel_vit_logits(sd, x, fix["backbone"], fix["cell"], fix["patch"], fix["pos"])
    assert torch.allclose(logits, fix["logits"], atol=2e-5, rtol=1e-4)
    loss = F.cross_entropy(logits, y)
    assert abs(float(loss) - fix["loss"]) < 1e-5
    loss.backward()
    for k, ref in fix["grads"].items():
        g = sd[k].grad
        assert g is not None, k
        assert abs(float(g.norm()) - ref["norm"]) <= 1e-3 * ref["norm"] + 1e-8, k
        assert torch.allclose(g.flatten()[:16], ref["head"], atol=1e-6 + 1e-3 * float(ref["head"].abs().max())), k


def test_forward_images_matches_reference(golden):
    """Feature3D_ViT2D_V2.forward_images (vit_3d_2d_pretrain.py:435-451), SURVEY.md section 8 row a13."""
    fix = golden("cfg1_forward_images")
    sd = O.init_voxel_state_dict(fix["backbone"], 6, 5, 40, "default", seed=fix["weight_seed"])
    sd.update(O.init_image_branch_state_dict(fix["backbone"], seed=fix["image_seed"]))
    assert abs(O.state_dict_checksum(sd) - fix["sd_checksum"]) <= 1e-6 * abs(fix["sd_checksum"])
    x = torch.randn(fix["B"], 3, 224, 224, generator=torch.Generator().manual_seed(fix["input_seed"]))
    with torch.no_grad():
        logits = O.voxel_vit_forward_images(sd, x, fix["backbone"])
    assert torch.allclose(logits, fix["logits"], atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("name", ["cfg4_point_cls_tiny1024", "cfg5_point_seg_tiny2048", "cfg4_point_cls_tiny1024_sharp",
                                  "cfg5_point_seg_tiny2048_sharp"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_point_models_match_reference(golden, name, mode):
    fix = golden(name)
    sd = O.init_point_state_dict(fix["backbone"], fix["input_dim"], fix["n_classes"], seed=fix["weight_seed"])
    if fix.get("sharp"):
        sd = O.sharpen_point_state_dict(sd, head_gain=fix["head_gain"])
    assert abs(O.state_dict_checksum(sd) - fix["sd_checksum"]) <= 1e-6 * abs(fix["sd_checksum"])
    x, _ = O.synthetic_points(fix["B"], fix["N"], extra=fix["input_dim"] - 3, seed=fix["input_seed"], n_classes=fix["n_classes"])
    starts = [s.numpy() for s in fix["fps_starts"]]
    with torch.no_grad():
        logits = O.point_vit_logits(sd, x, fix["backbone"], fix["N"], 16, starts, training=(mode == "train"), seg=fix["seg"])
    assert torch.allclose(logits, fix[mode]["logits"], atol=2e-5, rtol=1e-4)


def test_point_ops_match_reference(golden):
    fix = golden("pointops")
    for c in fix["cases"]:
        xyz, q = c["xyz"].numpy(), c["query"].numpy()
        assert np.array_equal(O.knn_np(xyz, q, c["K"]), c["knn"].numpy())
        d = O.square_distance_np(q, xyz)
        assert np.array_equal(np.sort(d, axis=-1)[:, :, :c["K"]], c["knn_dist"].numpy())
        assert np.array_equal(O.ball_query_np(c["radius"], c["nsample"], xyz, q), c["ball"].numpy())
        assert np.array_equal(O.fps_np(xyz, c["S"], c["fps_start"].numpy()), c["fps"].numpy())
        pts = torch.randn(c["B"], c["N"], 7)  # gather is checked through its index algebra
        assert torch.equal(O.index_points(pts, c["knn"]), pts[torch.arange(c["B"])[:, None, None], c["knn"]])
    t = fix["tie_case"]
    assert np.array_equal(O.knn_np(t["xyz"].numpy(), t["xyz"].numpy(), t["K"]), t["knn_stable"].numpy())


def test_point_ops_edge_cases():
    xyz = np.zeros((1, 16, 3), np.float32)  # all points identical: every distance ties -> index order
    assert np.array_equal(O.knn_np(xyz, xyz[:, :2], 16)[0, 0], np.arange(16))
    far = np.full((1, 4, 3), 5.0, np.float32)
    assert (O.ball_query_np(0.2, 8, xyz, far) == 16).all()  # no hit -> filled with N, as the reference does
    assert np.array_equal(O.fps_np(xyz, 4, np.array([3]))[0], np.array([3, 0, 0, 0]))  # argmax ties -> first index


def test_reference_harness_agrees_when_available():
    import reference_harness as H
    if not H.available():
        pytest.skip("reference tree not present (GPU box)")
    ref = H.load()
    emb = ref.embed.VoxelEmbed(30, 6, 5, embed_dim=384)
    m = ref.vit.Feature3D_ViT2D_V2(embed_layer=emb, n_classes=40, transformer_backbone="deit_small_patch16_224",
                                   pretrained=False, pos_embedding="default").eval()
    x, _ = O.synthetic_voxels(2, 30)
    with torch.no_grad():
        assert torch.allclose(m(x), O.voxel_vit_logits(m.state_dict(), x, "deit_small_patch16_224", 6, 5), atol=1e-6)
