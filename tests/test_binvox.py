"""binvox reader.

CPU: the numpy oracle (oracle/binvox_np.py) against files written and read back by the REAL reference functions
(tests/golden/binvox.pt, made by tests/golden/make_golden_binvox.py), bit-exact in both directions, and the host-side
header / payload parsing of the product module. GPU: the device expansion through the C ABI against the same fixtures
and, at the BASELINE cfg3 size (128^3), against the oracle and through the write -> read round trip."""
import io

import numpy as np
import pytest
import torch

import binvox_np as BO


def test_oracle_reader_and_writer_match_reference(golden):
    for c in golden("binvox")["cases"]:
        got = BO.read_as_3d_array(c["bytes"])
        assert got.dtype == bool and np.array_equal(got, c["dense"].numpy().astype(bool)), c["name"]
        assert np.array_equal(BO.read_as_3d_array(c["bytes"], fix_coords=False), c["dense_xzy"].numpy().astype(bool))
        dims, translate, scale, _ = BO.read_header(c["bytes"])
        assert dims == c["dims"] and translate == c["translate"] and scale == c["scale"]
        # the writer restatement reproduces the reference's bytes (same header formatting, same 255-capped runs)
        assert BO.write(c["dense"].numpy(), translate=c["translate"], scale=c["scale"]) == c["bytes"], c["name"]


def test_host_parsing_and_errors(golden):
    from simple3d_former_b200 import binvox_rw as P
    c = golden("binvox")["cases"][0]
    dims, translate, scale = P.read_header(io.BytesIO(c["bytes"]))
    assert dims == c["dims"] and translate == c["translate"] and scale == c["scale"]
    d, _, _, payload = P._payload(c["bytes"])
    assert d == c["dims"] and len(payload) % 2 == 0
    with pytest.raises(IOError):
        P.read_header(io.BytesIO(b"#notbinvox\n"))
    with pytest.raises(ValueError):
        P._payload(b"#binvox 1\ndim 4 4 8\ntranslate 0 0 0\nscale 1\ndata\n\x00\x80")
    with pytest.raises(RuntimeError):  # no CPU fallback
        P.load_voxel_batch([c["bytes"]], device="cpu")


@pytest.mark.gpu
def test_device_expansion_matches_reference(golden):
    from simple3d_former_b200 import binvox_rw as P
    cases = golden("binvox")["cases"]
    for c in cases:
        for dtype in (torch.uint8, torch.int32, torch.float32):
            got = P.load_voxel_batch([c["bytes"]], dtype=dtype)
            V = c["dims"][0]
            assert got.shape == (1, 1, V, V, V) and got.dtype == dtype
            assert torch.equal(got[0, 0].cpu().to(torch.uint8), c["dense"]), (c["name"], dtype)
        xzy = P.load_voxel_batch([c["bytes"]], fix_coords=False)
        assert torch.equal(xzy[0, 0].cpu(), c["dense_xzy"]), c["name"]
        v = P.read_as_3d_array(io.BytesIO(c["bytes"]))
        assert v.dims == c["dims"] and v.translate == c["translate"] and v.scale == c["scale"] and v.axis_order == "xyz"
        assert torch.equal(v.data.cpu(), c["dense"])
    # a batch of same-size models in one call (payloads of different lengths back to back)
    same = [c for c in cases if c["dims"][0] == 30]
    got = P.load_voxel_batch([c["bytes"] for c in same])
    for i, c in enumerate(same):
        assert torch.equal(got[i, 0].cpu(), c["dense"]), c["name"]
    # truncated stream: the reference's reshape raises, so do we
    c = cases[0]
    with pytest.raises(ValueError):
        P.load_voxel_batch([c["bytes"][:-2]])


@pytest.mark.gpu
def test_device_expansion_full_size_round_trip():
    """cfg3 size: 128^3 occupancy at p = 0.1 (like bench.py's synthetic voxels), 4 models per call."""
    from simple3d_former_b200 import binvox_rw as P
    rng = np.random.default_rng(3)
    dense = [rng.random((128, 128, 128)) < 0.1 for _ in range(3)]
    solid = np.zeros((128, 128, 128), bool)
    solid[17:101, 40:41, :] = True
    solid[:, :, 127] = True
    dense.append(solid)
    files = [BO.write(d) for d in dense]
    got = P.load_voxel_batch(files).cpu().numpy()
    for i, d in enumerate(dense):
        assert np.array_equal(got[i, 0].astype(bool), d)
        assert np.array_equal(BO.read_as_3d_array(files[i]), d)
    # idempotence: re-encoding the device result gives the same file
    assert BO.write(got[0, 0].astype(bool)) == files[0]


def test_oracle_round_trip_random_grids():
    """write -> read is the identity for cubic grids (the reference's reshape(dims) + transpose is not an identity for
    non-cubic dims -- its own TODO at utils/binvox_rw.py:228 -- and the product rejects those), run lengths are capped at
    255 and the pair stream encodes exactly V^3 voxels (the properties the device expansion relies on)."""
    rng = np.random.default_rng(5)
    for dims, p in [((1, 1, 1), 0.5), ((7, 7, 7), 0.3), ((16, 16, 16), 0.0), ((16, 16, 16), 1.0), ((33, 33, 33), 0.02),
                    ((20, 20, 20), 0.97)]:
        dense = rng.random(dims) < p
        raw = BO.write(dense)
        assert np.array_equal(BO.read_as_3d_array(raw), dense)
        d, _, _, pos = BO.read_header(raw)
        pairs = np.frombuffer(raw, dtype=np.uint8, offset=pos)
        assert list(d) == list(dims) and pairs.size % 2 == 0
        counts = pairs[1::2].astype(np.int64)
        assert counts.min() >= 1 and counts.max() <= 255 and counts.sum() == int(np.prod(dims))
        vals = pairs[0::2]
        same = vals[1:] == vals[:-1]
        assert np.all(counts[:-1][same] == 255)  # a value only repeats across the 255 cap
