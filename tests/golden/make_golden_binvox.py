#!/usr/bin/env python
"""Golden vectors for the binvox reader: files written and read back by the REAL reference functions
(/root/reference/utils/binvox_rw.py :: write / read_as_3d_array), with the numpy aliases the file still uses
(np.bool, np.int -- removed in numpy 1.24) patched back for the duration of this script. Run in the build container:

    python tests/golden/make_golden_binvox.py        ->  tests/golden/binvox.pt
"""
import importlib.util
import io
import os

import numpy as np
import torch

REF = os.environ.get("S3D_REFERENCE", "/root/reference")
np.bool = bool  # noqa: the reference predates numpy 1.24
np.int = int    # noqa
spec = importlib.util.spec_from_file_location("ref_binvox_rw", os.path.join(REF, "utils", "binvox_rw.py"))
rw = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rw)

rng = np.random.default_rng(9)


def case(name, dense):
    dims = list(dense.shape)
    vox = rw.Voxels(dense.astype(bool), dims, [0.25, -1.5, 3.0], 41.5, "xyz")
    fp = io.BytesIO()
    rw.write(vox, fp)
    raw = fp.getvalue()
    back = rw.read_as_3d_array(io.BytesIO(raw))
    xzy = rw.read_as_3d_array(io.BytesIO(raw), fix_coords=False)
    assert np.array_equal(back.data, dense.astype(bool))
    return dict(name=name, bytes=raw, dense=torch.from_numpy(back.data.astype(np.uint8)),
                dense_xzy=torch.from_numpy(xzy.data.astype(np.uint8)), dims=back.dims, translate=back.translate,
                scale=back.scale)


V = 30
cases = [case("random_p10_v30", rng.random((V, V, V)) < 0.1),
         case("random_p50_v32", rng.random((32, 32, 32)) < 0.5),
         case("empty_v30", np.zeros((V, V, V), bool)),
         case("full_v30", np.ones((V, V, V), bool)),
         case("alternating_v16", (np.indices((16, 16, 16)).sum(0) % 2).astype(bool))]
blob = np.zeros((32, 32, 32), bool)
blob[5:20, 8:30, 3:9] = True  # long runs that cross x-slab boundaries and the 255 cap
cases.append(case("box_v32", blob))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "binvox.pt")
torch.save({"cases": cases}, out)
print("wrote", out, os.path.getsize(out), "bytes;", [(c["name"], len(c["bytes"])) for c in cases])
