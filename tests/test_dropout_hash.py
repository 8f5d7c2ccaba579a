"""CPU: the numpy restatement of the dropout mask hash used by the GPU parity tests (tests/test_dropout_gpu.py::keep_mask)
is bit-identical to the C++ implementation in csrc/common.cuh (compiled for the host with nvcc; no GPU needed)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_numpy_hash_matches_cxx(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("test_dropout_gpu", os.path.join(ROOT, "tests", "test_dropout_gpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    exe = str(tmp_path / "drop_hash_host")
    r = subprocess.run(["nvcc", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "simple3d_former_b200", "csrc"),
                        "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "aux", "drop_hash_host.cu")],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("nvcc host build unavailable: " + r.stderr[-300:])
    for seed, site, p in [(20210915, 1, 0.1), (123456789, 4, 0.1), (7, 2, 0.5)]:
        rows, cols = 64, 96
        thresh = int(p * 65536.0 + 0.5)
        out = subprocess.run([exe, str(seed), str(site), str(rows), str(cols), str(thresh)], capture_output=True, text=True,
                             check=True).stdout.split()
        got = np.array([[ch == "1" for ch in line] for line in out])
        want = mod.keep_mask(seed, site, np.arange(rows, dtype=np.uint64) * np.uint64(7919) + np.uint64(3), np.arange(cols), p)
        assert got.shape == want.shape and np.array_equal(got, want)
        assert abs(want.mean() - (1 - p)) < 0.03
