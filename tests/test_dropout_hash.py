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
        thresh = int(p * 16384.0 + 0.5)
        out = subprocess.run([exe, str(seed), str(site), str(rows), str(cols), str(thresh)], capture_output=True, text=True,
                             check=True).stdout.split()
        got = np.array([[ch == "1" for ch in line] for line in out])
        want = mod.keep_mask(seed, site, np.arange(rows, dtype=np.uint64) * np.uint64(7919) + np.uint64(3), np.arange(cols), p)
        assert got.shape == want.shape and np.array_equal(got, want)
        assert abs(want.mean() - (1 - p)) < 0.03


def test_mask_statistics():
    """Keep rate, pair statistics at short row / column distances and the variance of row / column sums of the mask
    (a single multiply-fold round fails these: P(drop, drop) at column distance 2 is 2x the product of the marginals)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("test_dropout_gpu", os.path.join(ROOT, "tests", "test_dropout_gpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n, p = 2048, 0.1
    d = ~mod.keep_mask(20210915, 1, np.arange(n) + 12544 * 7, np.arange(n), p)
    q = d.mean()
    assert abs(q - p) < 2e-3
    for a, b in ((d[:, :-1], d[:, 1:]), (d[:, :-2], d[:, 2:]), (d[:, :-4], d[:, 4:]), (d[:-1], d[1:]), (d[:-2], d[2:]),
                 (d[:-1, :-1], d[1:, 1:])):
        assert abs((a & b).mean() / (q * q) - 1.0) < 0.03
    binom = n * q * (1 - q)
    assert abs(d.sum(1).var() / binom - 1.0) < 0.12 and abs(d.sum(0).var() / binom - 1.0) < 0.12
