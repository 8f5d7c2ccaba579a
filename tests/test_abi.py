"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/s3d_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "s3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s3d_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from simple3d_former_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        L.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in s3d_b200.h but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(L.SIGNATURES) == set(names)


def test_version_and_error_strings():
    from simple3d_former_b200 import _lib as L
    lib = L.lib()
    assert lib.s3d_abi_version() == 4
    assert lib.s3d_error_string(0) == b"ok"
    assert b"aligned" in lib.s3d_error_string(-3)
    assert b"NULL" in lib.s3d_error_string(-4)


def test_argument_errors_do_not_need_a_gpu():
    """Negative status codes are produced by host-side validation before anything is launched."""
    from simple3d_former_b200 import _lib as L
    lib = L.lib()
    assert lib.s3d_knn(None, None, None, None, 1, 16, 4, 16, None) == -4  # NULL pointers
    assert lib.s3d_knn(None, None, None, None, 1, 8, 4, 16, None) == -1  # K > N
    assert lib.s3d_fps(None, None, None, 1, 100000, 4, None) == -1  # N too large
    assert lib.s3d_layernorm_fwd(None, None, None, None, None, None, None, None, None, 4, 770, 1e-6, None) == -1
    assert lib.s3d_attn_fwd(None, None, None, None, None, 1, 1, 4, 40, 0, 0, 0, 0, 0, 0, 1.0, None, 0, 0.0, None) == -2  # head_dim (48 / 64 / 96 / 192 / 256 exist)
    assert lib.s3d_gemm_bf16(None, None, None, 0, 1, 1, 8, 8, 8, 0, 0, 0, 1.0, None, None, 0, 0, None, 0, None, 0, 1, 0,
                             0, 0, 0, 0, 0, 0, None) == -1


def test_product_raises_without_cuda():
    import torch
    from simple3d_former_b200.vision_transformer import Block
    blk = Block(64, 1)
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, 4, 64))


def test_product_does_not_import_oracle():
    """The shipped package must never route through the oracle."""
    pkg = os.path.join(ROOT, "simple3d_former_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "s3d_oracle" not in text and "reference_harness" not in text, os.path.join(dp, f)
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), os.path.join(dp, f)


def test_attn_bwd_workspace_bytes_host_side():
    """s3d_attn_bwd_workspace_bytes is pure host arithmetic (no CUDA call): two bf16 score matrices in 64-key panels,
    each rounded up to 1 KiB; 0 when the single-score-pass backward does not apply."""
    from simple3d_former_b200 import _lib as L
    f = L.lib().s3d_attn_bwd_workspace_bytes
    B, H, N, dh = 15, 4, 12544, 192
    one = (2 * B * H * N * ((N + 63) // 64 * 64) + 1023) // 1024 * 1024
    assert f(B, H, N, dh) == 2 * one
    assert f(64, 3, 197, 256) == 2 * ((2 * 64 * 3 * 197 * 256 + 1023) // 1024 * 1024)  # head_dim 256: spill-only form
    assert f(12544, 3, 15, 256) == 0      # 15-token sequences keep the warp-per-sequence kernels
    assert f(2, 4, 1024, 100) == 0        # head_dim without a tcgen05 kernel
    assert f(0, 4, 1024, 64) == 0
