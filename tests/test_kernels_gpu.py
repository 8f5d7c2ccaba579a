"""GPU: kernel-level parity through the C ABI (every entry point of include/s3d_b200.h) against plain fp32 torch on the
same inputs. The checks live in tools/gpu_probe.py (also usable stand-alone for bring-up); each group prints [PASS]/[FAIL]
lines with the measured error and its tolerance, and a group passes when it printed no [FAIL] and at least one [PASS]."""
import contextlib
import importlib.util
import io
import os

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("gpu_probe", os.path.join(ROOT, "tools", "gpu_probe.py"))
probe = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(probe)


@pytest.mark.parametrize("group", ["gemm_k", "gemm_mn", "gemm_epi", "gemm_batched", "ln", "attn_fwd", "attn_bwd", "attn_tc", "attn_spill", "sgemm",
                                   "elementwise", "points"])
def test_kernel_group(group):
    import torch
    assert torch.cuda.is_available()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        getattr(probe, "g_" + group)()
        torch.cuda.synchronize()
    out = buf.getvalue()
    fails = [ln for ln in out.splitlines() if "[FAIL]" in ln]
    assert not fails, "\n".join(fails)
    assert "[PASS]" in out
