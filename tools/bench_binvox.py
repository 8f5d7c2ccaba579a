#!/usr/bin/env python
"""Device-side binvox expansion at the cfg3 batch (64 models of 128^3, occupancy p = 0.1): CUDA-event time of
s3d_binvox_scan + s3d_binvox_expand, achieved GB/s on the algorithmic bytes (payload read + uint8 grid written), the
PCIe bytes saved against shipping the dense grid, and the oracle (numpy restatement of the reference reader) beside it."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import binvox_np as BO  # noqa: E402
from simple3d_former_b200 import _lib as L  # noqa: E402

B, V = 64, 128
rng = np.random.default_rng(9)
files = [BO.write(rng.random((V, V, V)) < 0.1) for _ in range(4)]
files = [files[i % 4] for i in range(B)]
payloads = [f[f.index(b"data\n") + 5:] for f in files]
payloads = [p + b"\0" * ((-len(p)) % 16) for p in payloads]  # 16-byte aligned models (empty padding runs)
offs = np.concatenate(([0], np.cumsum([len(p) for p in payloads])))
dev = torch.device("cuda:0")
host = torch.frombuffer(bytearray(b"".join(payloads)), dtype=torch.uint8).pin_memory()
payload = host.to(dev)
offsets = torch.tensor(offs, dtype=torch.long, device=dev)
for _ in range(3):
    grid, totals = L.binvox_expand(payload, offsets, V)
assert int((totals != V ** 3).sum()) == 0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = []
for _ in range(10):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    grid, totals = L.binvox_expand(payload, offsets, V)
    e.record()
    torch.cuda.synchronize()
    ms.append(s.elapsed_time(e))
t = sorted(ms)[len(ms) // 2]
alg = payload.numel() + B * V ** 3  # algorithmic: RLE payload read once + uint8 grid written once
t0 = time.perf_counter()
for f in files[:4]:
    BO.read_as_3d_array(f)
cpu_ms = (time.perf_counter() - t0) / 4 * 1e3
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
print(json.dumps({"workload": f"binvox expand, {B} x {V}^3, p=0.1", "ms": t, "models_per_s": B / t * 1e3,
                  "voxels_per_s": B * V ** 3 / t * 1e3, "algorithmic_bytes": alg, "achieved_gbs": alg / t / 1e6,
                  "hbm_peak_gbs": peaks["hbm_gbs"], "frac": alg / t / 1e6 / peaks["hbm_gbs"],
                  "h2d_payload_bytes": payload.numel(), "h2d_dense_uint8_bytes": B * V ** 3,
                  "h2d_reference_int32_bytes": 4 * B * V ** 3,
                  "cpu_oracle_ms_per_model": cpu_ms, "cpu_oracle_models_per_s": 1e3 / cpu_ms}))
