#!/usr/bin/env python
"""GPU time per launch of the small GEMMs of cfg2 / cfg4 (CUDA-graph replay of 20 back-to-back launches, so host launch
overhead is excluded), per (BN, cluster) variant. Used to tune the tile / cluster heuristic for latency-bound shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from simple3d_former_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = [("cfg2 qkv fwd", 1664, 1152, 384, False, False, False), ("cfg2 fc1 fwd", 1664, 1536, 384, False, False, False),
          ("cfg2 fc2 fwd f32", 1664, 384, 1536, False, False, True), ("cfg2 dX fc1", 1664, 384, 1536, False, True, False),
          ("cfg2 dW fc1", 1536, 384, 1664, True, True, True), ("cfg2 dW proj", 384, 384, 1664, True, True, True),
          ("cfg4 qkv fwd", 32896, 576, 192, False, False, False), ("cfg4 fc2 fwd f32", 32896, 192, 768, False, False, True),
          ("cfg4 dW fc1", 768, 192, 32896, True, True, True)]
for (name, M, N, K, amn, bmn, f32) in shapes:
    a = torch.randn((K, M) if amn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if bmn else (N, K), device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    for (bn, cl) in [(0, 0), (256, 2), (256, 1), (128, 2), (128, 1), (64, 1)]:
        if bn > 0 and N < bn // 2:
            continue
        try:
            for _ in range(2):
                L.gemm(a, b, a_mn=amn, b_mn=bmn, out=out, force_bn=bn, force_cluster=cl)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    L.gemm(a, b, a_mn=amn, b_mn=bmn, out=out, force_bn=bn, force_cluster=cl)
            for _ in range(3):
                g.replay()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                g.replay()
            e.record()
            torch.cuda.synchronize()
            us = s.elapsed_time(e) / 200 * 1e3
            print(f"{name:18s} M{M} N{N} K{K} BN{bn} cl{cl}: {us:7.2f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
        except Exception as exc:  # unsupported variant
            print(f"{name:18s} BN{bn} cl{cl}: {type(exc).__name__} {exc}", flush=True)
            torch.cuda.synchronize()
