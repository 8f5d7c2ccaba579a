#!/usr/bin/env python
"""Per-kernel breakdown (ours AND the remaining PyTorch ops) of one eager training step of a bench config.

  python tools/torch_profile.py cfg4 [rows]

Uses torch.profiler (CUPTI), so the absolute times are inflated; only the shares are used (which PyTorch-side ops are
still worth replacing with a kernel of ours). Output goes to stdout as a table sorted by device time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    cfg = bench.CONFIGS[name]
    from simple3d_former_b200 import _lib as L
    from simple3d_former_b200.dp import DataParallelTrainer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L.lib()
    model, exclude = bench.build_model(cfg, dev)
    trainer = DataParallelTrainer(model, lr=1e-3, exclude=exclude)
    lf = bench.loss_fn_for(cfg)
    x, y = bench.synthetic_batch(cfg, cfg["B"], seed=9)
    if cfg["kind"] == "voxel":
        x = x.to(torch.uint8)
    x, y = x.to(dev), y.to(dev)
    if cfg["kind"] == "point":
        model.set_fps_starts([torch.zeros(cfg["B"], dtype=torch.long, device=dev)] * 2)
    for _ in range(3):
        trainer.step(x, y, lf)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            trainer.step(x, y, lf)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=rows, max_name_column_width=90))


if __name__ == "__main__":
    main()
