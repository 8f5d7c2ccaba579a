"""Peak / live HBM of one eager cfg3 training step (torch allocator statistics), with and without the attention-backward
workspace: python tools/step_mem.py [cfg3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    cfg = bench.CONFIGS[name]
    from simple3d_former_b200 import _lib as L
    from simple3d_former_b200.dp import DataParallelTrainer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model, exclude = bench.build_model(cfg, dev)
    trainer = DataParallelTrainer(model, lr=1e-3, exclude=exclude)
    lf = bench.loss_fn_for(cfg)
    x, y = bench.synthetic_batch(cfg, cfg["B"], seed=9)
    if cfg["kind"] == "voxel":
        x = x.to(torch.uint8)
    x, y = x.to(dev), y.to(dev)
    print(f"static (weights, grads, optimizer state, batch): {torch.cuda.memory_allocated() / 2**30:.2f} GiB")
    for ws in (64, 0):
        L._ATTN_WS_MAX = ws << 30
        trainer.step(x, y, lf)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        trainer.step(x, y, lf)
        torch.cuda.synchronize()
        print(f"workspace cap {ws} GiB: peak allocated {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB, "
              f"reserved {torch.cuda.memory_reserved() / 2**30:.2f} GiB")


if __name__ == "__main__":
    main()
