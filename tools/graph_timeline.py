#!/usr/bin/env python
"""Timeline of ONE CUDA-graph replay of a bench config's training step (CUPTI through torch.profiler): span, summed
kernel time, idle gaps between kernels, and the kernels ranked by total device time. Answers "where does the step go
that the per-launch event timing of bench.py does not see" (torch glue kernels, gaps, clock differences).

    python tools/graph_timeline.py cfg3 [rows]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cfg = bench.CONFIGS[name]
    from simple3d_former_b200.dp import DataParallelTrainer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model, exclude = bench.build_model(cfg, dev)
    opt = cfg.get("opt", ("adam",))
    trainer = DataParallelTrainer(model, lr=1e-3, exclude=exclude) if opt[0] == "adam" else \
        DataParallelTrainer(model, lr=1e-3, exclude=exclude, optimizer="sgd", momentum=0.9)
    lf = bench.loss_fn_for(cfg)
    x, y = bench.synthetic_batch(cfg, cfg["B"], seed=9)
    if cfg["kind"] == "voxel":
        x = x.to(torch.uint8)
    x, y = x.to(dev), y.to(dev)
    if cfg["kind"] == "point":
        model.set_fps_starts([torch.zeros(cfg["B"], dtype=torch.long, device=dev)] * 2)

    def step():
        return trainer.step(x, y, lf)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        graph.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy, cur_end, gaps = 0.0, t0, []
    for e in evs:
        s, en = e.time_range.start, e.time_range.end
        if s > cur_end:
            gaps.append((s - cur_end, e.name))
        if en > cur_end:
            busy += en - max(s, cur_end)
            cur_end = en
    tot = collections.Counter()
    cnt = collections.Counter()
    for e in evs:
        key = e.name.replace("(anonymous namespace)::", "").split("(")[0][-70:]
        tot[key] += e.time_range.end - e.time_range.start
        cnt[key] += 1
    print(f"{name}: span {(t1 - t0) / 1e3:.2f} ms, device busy {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms in "
          f"{len(gaps)} gaps, {len(evs)} kernels / copies, summed kernel time {sum(tot.values()) / 1e3:.2f} ms")
    ours = sum(v for k, v in tot.items() if "s3d" in k)
    print(f"our kernels {ours / 1e3:.2f} ms, everything else (torch element-wise / reductions / memsets / NCCL) "
          f"{(sum(tot.values()) - ours) / 1e3:.2f} ms")
    for k, v in tot.most_common(rows):
        print(f"{v / 1e3:9.3f} ms  x{cnt[k]:4d}  {k}")
    gaps.sort(reverse=True)
    print("largest gaps (us, before kernel):", [(round(g, 1), n.split("(")[0][-40:]) for g, n in gaps[:8]])


if __name__ == "__main__":
    main()
