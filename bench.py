#!/usr/bin/env python
"""Benchmark of the Simple3D-Former encoder hot path on B200 (metric: voxels/s or points/s, forward + backward + Adam).

  python bench.py --gpus N --steps K --warmup W [--config cfg2|cfg3|cfg4|cfg5] [--impl ours|reference]

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); weak scaling: the per-GPU batch is fixed and
gradients are averaged with NCCL allreduce. Rank 0 prints ONE JSON line. A "step" is one training step over one
synthetic batch resident in HBM (`value`) or starting from pinned host memory and ending with the loss read back
(`e2e`). `--impl reference` times the reference's CPU implementation of the same step (the oracle port: the reference
tree does not exist on the GPU box) on the host cores with a bounded per-step sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[1]: deit_small + VoxelEmbed (cell 6 / patch 5), 30^3 voxels, batch 64, bf16, 1xB200
    "cfg2": dict(kind="voxel", model="deit_small_patch16_224+VoxelEmbed(30,6,5)", backbone="deit_small_patch16_224", V=30,
                 cell=6, patch=5, pos="default", average=True, B=64, n_classes=40, unit="voxels/s", per_sample=30 ** 3,
                 cpu_B=64),
    # configs[2]: deit_base (3 heads) + VoxelEmbed_no_average (cell 9 / patch 14) + group_embed, 128^3, batch 64/GPU
    "cfg3": dict(kind="voxel", model="deit_base_patch16_224+VoxelEmbed_no_average(128,9,14)+group_embed",
                 backbone="deit_base_patch16_224", V=128, cell=9, patch=14, pos="group_embed", average=False, B=64,
                 n_classes=55, unit="voxels/s", per_sample=128 ** 3, cpu_B=2),
    # configs[3]: 3DViT point classification, 1024 points, kNN K=16, batch 128/GPU
    "cfg4": dict(kind="point", model="PointTransformerCls(deit_tiny)", backbone="deit_tiny_patch16_224", seg=False, N=1024,
                 input_dim=6, n_classes=40, B=128, unit="points/s", per_sample=1024, cpu_B=8, opt=("sgd", 0.01)),
    # configs[4]: ShapeNetPart part segmentation, 2048 points x 50 parts, batch 32/GPU
    "cfg5": dict(kind="point", model="PointTransformerSeg(deit_tiny)", backbone="deit_tiny_patch16_224", seg=True, N=2048,
                 input_dim=22, n_classes=50, B=32, unit="points/s", per_sample=2048, cpu_B=4, opt=("sgd", 0.05)),
}


# Default workload: BASELINE.json quotes its metric "@1/2/4/8 B200" and north_star's targets (>= 60 % tensor-pipe on the
# attention kernel, >= 6x 1->8 scaling) on "deit_base voxel classification at batch 64/GPU" = configs[2] (cfg3), which fits
# one GPU; configs[1] (cfg2, a 160-us-roofline launch-bound step) is reported alongside it at N=1 as `secondary`.
DEFAULT_CONFIG = "cfg3"

# ncu evidence: `ncu --set full --clock-control none` captures of single launches at the bench shapes, post-processed by
# tools/ncu_table.py into profiles/ncu_table.json (per launch: dram__bytes_read.sum + dram__bytes_write.sum = "traffic",
# sm__pipe_tensor_cycles_active % and gpu__time_duration -- cold-cache, serialised: compare shares, not absolutes).
# Keys are KernelProfile.shape_key strings; attention kernels carry the dropout flag of the variant that was captured,
# so the table entry always describes the variant the timed step runs.
def load_ncu_table():
    path = os.path.join(ROOT, "profiles", "ncu_table.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return {"entries": {}, "source": None}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ----------------------------------------------------------------------------------------------------------------
# synthetic data: binary occupancy grids / unit-ball point clouds, random labels
# ----------------------------------------------------------------------------------------------------------------
def synthetic_batch(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    if cfg["kind"] == "voxel":
        V = cfg["V"]
        x = (torch.rand(B, 1, V, V, V, generator=g) < 0.1).float()
        y = torch.randint(0, cfg["n_classes"], (B,), generator=g)
        return x, y
    N, d = cfg["N"], cfg["input_dim"]
    xyz = torch.rand(B, N, 3, generator=g) * 2 - 1
    xyz = xyz / xyz.norm(dim=-1).max(dim=1, keepdim=True)[0][..., None].clamp_min(1e-6)
    feats = F.normalize(torch.randn(B, N, 3, generator=g), dim=-1)
    x = torch.cat([xyz, feats], dim=-1)
    if d > 6:
        onehot = F.one_hot(torch.randint(0, d - 6, (B,), generator=g), d - 6).float()
        x = torch.cat([x, onehot[:, None, :].expand(-1, N, -1)], dim=-1)
    y = torch.randint(0, cfg["n_classes"], (B, N) if cfg["seg"] else (B,), generator=g)
    return x.contiguous(), y


def loss_fn_for(cfg):
    n = cfg["n_classes"]
    return lambda logits, y: F.cross_entropy(logits.reshape(-1, n), y.reshape(-1))


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------------------------
def _reference_modules():
    """The reference's OWN modules from oracle/_ref (byte-for-byte copies made by oracle/make_ref.py in the build
    container; they travel with the snapshot) on the plain-PyTorch timm 0.3.2 restatement -- or None when absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_harness as H
    if not os.path.isfile(os.path.join(H.REF_COPY, "models", "vit_3d_2d_pretrain.py")):
        return None
    return H, H.load(timm="shim", root=H.REF_COPY)


def _build_reference_model(cfg, H, mods, O):
    if cfg["kind"] == "voxel":
        D = O.BACKBONES[cfg["backbone"]]["embed_dim"]
        emb = (mods.embed.VoxelEmbed if cfg["average"] else mods.embed.VoxelEmbed_no_average)(cfg["V"], cfg["cell"],
                                                                                        cfg["patch"], embed_dim=D)
        model = mods.vit.Feature3D_ViT2D_V2(embed_layer=emb, n_classes=cfg["n_classes"],
                                            transformer_backbone=cfg["backbone"], pretrained=False,
                                            pos_embedding=cfg["pos"])
        for p_ in [model.head.weight, model.head.bias, model.pos_embed, *model.patch_embed.parameters()]:
            p_.requires_grad = False  # what the reference's pretrained path does (vit_3d_2d_pretrain.py:428-432)
        return model
    return (mods.point.PointTransformerSeg if cfg["seg"] else mods.point.PointTransformerCls)(
        H.point_cfg(cfg["N"], cfg["n_classes"], cfg["input_dim"], backbone=cfg["backbone"]))


def cpu_reference_step_fn(cfg, B):
    """One training step of the reference on the host cores. Returns (step_fn, kind): kind "reference" = the unmodified
    reference modules (oracle/_ref) driven exactly as train_cls_voxel.py:270-288 / train_cls.py:103-128 drive them;
    kind "port" = the functional oracle restatement (only when oracle/_ref is absent)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import s3d_oracle as O
    torch.set_num_threads(os.cpu_count())
    kind, lr = cfg.get("opt", ("adam", 1e-3))  # voxel: Adam (train_cls_voxel.py:195); point: SGD momentum 0.9
    x, y = synthetic_batch(cfg, B, seed=9)     # (train_cls.py:91, train_partseg.py:95 with config/*.yaml optimizer: SGD)
    lf = loss_fn_for(cfg)
    ref = _reference_modules()
    if ref is not None:
        import contextlib
        H, mods = ref
        torch.manual_seed(9)
        with contextlib.redirect_stdout(sys.stderr):  # the reference's constructors print the backbone name
            model = _build_reference_model(cfg, H, mods, O)
        model.train()
        params = [p_ for p_ in model.parameters() if p_.requires_grad]
        opt = torch.optim.SGD(params, lr=lr, momentum=0.9) if kind == "sgd" else torch.optim.Adam(params, lr=lr)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = lf(model(x), y)
            loss.backward()
            opt.step()
            return float(loss.detach())

        return step, "reference"
    if cfg["kind"] == "voxel":
        sd = O.init_voxel_state_dict(cfg["backbone"], cfg["cell"], cfg["patch"], cfg["n_classes"], cfg["pos"], seed=9)
        frozen = ()
    else:
        sd = O.init_point_state_dict(cfg["backbone"], cfg["input_dim"], cfg["n_classes"], seed=9)
        frozen = tuple(k for k in sd if "running_" in k)
    params = {k: v.requires_grad_(True) for k, v in sd.items() if k not in frozen}
    sd = {**sd, **params}
    if kind == "sgd":
        opt = torch.optim.SGD(list(params.values()), lr=lr, momentum=0.9)
    else:
        opt = torch.optim.Adam(list(params.values()), lr=lr)
    starts = None
    if cfg["kind"] == "point":
        starts = [torch.zeros(B, dtype=torch.long).numpy(), torch.zeros(B, dtype=torch.long).numpy()]

    def step():
        opt.zero_grad(set_to_none=True)
        if cfg["kind"] == "voxel":
            logits = O.voxel_vit_logits(sd, x, cfg["backbone"], cfg["cell"], cfg["patch"], cfg["pos"])
        else:
            logits = O.point_vit_logits(sd, x, cfg["backbone"], cfg["N"], 16, starts, training=True, seg=cfg["seg"])
        loss = lf(logits, y)
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, "port"


def run_cpu(cfg, B, steps, warmup, budget_s=None):
    step, kind = cpu_reference_step_fn(cfg, B)
    for _ in range(warmup):
        step()
    times = []
    t_begin = time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s and i >= 1:
            break
    mean = sum(times) / len(times)
    return B * cfg["per_sample"] / mean, mean * 1e3, len(times), kind


def cpu_sample_note(cfg, B, n, ms, kind, cores):
    what = ("the reference's own modules (oracle/_ref: unmodified models/*.py, data/pointnet_util.py on the timm-0.3.2 "
            "restatement)" if kind == "reference" else "oracle port of the reference modules")
    note = f"{what}, torch CPU fp32, {cores} threads, batch {B} fwd+bwd+optimizer, {n} steps, {ms:.0f} ms/step"
    if B != cfg["B"]:
        note += (f"; the GPU arm runs batch {cfg['B']}/GPU -- the CPU sample is bounded to batch {B} and its rate is per "
                 f"sample" + ("; group_embed attention costs O(batch^2) per step, so batch " + str(cfg["B"]) +
                              " would be ~17 % SLOWER per sample on the CPU (a same-batch CPU rate would be lower, "
                              "the GPU/CPU ratio higher)" if cfg.get("pos") == "group_embed" else ""))
    return note


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def build_model(cfg, device):
    import types
    from simple3d_former_b200.embed_layer_3d_modality import VoxelEmbed, VoxelEmbed_no_average
    from simple3d_former_b200.models import BACKBONES, Feature3D_ViT2D_V2, PointTransformerCls, PointTransformerSeg
    torch.manual_seed(9)  # train_cls_voxel.py:383
    if cfg["kind"] == "voxel":
        D = BACKBONES[cfg["backbone"]]["embed_dim"]
        emb = (VoxelEmbed if cfg["average"] else VoxelEmbed_no_average)(cfg["V"], cfg["cell"], cfg["patch"], embed_dim=D)
        model = Feature3D_ViT2D_V2(embed_layer=emb, n_classes=cfg["n_classes"], transformer_backbone=cfg["backbone"],
                                   pretrained=False, pos_embedding=cfg["pos"])
        model.freeze_image_branch()
        exclude = ()
    else:
        mc = types.SimpleNamespace(nblocks=4, nneighbor=16, transformer_backbone=cfg["backbone"], pretrained=False,
                                   head="Linear", transformer_dim=512)
        pc = types.SimpleNamespace(num_point=cfg["N"], num_class=cfg["n_classes"], input_dim=cfg["input_dim"], model=mc)
        model = (PointTransformerSeg if cfg["seg"] else PointTransformerCls)(pc)
        exclude = model.unused_parameter_names()
    return model.to(device).train(), exclude


class KernelProfile:
    """CUDA-event timing of every C-ABI launch (installed into _lib.call for an instrumented pass)."""

    def __init__(self):
        self.records = []

    def wrap(self, L):
        orig = L.call

        def call(name, *args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(name, *args)
            e.record()
            self.records.append((name, args, s, e))

        L.call = call
        return orig

    @staticmethod
    def work(name, a):
        """Algorithmic (flops, bytes) of one launch."""
        if name == "s3d_gemm_bf16":
            M, N, K, batch = a[3], a[4], a[5], a[21]
            out_b = (4.0 if a[11] else 2.0) + (4.0 if a[14] else 0.0) + (2.0 if a[17] else 0.0) + (2.0 if a[19] else 0.0)
            return 2.0 * M * N * K * batch, batch * (2.0 * (M * K + N * K) + out_b * M * N)
        if name == "s3d_attn_fwd":
            B, H, N, dh = a[5], a[6], a[7], a[8]
            return 4.0 * B * H * N * N * dh, 2.0 * 4 * B * H * N * dh
        if name == "s3d_attn_bwd":
            B, H, N, dh = a[10], a[11], a[12], a[13]
            return 10.0 * B * H * N * N * dh, 2.0 * 8 * B * H * N * dh
        if name == "s3d_layernorm_fwd":  # read x, [addend]; write [sum], [bf16], [f32]
            per = 4.0 + (4.0 if a[1] else 0.0) + (4.0 if a[2] else 0.0) + (2.0 if a[5] else 0.0) + (4.0 if a[6] else 0.0)
            return 0.0, a[9] * a[10] * per
        if name == "s3d_layernorm_bwd":  # read x, dy (bf16 / f32), [dres]; write dx, [bf16 copy]
            per = 4.0 + (2.0 if a[1] else 4.0) + (4.0 if a[6] else 0.0) + 4.0 + (2.0 if a[8] else 0.0)
            return 0.0, a[12] * a[13] * per
        if name == "s3d_colsum_bf16":
            return 0.0, a[2] * a[3] * 2.0
        return 0.0, 0.0

    @staticmethod
    def shape_key(name, a):
        if name == "s3d_gemm_bf16":
            return f"{name}[M={a[3]},N={a[4]},K={a[5]},a_mn={a[9]},b_mn={a[10]},epi={a[16]},f32={a[11]}]"
        if name == "s3d_attn_fwd":
            return f"{name}[B={a[5]},H={a[6]},N={a[7]},dh={a[8]},drop={int(bool(a[16]) and a[18] > 0)}]"
        if name == "s3d_attn_bwd":
            return f"{name}[B={a[10]},H={a[11]},N={a[12]},dh={a[13]},drop={int(bool(a[21]) and a[23] > 0)}]"
        return name

    def summary(self):
        """Returns (per-family totals, per-(kernel, shape) totals)."""
        torch.cuda.synchronize()
        fam, shapes = {}, {}
        for name, args, s, e in self.records:
            ms = s.elapsed_time(e)
            fl, by = self.work(name, args)
            for table, key in ((fam, name), (shapes, self.shape_key(name, args))):
                f = table.setdefault(key, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
                f["ms"] += ms
                f["flops"] += fl
                f["bytes"] += by
                f["launches"] += 1
        return fam, shapes


def run_ours(args, cfg, rank, world, local_rank):
    from simple3d_former_b200 import _lib as L
    from simple3d_former_b200.dp import DataParallelTrainer
    import torch.distributed as dist

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    L.lib()  # fail loudly if the CUDA extension is missing
    model, exclude = build_model(cfg, device)
    opt_kind, opt_lr = cfg.get("opt", ("adam", 1e-3))
    trainer = DataParallelTrainer(model, lr=opt_lr, exclude=exclude, optimizer=opt_kind, momentum=0.9)
    lf = loss_fn_for(cfg)
    B = cfg["B"]
    xh, yh = synthetic_batch(cfg, B, seed=9 + rank)
    if cfg["kind"] == "voxel":
        xh = xh.to(torch.uint8)  # binvox occupancy is binary: ship 1 byte per voxel, the patch-gather kernel reads it as is
    xh, yh = xh.pin_memory(), yh.pin_memory()
    x = xh.to(device)
    y = yh.to(device)
    if cfg["kind"] == "point":
        model.set_fps_starts([torch.zeros(B, dtype=torch.long, device=device)] * 2)

    def eager_step():
        return trainer.step(x, y, lf)

    # launches per step (C-ABI calls; attn_bwd enqueues 3 kernels)
    L.LAUNCHES = 0
    n_before = L.LAUNCHES
    eager_step()
    torch.cuda.synchronize()
    launches_per_step = L.LAUNCHES - n_before
    use_graph = not args.no_graph  # N > 1: the NCCL bucket all-reduces are captured with the step (dp.py)
    step_fn = eager_step
    graph_note = "eager"
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    eager_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = eager_step()

            def step_fn():
                graph.replay()
                return static_loss

            graph_note = "cuda_graph(whole step" + (", NCCL bucket all-reduces captured)" if world > 1 else ")")
        except Exception as exc:  # capture is an optimisation, never a requirement
            graph_note = f"eager (graph capture failed: {type(exc).__name__}: {str(exc)[:120]})"
            step_fn = eager_step
            torch.cuda.synchronize()
        if world > 1:  # all ranks must run the same mode (a graph rank and an eager rank would still match collectives,
            ok = torch.tensor([1 if graph_note.startswith("cuda_graph") else 0], device=device)  # but time differently)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0 and graph_note.startswith("cuda_graph"):
                step_fn, graph_note = eager_step, "eager (graph capture failed on another rank)"

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps):
        """Device-resident steps: sum of per-step CUDA-event durations (the L2 flush between steps is outside the pairs)."""
        pairs = []
        for _ in range(n_steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            step_fn()
            e.record()
            pairs.append((s, e))
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in pairs)

    copy_stream = torch.cuda.Stream(device=device)
    stage_x = [torch.empty_like(x), torch.empty_like(x)]
    stage_y = [torch.empty_like(y), torch.empty_like(y)]
    loss_hist = torch.zeros(max(args.steps, 1)).pin_memory()

    def timed_e2e(n_steps):
        """End to end through the public API: every step's batch starts in pinned host memory and its loss ends in host
        memory, all inside ONE timed region (first event before the first H2D copy, last event after the last loss has
        been copied back). Input feeding is double-buffered the way a training loop with a pinned-memory loader does it:
        the H2D copy of batch i+1 runs on a copy stream while step i computes; the step then takes its batch with a
        device-to-device copy into the (CUDA-graph) input tensors. Steps run back to back (no L2 flush: the per-step
        working set is far larger than the 126 MB L2)."""
        main = torch.cuda.current_stream()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        taken = [torch.cuda.Event(), torch.cuda.Event()]
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def feed(i):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(taken[b])  # the step that used this staging buffer has copied it out
                stage_x[b].copy_(xh, non_blocking=True)
                stage_y[b].copy_(yh, non_blocking=True)
                ready[b].record(copy_stream)

        start.record(main)
        copy_stream.wait_event(start)
        feed(0)
        for i in range(n_steps):
            b = i & 1
            if i + 1 < n_steps:
                feed(i + 1)
            main.wait_event(ready[b])
            x.copy_(stage_x[b], non_blocking=True)
            y.copy_(stage_y[b], non_blocking=True)
            taken[b].record(main)
            loss = step_fn()
            loss_hist[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        stop.record(main)
        stop.synchronize()
        return start.elapsed_time(stop)

    for _ in range(max(args.warmup, 3)):
        step_fn()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms = timed(args.steps)
    barrier()
    e2e_ms = timed_e2e(args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()

    # instrumented pass: per-kernel CUDA-event timing of 2 eager steps (separate from the timed region)
    prof = KernelProfile()
    orig = prof.wrap(L)
    try:
        for _ in range(2):
            eager_step()
        fam, shapes = prof.summary()
    finally:
        L.call = orig
    if rank != 0:
        return None
    pk = peaks()
    # dominant kernel = the kernel family (one __global__ template behind one C-ABI entry point) with the largest share
    # of the step's kernel time; achieved = sum of algorithmic flops (bytes) of its launches / sum of their durations
    def rate(f):
        tb = f["flops"] > 0
        a = (f["flops"] / (f["ms"] * 1e-3) / 1e12) if tb else (f["bytes"] / (f["ms"] * 1e-3) / 1e9)
        return tb, a

    name, f = max(fam.items(), key=lambda kv: kv[1]["ms"])
    tensor_bound, achieved = rate(f)
    peak = pk["tf_sustained"] if tensor_bound else pk["hbm"]
    step_kernel_ms = sum(v["ms"] for v in fam.values())
    shapes_sorted = sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])
    roofline = {"kernel": name, "bound": "tensor" if tensor_bound else "hbm", "achieved": round(achieved, 2),
                "peak": peak, "peak_source": pk["src"] + (" sustained bf16" if tensor_bound else " copy"),
                "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": round(achieved / peak, 4),
                "traffic": None,
                "algorithmic_per_launch": round((f["flops"] if tensor_bound else f["bytes"]) / f["launches"], 1),
                "launches_per_step": f["launches"] // 2, "avg_launch_us": round(1e3 * f["ms"] / f["launches"], 2),
                "share_of_kernel_time": round(f["ms"] / step_kernel_ms, 3),
                "timing": "cuda events around every C-ABI launch, instrumented eager pass of 2 steps after the timed region",
                "families": {k: {"ms_per_step": round(v["ms"] / 2, 3), "achieved": round(rate(v)[1], 1),
                                 "unit": "TFLOP/s" if rate(v)[0] else "GB/s"}
                             for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
                "top_shapes": {k: {"ms_per_step": round(v["ms"] / 2, 3), "achieved": round(rate(v)[1], 1),
                                   "unit": "TFLOP/s" if rate(v)[0] else "GB/s"} for k, v in shapes_sorted[:24]}}
    # ncu evidence: DRAM traffic of the dominant family's largest launch (same launch the algorithmic figure next to it
    # refers to) and the per-launch captures of every top shape that has one
    ncu = load_ncu_table()
    NCU = ncu.get("entries", {})
    fam_shapes = [(k, v) for k, v in shapes_sorted if k.split("[")[0] == name]
    if fam_shapes and fam_shapes[0][0] in NCU:
        k0, v0 = fam_shapes[0]
        roofline["traffic"] = NCU[k0]["traffic"]
        roofline["traffic_launch"] = k0
        roofline["traffic_algorithmic"] = round((v0["bytes"]) / v0["launches"], 1)
    roofline["ncu"] = {k: NCU[k] for k, _ in shapes_sorted[:24] if k in NCU}
    roofline["ncu_source"] = ncu.get("source")
    # whole-step figure, so the dominant family's fraction cannot be mistaken for it: algorithmic tensor FLOPs of ALL
    # launches of a step (GEMMs 2MNK; attention 4 / 10 N^2 dh per (batch, head) forward / backward) over the timed step
    step_flops = sum(v["flops"] for v in fam.values()) / 2.0
    roofline["step_frac"] = round(step_flops / (dev_ms / args.steps * 1e-3) / 1e12 / pk["tf_sustained"], 4)
    roofline["step_algorithmic_tflop"] = round(step_flops / 1e12, 3)
    samples = B * world * args.steps
    value = samples * cfg["per_sample"] / (dev_ms * 1e-3)
    e2e_value = samples * cfg["per_sample"] / (e2e_ms * 1e-3)
    h2d = xh.numel() * xh.element_size() + yh.numel() * yh.element_size()
    out = {
        "metric": "voxels/sec fwd+bwd" if cfg["kind"] == "voxel" else "points/sec fwd+bwd",
        "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg['model']}, batch {B}/GPU, fwd+bwd+{'SGD' if cfg.get('opt', ('adam',))[0] == 'sgd' else 'Adam'}, bf16 operands / fp32 accumulate",
                   "global_batch": B * world, "parallelism": f"dp{world}", "launch": graph_note,
                   "l2": "256 MB flush between timed steps; per-step working set (weights + Adam state + activations) >> 126 MB L2",
                   "input": "uint8 occupancy grid (1 B/voxel)" if cfg["kind"] == "voxel" else "fp32 points",
                   "samples_per_s": samples / (dev_ms * 1e-3)},
        "e2e": {"value": e2e_value, "unit": cfg["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps,
                "how": "one timed region over all steps; pinned host batch -> H2D on a copy stream (double-buffered, "
                       "overlapping the previous step) -> step -> loss copied to pinned host memory"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "hbm_peak_allocated_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, ms, n, kind = run_cpu(cfg, cfg["cpu_B"], steps=50, warmup=1, budget_s=15.0)
        cores = torch.get_num_threads()
        out["cpu_baseline"] = {"value": v, "unit": cfg["unit"], "cores": cores, "kind": kind,
                               "sample": cpu_sample_note(cfg, cfg["cpu_B"], n, ms, kind, cores)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--no-secondary", action="store_true", help="skip the extra cfg2 / cfg4 / cfg5 measurements")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        B = cfg["cpu_B"]
        v, ms, n, kind = run_cpu(cfg, B, steps=args.steps, warmup=args.warmup)
        cores = torch.get_num_threads()
        print(json.dumps({
            "impl": "reference", "metric": "voxels/sec fwd+bwd" if cfg["kind"] == "voxel" else "points/sec fwd+bwd",
            "value": v, "unit": cfg["unit"], "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg['model']}, CPU fp32, bounded sample batch {B}, fwd+bwd+optimizer"},
            "cpu_baseline": {"value": v, "unit": cfg["unit"], "cores": cores, "kind": kind,
                             "sample": cpu_sample_note(cfg, B, n, ms, kind, cores)},
            "e2e": {"value": v, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    out = run_ours(args, cfg, rank, world, local_rank)
    if args.config == DEFAULT_CONFIG and not args.no_secondary:
        # the other BASELINE configs measured the same way at the same N, reported inside the same JSON line:
        # configs[1] (cfg2), configs[3] (cfg4: point classification), configs[4] (cfg5: part segmentation)
        import copy
        secondary = {}
        for name in ("cfg2", "cfg4", "cfg5"):
            a2 = copy.copy(args)
            a2.config, a2.no_cpu_baseline, a2.steps = name, True, max(args.steps, 20)
            torch.cuda.empty_cache()
            sec = run_ours(a2, CONFIGS[name], rank, world, local_rank)
            if rank == 0:
                secondary[name] = {k: sec[k] for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling", "config",
                                                        "e2e", "gpu_launches")}
                secondary[name]["roofline"] = {k: sec["roofline"][k] for k in ("kernel", "achieved", "peak", "unit", "frac",
                                                                               "share_of_kernel_time", "step_frac")}
        if rank == 0:
            out["secondary"] = secondary
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
