"""Data-parallel training step for the hot path: one process per GPU, gradients-only exchange.

Equivalent of the reference's `DistributedDataParallel(model, device_ids=[gpu])` + `Adam` (train_cls_voxel.py:154-165,
195-198, 270-288), laid out for B200:

  * every trainable parameter is a view into ONE flat fp32 buffer; gradients accumulate into one flat fp32 buffer and the
    bf16 tensor-core shadows live in one flat bf16 buffer (`p._s3d_shadow`), so the optimizer is a single fused kernel
    (`s3d_adam_step`: Adam + 1/world gradient averaging + bf16 shadow refresh in one HBM pass);
  * the flat gradient buffer is cut into ~25 MB buckets in backward order; a bucket's `ncclAllReduce` is enqueued on a
    side stream as soon as its last gradient has been accumulated (post-accumulate-grad hooks), overlapping NVLink
    traffic with the remaining backward kernels; there is no activation exchange;
  * parameters that never receive gradients (image head / pos_embed / patch_embed, dead point-model layers) are
    excluded, which is what keeps the reference's DDP from erroring.

`FlatGradBuckets` is the host-side logic (layout, bucket bookkeeping, collective launches); it is device agnostic so the
N > 1 path is covered by world_size-2 `gloo` tests on CPU. `DataParallelTrainer` adds the CUDA kernels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _align(n, a=64):
    return (n + a - 1) // a * a


class FlatGradBuckets:
    """Flat parameter / gradient storage with bucketed, hook-driven allreduce (SUM; averaging is folded into the
    optimizer kernel as grad_scale = 1 / world)."""

    def __init__(self, named_params, bucket_mb=25.0, process_group=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        named = list(named_params)
        named.reverse()  # gradients become ready roughly in reverse registration order
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        self.device = dev
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += _align(p.numel())
        self.total = total
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[o:o + n].view(p.shape)
            p.grad = self.flat_g[o:o + n].view(p.shape)
        self.buckets = []  # [start, end, n_params]
        limit = max(1, int(bucket_mb * 1024 * 1024 / 4))
        start, count = 0, 0
        self._bucket_of = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            self._bucket_of[i] = len(self.buckets)
            count += 1
            end = o + _align(p.numel())
            if end - start >= limit or i == len(self.params) - 1:
                self.buckets.append([start, end, count])
                start, count = end, 0
        self._pending = [b[2] for b in self.buckets]
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._uses = [0] * len(self.params)      # in-place gradient writes seen this step (gradient sinks)
        self._expected = None                    # learned in the first step: writes per parameter per step
        self._comm_stream = torch.cuda.Stream(device=dev) if (self.world > 1 and dev.type == "cuda") else None
        self._handles = []
        self.launch_order = []  # bucket ids in the order their collectives were enqueued (observability / tests)
        if self.world > 1:
            # identical initial weights on every rank (DDP's constructor broadcast, train_cls_voxel.py:154-159)
            dist.broadcast(self.flat_p, src=0, group=self.group)
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    def enable_sinks(self):
        """Kernels accumulate parameter gradients straight into the flat buffer (functional._wgrad/_bgrad/_ln_bwd) and
        report each write through note_write(); autograd then never materialises or adds those gradients."""
        for p in self.params:
            p._s3d_grad_sink = p.grad
            p._s3d_owner = self

    def note_write(self, p):
        """Bookkeeping only. Readiness of a bucket is signalled by autograd's post-accumulate-grad hook, which the engine
        runs once per parameter after ALL of its uses have been processed -- also for sink parameters whose backward
        returned None (measured: torch 2.11). Buckets whose hooks did not all fire are launched in sync_gradients()."""
        self._uses[self._index[id(p)]] += 1

    def _make_hook(self, i):
        b = self._bucket_of[i]

        def hook(_param):
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch_bucket(b)

        return hook

    def _launch_bucket(self, b):
        s, e, _ = self.buckets[b]
        self.launch_order.append(b)
        view = self.flat_g[s:e]
        if self._comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                self._handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero_grad(self):
        self.flat_g.zero_()
        self._pending = [b[2] for b in self.buckets]
        self._uses = [0] * len(self.params)
        self.launch_order = []

    def sync_gradients(self):
        """Waits for the bucket allreduces (launching any bucket whose hooks did not all fire, e.g. unused params)."""
        if self.world == 1:
            return
        for b, pend in enumerate(self._pending):
            if pend > 0:
                self._pending[b] = 0
                self._launch_bucket(b)
        for h in self._handles:
            h.wait()
        self._handles = []
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        if self._expected is None and any(self._uses):
            self._expected = list(self._uses)  # observability: in-place writes per parameter per step


class DataParallelTrainer:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, bucket_mb=25.0,
                 process_group=None, exclude=(), optimizer="adam", momentum=0.9):
        """optimizer: "adam" (train_cls_voxel.py:195) or "sgd" (momentum SGD of train_cls.py:91 / train_partseg.py:95)."""
        from . import _lib as L
        self._L = L
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        if optimizer not in ("adam", "sgd"):
            raise ValueError("optimizer must be 'adam' or 'sgd'")
        self.optimizer, self.momentum = optimizer, momentum
        exclude = set(exclude)
        for n, p in model.named_parameters():
            if n in exclude:
                p.requires_grad_(False)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if named[0][1].device.type != "cuda":
            raise RuntimeError("DataParallelTrainer needs the model on a CUDA device (no CPU path)")
        self.flat = FlatGradBuckets(named, bucket_mb=bucket_mb, process_group=process_group)
        f = self.flat
        self.world = f.world
        dev = f.device
        self.flat_m = torch.zeros(f.total, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(f.total, device=dev, dtype=torch.float32)
        self.flat_s = torch.zeros(f.total, device=dev, dtype=torch.bfloat16)
        for p, o in zip(f.params, f.offsets):
            n = p.numel()
            p._s3d_shadow = self.flat_s[o:o + n].view(p.shape[0], -1) if p.dim() >= 2 else self.flat_s[o:o + n]
        L.cast_bf16(f.flat_p, out=self.flat_s)
        f.enable_sinks()
        self.step_count = 0
        self.step_t = torch.zeros(1, device=dev, dtype=torch.int32)  # device-side counter (CUDA-graph replays)

    def zero_grad(self):
        self.flat.zero_grad()

    def sync_gradients(self):
        self.flat.sync_gradients()

    def optimizer_step(self):
        self.step_count += 1
        self.step_t += 1
        f = self.flat
        if self.optimizer == "sgd":
            self._L.sgd_momentum_step(f.flat_p, f.flat_g, self.flat_m, self.flat_s, self.lr, self.momentum,
                                      self.weight_decay, self.step_count, grad_scale=1.0 / self.world,
                                      step_tensor=self.step_t)
            return
        self._L.adam_step(f.flat_p, f.flat_g, self.flat_m, self.flat_v, self.flat_s, self.lr, self.betas[0],
                          self.betas[1], self.eps, self.weight_decay, self.step_count, grad_scale=1.0 / self.world,
                          step_tensor=self.step_t)

    def step(self, x, y, loss_fn):
        """forward + backward + gradient allreduce + Adam. Returns the (device) loss tensor."""
        self.zero_grad()
        loss = loss_fn(self.model(x), y)
        loss.backward()
        self.sync_gradients()
        self.optimizer_step()
        return loss
