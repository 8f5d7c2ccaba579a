"""CPU, world_size 2, gloo: the host-side data-parallel logic (flat layout, buckets, hook-driven allreduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 32), nn.ReLU(), nn.Linear(32, 4))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simple3d_former_b200.dp import FlatGradBuckets
    model = _model()
    if rank == 1:  # ranks start from different weights: the constructor broadcast must fix that
        for p in model.parameters():
            p.data.add_(1.0)
    flat = FlatGradBuckets(model.named_parameters(), bucket_mb=0.0005)  # ~128 floats per bucket -> several buckets
    torch.manual_seed(100)
    x = torch.randn(8, 16)
    y = torch.randn(8, 4)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    flat.zero_grad()
    loss = ((model(xs) - ys) ** 2).sum()
    loss.backward()
    flat.sync_gradients()
    q.put((rank, flat.flat_p.clone(), flat.flat_g.clone(), len(flat.buckets), list(flat.launch_order)))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_buckets_allreduce_matches_full_batch():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, p0, g0, nb0, order0), (_, p1, g1, nb1, order1) = res
    assert nb0 == nb1 and nb0 >= 3
    assert torch.equal(p0, p1), "constructor broadcast must make the weights identical"
    assert torch.equal(g0, g1), "summed gradients must be identical on every rank"
    assert sorted(order0) == list(range(nb0)) and order0 == order1
    assert order0[0] == 0, "the bucket holding the last layers' gradients is reduced first (backward order)"
    # reference: single process over the concatenated batch (sum-loss => summed gradients are equal)
    from simple3d_former_b200.dp import FlatGradBuckets
    model = _model()
    flat = FlatGradBuckets(model.named_parameters(), bucket_mb=0.0005)
    torch.manual_seed(100)
    x = torch.randn(8, 16)
    y = torch.randn(8, 4)
    flat.zero_grad()
    ((model(x) - y) ** 2).sum().backward()
    assert torch.allclose(flat.flat_g, g0, atol=1e-5)
    assert torch.equal(flat.flat_p, p0)


def test_flat_views_alias_parameters():
    from simple3d_former_b200.dp import FlatGradBuckets
    model = _model()
    want = {n: p.detach().clone() for n, p in model.named_parameters()}
    flat = FlatGradBuckets(model.named_parameters())
    for n, p in model.named_parameters():
        assert torch.equal(p, want[n])
        assert p.data_ptr() >= flat.flat_p.data_ptr()
        assert p.data_ptr() < flat.flat_p.data_ptr() + flat.flat_p.numel() * 4
    flat.flat_p.mul_(2.0)
    for n, p in model.named_parameters():
        assert torch.equal(p, want[n] * 2.0)
    assert flat.world == 1 and flat.total % 64 == 0
