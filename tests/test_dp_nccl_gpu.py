"""GPU (needs >= 2 devices, skipped otherwise): the NCCL data-parallel step. Each rank's gradient-sink / bucketed allreduce
result must equal the sum over ranks of the gradients plain autograd produces for that rank's batch, and the parameters
after one fused Adam step must be identical on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import s3d_oracle as O
    from simple3d_former_b200.dp import DataParallelTrainer
    from simple3d_former_b200.embed_layer_3d_modality import VoxelEmbed_no_average
    from simple3d_former_b200.models import Feature3D_ViT2D_V2
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sd = O.init_voxel_state_dict("deit_base_patch16_224", 9, 4, 55, "group_embed", seed=9)

    def build():
        m = Feature3D_ViT2D_V2(embed_layer=VoxelEmbed_no_average(36, 9, 4, embed_dim=768), n_classes=55,
                               transformer_backbone="deit_base_patch16_224", pretrained=False, pos_embedding="group_embed")
        m.load_state_dict(sd, strict=False)
        m.freeze_image_branch()
        m.group_embed.dropout_p = 0.0  # this test compares gradients across separate forwards: no stochastic masks
        return m.to(dev).train()

    x, y = O.synthetic_voxels(3, 36, seed=100 + rank, n_classes=55)  # a different batch on every rank
    x, y = x.to(dev), y.to(dev)
    plain = build()
    F.cross_entropy(plain(x), y).backward()
    want = {n: p.grad.detach().clone() for n, p in plain.named_parameters() if p.grad is not None}
    for g in want.values():
        dist.all_reduce(g)  # expected: sum over ranks of the local gradients
    model = build()
    trainer = DataParallelTrainer(model, lr=1e-3, bucket_mb=4.0)
    worst = 0.0
    for step in range(2):  # step 0 learns the per-parameter write counts, step 1 uses hook-driven bucket launches
        trainer.zero_grad()
        F.cross_entropy(model(x), y).backward()
        trainer.sync_gradients()
        torch.cuda.synchronize()
        for n, p in model.named_parameters():
            if p.requires_grad:
                scale = want[n].abs().max().item() + 1e-12
                worst = max(worst, (p.grad - want[n]).abs().max().item() / scale)
    order = list(trainer.flat.launch_order)
    trainer.optimizer_step()
    torch.cuda.synchronize()
    chk = trainer.flat.flat_p.double().sum().item()
    q.put((rank, worst, chk, order, len(trainer.flat.buckets)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_and_step():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, w0, c0, order0, nb), (_, w1, c1, order1, _) = res
    assert w0 <= 2e-3 and w1 <= 2e-3, (w0, w1)
    assert c0 == c1, "parameters diverged between ranks after the fused Adam step"
    assert sorted(order0) == list(range(nb)) and nb >= 4
    assert order0 == order1
