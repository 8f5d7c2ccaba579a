import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "oracle"))
import torch, torch.distributed as dist, torch.nn.functional as F
import s3d_oracle as O
from simple3d_former_b200.dp import DataParallelTrainer
from simple3d_former_b200.embed_layer_3d_modality import VoxelEmbed_no_average
from simple3d_former_b200.models import Feature3D_ViT2D_V2
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
sd = O.init_voxel_state_dict("deit_base_patch16_224", 9, 4, 55, "group_embed", seed=9)
def build():
    m = Feature3D_ViT2D_V2(embed_layer=VoxelEmbed_no_average(36, 9, 4, embed_dim=768), n_classes=55, transformer_backbone="deit_base_patch16_224", pretrained=False, pos_embedding="group_embed")
    m.load_state_dict(sd, strict=False); m.freeze_image_branch(); return m.to(dev).train()
x, y = O.synthetic_voxels(3, 36, seed=100 + rank, n_classes=55); x, y = x.to(dev), y.to(dev)
plain = build(); F.cross_entropy(plain(x), y).backward()
local = {n: p.grad.detach().clone() for n, p in plain.named_parameters() if p.grad is not None}
want = {n: g.clone() for n, g in local.items()}
for g in want.values(): dist.all_reduce(g)
model = build(); tr = DataParallelTrainer(model, lr=1e-3, bucket_mb=4.0)
for step in range(2):
    tr.zero_grad(); F.cross_entropy(model(x), y).backward(); tr.sync_gradients(); torch.cuda.synchronize()
    rows = []
    for n, p in model.named_parameters():
        if p.requires_grad:
            sc = want[n].abs().max().item() + 1e-12
            rows.append(((p.grad - want[n]).abs().max().item() / sc, (p.grad - local[n]).abs().max().item() / sc, n))
    rows.sort(reverse=True)
    if rank == 0:
        print(f"step {step}: buckets {len(tr.flat.buckets)} launch_order {tr.flat.launch_order[:12]}... n_bad(>2e-3)={sum(r[0] > 2e-3 for r in rows)} of {len(rows)}")
        for r in rows[:6]: print("   err_vs_sum %.3e  err_vs_local %.3e  %s" % r)
        fl = tr.flat
        idx = {n: i for i, n in enumerate(fl.names)}
        for r in rows:
            if r[0] > 2e-3:
                i = idx[r[2]]
                print("   BAD %-40s bucket %2d uses %d expected %s err_sum %.2e err_local %.2e" % (r[2], fl._bucket_of[i], fl._uses[i], None if fl._expected is None else fl._expected[i], r[0], r[1]))
        print("   bucket sizes (params):", [b[2] for b in fl.buckets])
        print("   pending after sync:", fl._pending)
dist.barrier(); dist.destroy_process_group()
