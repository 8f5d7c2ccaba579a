"""Runs the same forward+backward twice (and once under the trainer) and reports bitwise / relative differences."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch, torch.nn.functional as F
import s3d_oracle as O
from simple3d_former_b200.embed_layer_3d_modality import VoxelEmbed_no_average
from simple3d_former_b200.models import Feature3D_ViT2D_V2
from simple3d_former_b200.dp import DataParallelTrainer

dev = torch.device("cuda:0")
sd = O.init_voxel_state_dict("deit_base_patch16_224", 9, 4, 55, "group_embed", seed=9)
x, y = O.synthetic_voxels(3, 36, seed=9, n_classes=55)
x, y = x.to(dev), y.to(dev)

def build():
    m = Feature3D_ViT2D_V2(embed_layer=VoxelEmbed_no_average(36, 9, 4, embed_dim=768), n_classes=55,
                           transformer_backbone="deit_base_patch16_224", pretrained=False, pos_embedding="group_embed")
    m.load_state_dict(sd, strict=False)
    m.freeze_image_branch()
    m.group_embed.dropout_p = float(os.environ.get("S3D_CHECK_DROPOUT", "0"))  # default: deterministic arithmetic
    return m.to(dev).train()

def run(m):
    for p in m.parameters():
        if p.grad is not None and not hasattr(p, "_s3d_grad_sink"):
            p.grad = None
    logits = m(x)
    F.cross_entropy(logits, y).backward()
    torch.cuda.synchronize()
    return logits.detach().clone(), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}

m = build()
seed0 = m.group_embed._drop_seed.clone()
l1, g1 = run(m)
m.group_embed._drop_seed.copy_(seed0)  # with dropout on, the same seed must reproduce the same masks
l2, g2 = run(m)
print("same model, run twice: logits bitwise equal:", torch.equal(l1, l2), "max diff", (l1 - l2).abs().max().item())
worst = sorted(((g1[n] - g2[n]).abs().max().item() / (g1[n].abs().max().item() + 1e-12), n) for n in g1)[-5:]
print("  worst grad rel diffs:", worst)
m2 = build()
l3, g3 = run(m2)
print("second model instance: logits bitwise equal:", torch.equal(l1, l3), "max diff", (l1 - l3).abs().max().item())
m3 = build()
tr = DataParallelTrainer(m3, lr=1e-3)
tr.zero_grad()
l4 = m3(x)
F.cross_entropy(l4, y).backward()
tr.sync_gradients()
torch.cuda.synchronize()
print("trainer: logits bitwise equal:", torch.equal(l1, l4.detach()), "max diff", (l1 - l4.detach()).abs().max().item())
g4 = {n: p.grad for n, p in m3.named_parameters() if p.requires_grad}
worst = sorted(((g1[n] - g4[n]).abs().max().item() / (g1[n].abs().max().item() + 1e-12), n) for n in g1)[-8:]
print("  worst grad rel diffs vs autograd path:", worst)
