#!/usr/bin/env python
"""Which stage of the point path contributes how much of the logit error against the reference's golden logits?
Runs the cfg4 / cfg5 fixtures with the fused set-abstraction and transition-up paths switched on and off."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import s3d_oracle as O  # noqa: E402
from simple3d_former_b200 import models as M  # noqa: E402
from simple3d_former_b200 import pointnet_util as P  # noqa: E402

dev = torch.device("cuda:0")
for name in ("cfg4_point_cls_tiny1024", "cfg5_point_seg_tiny2048"):
    fix = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)
    mc = types.SimpleNamespace(nblocks=4, nneighbor=16, transformer_backbone=fix["backbone"], pretrained=False,
                               head="Linear", transformer_dim=512)
    pc = types.SimpleNamespace(num_point=fix["N"], num_class=fix["n_classes"], input_dim=fix["input_dim"], model=mc)
    for mode in ("train", "eval"):
        for sa_f, tu_f in ((True, True), (False, True), (True, False), (False, False)):
            P.PointNetSetAbstraction.fused = sa_f
            M.TransitionUp.fused = tu_f
            model = (M.PointTransformerSeg if fix["seg"] else M.PointTransformerCls)(pc)
            model.load_state_dict(O.init_point_state_dict(fix["backbone"], fix["input_dim"], fix["n_classes"],
                                                          seed=fix["weight_seed"]), strict=False)
            model = model.to(dev).train(mode == "train")
            model.set_fps_starts([s.to(dev) for s in fix["fps_starts"]])
            x, _ = O.synthetic_points(fix["B"], fix["N"], extra=fix["input_dim"] - 3, seed=fix["input_seed"],
                                      n_classes=fix["n_classes"])
            with torch.no_grad():
                logits = model(x.to(dev)).cpu()
            ref = fix[mode]["logits"]
            err = (logits - ref).abs()
            print(f"{name} {mode} sa_fused={sa_f} tu_fused={tu_f}: max|err|={err.max():.5f} mean|err|={err.mean():.6f} "
                  f"max|logit|={ref.abs().max():.3f}")
