"""FPS kernel alone at the cfg4 / cfg5 sizes (time per selected point)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple3d_former_b200 import _lib as L
for (B, N, S) in [(128, 1024, 256), (32, 2048, 512), (128, 256, 64)]:
    xyz = torch.rand(B, N, 3, device="cuda") * 2 - 1
    start = torch.zeros(B, dtype=torch.long, device="cuda")
    out = torch.empty(B, S, dtype=torch.long, device="cuda")
    fn = lambda: L.call("s3d_fps", xyz.data_ptr(), start.data_ptr(), out.data_ptr(), B, N, S, L.stream())
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"B{B} N{N} npoint{S}: {ms * 1e3:.1f} us  {ms * 1e6 / S:.0f} ns per point")
