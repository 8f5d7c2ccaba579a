"""Bitwise run-to-run reproducibility of the attention kernels (no atomics anywhere in them): every shape is run REPS times
on the same inputs and compared with the first result. A data race in a kernel shows up here long before it moves a
tolerance-based parity test."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from simple3d_former_b200 import _lib as L  # noqa: E402

REPS = int(os.environ.get("REPS", "12"))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
bad = 0
for (B, N, H, dh, seqfirst, drop) in [(128, 257, 3, 64, False, False), (64, 197, 3, 256, False, False), (32, 513, 3, 64, False, False),
                                      (15, 12544, 4, 192, True, True), (15, 12544, 4, 192, True, False), (4, 700, 4, 96, True, True),
                                      (1000, 15, 3, 256, False, False), (64, 26, 6, 64, False, False)]:
    E = H * dh
    qkv = (torch.randn(B * N, 3 * E, device=dev, generator=g) * 0.5).bfloat16()
    dout = torch.randn(B * N, E, device=dev, generator=g).bfloat16()
    if seqfirst:
        qs, os_ = (3 * E, dh, B * 3 * E), (E, dh, B * E)
    else:
        qs, os_ = (N * 3 * E, dh, 3 * E), (N * E, dh, E)
    seed = torch.tensor([77], dtype=torch.int32, device=dev)
    kw = dict(drop_seed=seed, drop_site=1, drop_p=0.1) if drop else {}
    b = qkv.data_ptr()
    ref = None
    for rep in range(REPS):
        o = torch.zeros(B * N, E, device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(B, H, N, device=dev)
        L.attn_fwd(b, b + 2 * E, b + 4 * E, o, lse, B, H, N, dh, qs, os_, dh ** -0.5, **kw)
        dqkv = torch.zeros_like(qkv)
        delta = torch.zeros_like(lse)
        db = dqkv.data_ptr()
        L.attn_bwd(b, b + 2 * E, b + 4 * E, o, dout, lse, delta, db, db + 2 * E, db + 4 * E, B, H, N, dh, qs, os_, dh ** -0.5, **kw)
        torch.cuda.synchronize()
        cur = (o, lse, dqkv)
        if ref is None:
            ref = cur
            assert torch.isfinite(o.float()).all() and torch.isfinite(dqkv.float()).all(), "non-finite output"
        else:
            for name, a, c in zip(("out", "lse", "dqkv"), ref, cur):
                if not torch.equal(a, c):
                    bad += 1
                    print(f"  [FAIL] B{B} N{N} H{H} dh{dh} drop{int(drop)} rep {rep}: {name} differs, max abs "
                          f"{(a.float() - c.float()).abs().max().item():.3e}", flush=True)
    print(f"  [{'PASS' if bad == 0 else 'SEEN FAILURES'}] B{B} N{N} H{H} dh{dh} seqfirst{int(seqfirst)} drop{int(drop)}: {REPS} runs", flush=True)
print("determinism:", "ok" if bad == 0 else f"{bad} mismatches")
sys.exit(1 if bad else 0)
