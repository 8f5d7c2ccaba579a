"""binvox files -> occupancy grids on the GPU, with the reference's function surface (utils/binvox_rw.py).

The reference reads a file with numpy on the host (read_as_3d_array, :117-151), widens it to int32 [1,V,V,V]
(data/modelnet40.py:40-41) and ships 8 MiB per 128^3 model over PCIe. Here only the header is parsed on the host; the
run-length payload (tens of KB per model) is copied to the device as is and expanded there by s3d_binvox_scan /
s3d_binvox_expand into the uint8 (or int32 / float32) grid that VoxelEmbed consumes. Cubic grids only (every dataset
the reference trains on is cubic); no CPU fallback."""
from __future__ import annotations

import io

import torch

from . import _lib as L


class Voxels:
    """Same fields as the reference's Voxels (utils/binvox_rw.py:66-103); `data` is a CUDA tensor [V, V, V]."""

    def __init__(self, data, dims, translate, scale, axis_order):
        assert axis_order in ("xzy", "xyz")
        self.data, self.dims, self.translate, self.scale, self.axis_order = data, dims, translate, scale, axis_order


def read_header(fp):
    """utils/binvox_rw.py:105-115."""
    line = fp.readline().strip()
    if not line.startswith(b"#binvox"):
        raise IOError("Not a binvox file")
    dims = list(map(int, fp.readline().strip().split(b" ")[1:]))
    translate = list(map(float, fp.readline().strip().split(b" ")[1:]))
    scale = list(map(float, fp.readline().strip().split(b" ")[1:]))[0]
    fp.readline()
    return dims, translate, scale


def _payload(src):
    """file path / bytes / binary file object -> (dims, translate, scale, payload bytes)."""
    if isinstance(src, (bytes, bytearray, memoryview)):
        fp = io.BytesIO(bytes(src))
    elif isinstance(src, str):
        fp = open(src, "rb")
    else:
        fp = src
    try:
        dims, translate, scale = read_header(fp)
        data = fp.read()
    finally:
        if isinstance(src, str):
            fp.close()
    if len(dims) != 3 or not (dims[0] == dims[1] == dims[2]):
        raise ValueError(f"only cubic binvox grids are supported, got dims {dims}")
    if len(data) % 2:
        raise ValueError("binvox payload is not a sequence of (value, count) byte pairs")
    return dims, translate, scale, data


def load_voxel_batch(sources, device="cuda", dtype=torch.uint8, fix_coords=True, check=True):
    """B binvox models (paths, bytes or binary file objects) -> occupancy grid [B, 1, V, V, V] on `device`.

    Equivalent of stacking `np.int32(binvox_rw.read_as_3d_array(f).data)[np.newaxis]` over a batch
    (data/modelnet40.py:35-45) and `.to(device)`; dtype uint8 (default, what the patchify kernel reads), int32 or
    float32. check=True verifies on the host that every stream encodes exactly V^3 voxels (the reference's reshape
    raises otherwise); it costs one device-to-host read of B integers."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("simple3d_former_b200.binvox_rw expands binvox payloads on the GPU (no CPU fallback)")
    metas = [_payload(s) for s in sources]
    V = metas[0][0][0]
    if any(m[0][0] != V for m in metas):
        raise ValueError("all models of a batch must have the same grid size")
    # every model's pairs start on a 16-byte boundary (vector loads in the scan); the padding is (value 0, count 0)
    # pairs, i.e. empty runs
    chunks, offs = [], [0]
    for m in metas:
        pad = (-len(m[3])) % 16
        chunks.append(m[3] + b"\0" * pad)
        offs.append(offs[-1] + len(m[3]) + pad)
    host = torch.frombuffer(bytearray(b"".join(chunks)), dtype=torch.uint8).pin_memory()
    payload = host.to(dev, non_blocking=True)
    offsets = torch.tensor(offs, dtype=torch.long).to(dev)
    grid, totals = L.binvox_expand(payload, offsets, V, out_dtype=dtype, fix_coords=fix_coords)
    if check:
        bad = (totals != V ** 3).nonzero().flatten().tolist()
        if bad:
            raise ValueError(f"binvox stream of model(s) {bad} does not encode {V}^3 voxels")
    return grid


def read_as_3d_array(fp, fix_coords=True, device="cuda", dtype=torch.uint8):
    """utils/binvox_rw.py:117-151 with the dense grid produced on the GPU."""
    dims, translate, scale, data = _payload(fp)
    grid = load_voxel_batch([b"#binvox 1\ndim %d %d %d\ntranslate 0 0 0\nscale 1\ndata\n" % tuple(dims) + data],
                            device=device, dtype=dtype, fix_coords=fix_coords)
    return Voxels(grid[0, 0], dims, translate, scale, "xyz" if fix_coords else "xzy")
