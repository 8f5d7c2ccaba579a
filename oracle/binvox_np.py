"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's binvox reader / writer.

Follows utils/binvox_rw.py of the reference: read_header (:105-115), read_as_3d_array (:117-151: np.repeat(values,
counts) -> reshape(dims) in x-z-y order -> transpose(0, 2, 1)), write (:231-283: run-length state machine, runs capped
at 255). The reference file uses the removed aliases np.bool / np.int and does not import on numpy >= 1.24; this
restatement is pinned against the real functions (run with those aliases patched back) by
tests/golden/make_golden_binvox.py -> tests/golden/binvox.pt, and tests/test_oracle.py checks it against that fixture.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
import numpy as np


def read_header(buf: bytes):
    """-> (dims, translate, scale, payload_offset). utils/binvox_rw.py:105-115."""
    lines, pos = [], 0
    for _ in range(5):
        end = buf.index(b"\n", pos)
        lines.append(buf[pos:end].strip())
        pos = end + 1
    if not lines[0].startswith(b"#binvox"):
        raise IOError("Not a binvox file")
    dims = list(map(int, lines[1].split(b" ")[1:]))
    translate = list(map(float, lines[2].split(b" ")[1:]))
    scale = list(map(float, lines[3].split(b" ")[1:]))[0]
    return dims, translate, scale, pos


def read_as_3d_array(buf: bytes, fix_coords: bool = True):
    """-> bool array [dims] (x, y, z) (or (x, z, y) with fix_coords=False). utils/binvox_rw.py:117-151."""
    dims, _, _, pos = read_header(buf)
    raw = np.frombuffer(buf, dtype=np.uint8, offset=pos)
    values, counts = raw[::2], raw[1::2]
    data = np.repeat(values, counts).astype(bool).reshape(dims)
    if fix_coords:
        data = np.transpose(data, (0, 2, 1))
    return data


def write(dense_xyz: np.ndarray, translate=(0.0, 0.0, 0.0), scale=1.0) -> bytes:
    """Dense (x, y, z) occupancy -> binvox file bytes. utils/binvox_rw.py:231-283 (axis_order 'xyz' branch)."""
    dims = dense_xyz.shape
    head = ("#binvox 1\n" + "dim " + " ".join(map(str, dims)) + "\n" + "translate " + " ".join(map(str, translate)) +
            "\n" + "scale " + str(scale) + "\ndata\n").encode("ascii")
    flat = np.transpose(dense_xyz.astype(bool), (0, 2, 1)).reshape(-1).astype(np.uint8)
    change = np.flatnonzero(np.diff(flat)) + 1
    starts = np.concatenate(([0], change))
    lengths = np.diff(np.concatenate((starts, [flat.size])))
    states = flat[starts]
    full, rem = lengths // 255, lengths % 255  # the state machine dumps every 255 voxels and keeps counting
    n_out = full + (rem > 0)
    vals = np.repeat(states, n_out)
    cnts = np.full(vals.size, 255, dtype=np.uint8)
    last = np.cumsum(n_out) - 1
    cnts[last[rem > 0]] = rem[rem > 0].astype(np.uint8)
    pairs = np.empty(2 * vals.size, dtype=np.uint8)
    pairs[0::2], pairs[1::2] = vals, cnts
    return head + pairs.tobytes()
