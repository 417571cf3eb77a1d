"""Writes tests/golden/reference_inputs.bin: the INPUT arrays of golden_v1.npz in the flat format that
make_reference_fixtures.rs reads (b"AUSFIX01", u64 n, then per array: u64 name_len, name, u64 ndim, dims, f64 data), so
that the real crate analyses bit-identical samples.  See make_reference_fixtures.rs for the whole procedure."""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INPUTS = ("c1_x", "mp_x", "c2_x0", "c2_x1", "c3_x", "c3_lens", "c4_x", "an_mag")


def write_arrays(path, arrays):
    with open(path, "wb") as f:
        f.write(b"AUSFIX01")
        f.write(struct.pack("<Q", len(arrays)))
        for name in sorted(arrays):
            a = np.ascontiguousarray(arrays[name], dtype="<f8")
            nb = name.encode()
            f.write(struct.pack("<Q", len(nb)) + nb)
            f.write(struct.pack("<Q", a.ndim) + struct.pack(f"<{a.ndim}Q", *a.shape))
            f.write(a.tobytes())


def read_arrays(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"AUSFIX01", "not an AUSFIX01 file"
        (n,) = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            (ln,) = struct.unpack("<Q", f.read(8))
            name = f.read(ln).decode()
            (nd,) = struct.unpack("<Q", f.read(8))
            dims = struct.unpack(f"<{nd}Q", f.read(8 * nd))
            cnt = int(np.prod(dims)) if nd else 1
            out[name] = np.frombuffer(f.read(8 * cnt), dtype="<f8").reshape(dims).copy()
    return out


if __name__ == "__main__":
    g = np.load(os.path.join(HERE, "golden_v1.npz"))
    dst = os.path.join(HERE, "reference_inputs.bin")
    write_arrays(dst, {k: g[k] for k in INPUTS})
    print("wrote", dst, os.path.getsize(dst) // 1024, "KiB")
