"""Run oracle/_ref/ref_lists (the reference's own tree build + walks, see ref_harness/) and parse
its record stream.  TEST INFRASTRUCTURE ONLY: imported by tests/ and by tests/golden/make_golden.py,
never by the product package."""
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "ref_lists")
_DT = {0: np.int32, 1: np.float64, 2: np.int64}


def available():
    return os.path.isfile(REF_BIN) and os.access(REF_BIN, os.X_OK)


def parse(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    o = 0
    while o < len(buf):
        (nl,) = struct.unpack_from("<I", buf, o); o += 4
        name = buf[o:o + nl].decode(); o += nl
        dt, cnt = struct.unpack_from("<IQ", buf, o); o += 12
        dtype = np.dtype(_DT[dt])
        arr = np.frombuffer(buf, dtype=dtype, count=cnt, offset=o).copy(); o += cnt * dtype.itemsize
        out[name] = arr
    return out


def run(pos, box, maxleaf, nside, theta=0.4, do_ext=False, nproc=1, timeout=600):
    """pos: (N,3) float64.  Returns a list (one per rank) of dicts of numpy arrays."""
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    with tempfile.TemporaryDirectory() as td:
        pf = os.path.join(td, "pos.f64")
        pos.tofile(pf)
        prefix = os.path.join(td, "out")
        cmd = [REF_BIN, pf, str(pos.shape[0]), repr(float(box)), str(int(maxleaf)), str(int(nside)),
               repr(float(theta)), "1" if do_ext else "0", str(int(nproc)), prefix]
        subprocess.run(cmd, check=True, timeout=timeout, stdout=subprocess.DEVNULL)
        return [parse(f"{prefix}.rank{r}") for r in range(nproc)]


def read_gadget2_positions(path):
    """Positions block of a Gadget-2 format-1 snapshot (1_Indexing/src/snapshot.c:5-22,243-259):
    [i32 256][256-byte header][i32 256][i32 nbytes][N*3 float32][i32 nbytes]..."""
    with open(path, "rb") as f:
        (m,) = struct.unpack("<i", f.read(4)); assert m == 256
        hdr = f.read(256); f.read(4)
        npart = struct.unpack_from("<6i", hdr, 0)
        mass = struct.unpack_from("<6d", hdr, 24)
        box = struct.unpack_from("<d", hdr, 24 + 48 + 8 + 8 + 4 + 4 + 24 + 4 + 4)[0]
        (nb,) = struct.unpack("<i", f.read(4))
        n = sum(npart)
        assert nb == n * 12, (nb, n)
        pos = np.frombuffer(f.read(nb), dtype="<f4").reshape(n, 3).copy()
    return pos, box, mass, npart
