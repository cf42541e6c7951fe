"""Gadget-2 snapshot reader / writer of the host library (second half) against the reference's conventions
(1_Indexing/src/snapshot.c:5-22,211-293,397-503): header layout, float32 blocks, a^(3/2) velocity unit, slab reads; and
against the reference's own demo file where /root/reference is mounted (the committed golden positions were extracted
from that file by an independent reader, oracle/refrun.py)."""
import os
import struct

import numpy as np
import pytest
from conftest import ROOT

from p2p_b200 import snapshot

DEMO = "/root/reference/1_Indexing/demo/ic_lcdm.gdt2"


def test_write_then_read_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    n, box, z = 5000, 1.0e5, 49.0
    pos = rng.uniform(0, box, (n, 3)).astype(np.float32).astype(np.float64)
    vel = rng.normal(0, 50.0, (n, 3))
    path = tmp_path / "snap.gdt2"
    snapshot.write(path, pos, vel, box, 211.75, z, npart_total=4 * n, omega0=0.3, omega_lambda=0.7, hubble=0.7)
    raw = path.read_bytes()
    # the layout the reference's reader walks: [256][header][256] [12 n][positions][12 n] [12 n][velocities][12 n]
    assert len(raw) == 4 + 256 + 4 + 2 * (4 + 12 * n + 4)
    assert struct.unpack_from("<i", raw, 0)[0] == 256 and struct.unpack_from("<i", raw, 260)[0] == 256
    assert struct.unpack_from("<6i", raw, 4) == (0, n, 0, 0, 0, 0)
    assert struct.unpack_from("<i", raw, 264)[0] == 12 * n
    info = snapshot.header(path)
    assert (info.nfile, info.npart[1], info.npart_total[1], info.num_files) == (n, n, 4 * n, 1)
    assert (info.box, info.redshift, info.mass[1], info.omega0, info.hubble) == (box, z, 211.75, 0.3, 0.7)
    assert abs(info.time - 1.0 / (1.0 + z)) < 1e-15
    p, v, _ = snapshot.read(path)
    assert np.array_equal(p, pos)                                        # float32 values survive exactly
    unit = (1.0 / (1.0 + z)) ** 1.5
    on_disk = (vel.astype(np.float32).astype(np.float64) / unit).astype(np.float32)      # the reference's narrowing order
    assert np.array_equal(v, on_disk.astype(np.float64) * unit)
    assert np.abs(v - vel).max() < 1e-5 * np.abs(vel).max()
    # a rank's slab
    p2, v2, _ = snapshot.read(path, 1234, 777)
    assert np.array_equal(p2, pos[1234:2011]) and np.array_equal(v2, v[1234:2011])
    with pytest.raises(IOError):
        snapshot.read(path, n - 5, 10)
    with pytest.raises(IOError):
        snapshot.header(tmp_path / "missing.gdt2")


@pytest.mark.skipif(not os.path.isfile(DEMO), reason="reference demo file not mounted")
def test_reads_the_reference_demo_file():
    want = np.load(os.path.join(ROOT, "tests", "golden", "demo_lcdm_pos_f32.npy"))
    info = snapshot.header(DEMO)
    assert info.nfile == len(want) == 32768 and info.box == 100000.0
    pos, vel, _ = snapshot.read(DEMO)
    assert np.array_equal(pos, want.astype(np.float64))
    assert np.isfinite(vel).all() and np.abs(vel).max() > 0
    lo, hi = 32768 * 3 // 8, 32768 * 4 // 8                               # the slab of rank 3 of 8
    assert np.array_equal(snapshot.read(DEMO, lo, hi - lo, velocities=False)[0], want[lo:hi].astype(np.float64))
