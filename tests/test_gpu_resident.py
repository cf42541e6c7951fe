"""Device-resident stepping (integrator part): positions, velocities and ids stay on the device between
steps.  Mirror on the host: the same kick / drift / wrap arithmetic in numpy (fp64) around the one-step device path,
carrying the arrays in the order every tree build leaves them, as the reference's part[] does.  Bit-exact equality of
positions, velocities and ids after several steps: same order -> same tree -> same lists -> same FP32 sums."""
import numpy as np
import pytest
from conftest import DEMO_BOX, DEMO_MASS, DEMO_NSIDE, THETA

import oracle
import p2p_b200

pytestmark = pytest.mark.gpu


def test_resident_steps_equal_host_mirror(demo_pos):
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    bdl, bdr = np.zeros(3), np.full(3, DEMO_BOX)
    n = len(demo_pos)
    rng = np.random.default_rng(5)
    vel0 = rng.normal(0.0, 1.0, (n, 3))
    nstep, maxleaf = 3, 16

    a = p2p_b200.P2PContext(0)
    b = p2p_b200.P2PContext(0)
    try:
        for c in (a, b):
            c.set_physics(DEMO_MASS, eps, rs)
            c.set_box([0.0, 0.0, 0.0], DEMO_BOX)
        # scale the steps so that particles move ~0.2 cells per step, some across the periodic boundary
        b.tree_build(demo_pos, maxleaf, bdl, bdr, 0)
        b.clear_tasks(); b.tree_walk(THETA, rcut, DEMO_BOX, 0.5 * (bdr + bdl), bdr - bdl); b.build_csr(); b.compute()
        amean = np.linalg.norm(b.download_acc(), axis=1).mean()
        cell = DEMO_BOX / DEMO_NSIDE
        dkh = 0.5 * cell / amean                 # velocities of ~0.5 cell per unit time after one kick
        dd = 0.4

        # ---- resident path
        a.resident_load(demo_pos, vel0 * 0.2 * cell)
        for _ in range(nstep):
            a.resident_forces(maxleaf, bdl, bdr, THETA, rcut, DEMO_BOX)
            a.resident_kick(dkh)
            a.resident_drift(dd, DEMO_BOX)
        pa, va, ia = a.resident_download()

        # ---- host mirror around the one-step device path
        pos, vel, ids = demo_pos.copy(), vel0 * 0.2 * cell, np.arange(n, dtype=np.int64)
        wrapped = 0
        for _ in range(nstep):
            b.tree_build(pos, maxleaf, bdl, bdr, 0)
            b.clear_tasks(); b.tree_walk(THETA, rcut, DEMO_BOX, 0.5 * (bdr + bdl), bdr - bdl); b.build_csr(); b.compute()
            acc = b.download_acc()                                  # tree order
            D = b.tree_download()
            pos, vel, ids = D["pos"], vel[D["perm"]], ids[D["perm"]]
            vel = vel + acc * dkh
            pos = pos + vel * dd
            for k in range(3):
                col = pos[:, k]
                lo, hi = col < 0.0, col >= DEMO_BOX
                wrapped += int(lo.sum() + hi.sum())
                while (col < 0.0).any():
                    col[col < 0.0] += DEMO_BOX
                while (col >= DEMO_BOX).any():
                    col[col >= DEMO_BOX] -= DEMO_BOX
        assert wrapped > 50                                         # the periodic wrap was exercised
        assert np.array_equal(ia, ids) and sorted(ia.tolist()) == list(range(n))
        assert np.array_equal(va, vel)
        assert np.array_equal(pa, pos)
        assert np.abs(pa - demo_pos[ia]).max() > 0.1 * cell         # and the particles really moved
    finally:
        a.close()
        b.close()
