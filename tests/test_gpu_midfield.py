"""Mid-field on the device (P2M / M2M / M2L / L2L / L2P, csrc/midfield.cuh) against the REFERENCE ITSELF:
oracle/_ref/ref_lists runs the reference's unmodified fmm_prepare / fmm_task / fmm_ext with its own CPU operators
(1_Indexing/src/operator.c) while the P2P stubs return zeros, so its part[].acc is the mid-field alone.

With the reference's own parameters (theta 0.4, r_cut 4.5 r_s, MAXLEAF >= 8) the acceptance criterion never hands a
pair to the expansions (an accepted pair needs w_i + w_j < theta d with d <= r_cut = 5.6 cells): its M2L list is EMPTY
and the mid-field is exactly zero -- checked below.  The operators are therefore exercised with a wide opening angle and
tiny leaves, where the list is not empty.  Floating point: fp64 on both sides, different summation order -> 1e-11."""
import os

import numpy as np
import pytest
from conftest import DEMO_BOX, DEMO_NSIDE

import oracle
import p2p_b200
import refrun

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not os.path.isfile(refrun.REF_BIN), reason="oracle/_ref/ref_lists not built (needs /root/reference at build time)")


def _device_mid(pos, maxleaf, theta, literal_d6, periodic=True):
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(pos))
    ctx = p2p_b200.P2PContext(0)
    try:
        ctx.set_physics(1.0, eps, rs)                        # the harness runs with MASSPART = 1
        ctx.set_box([0.0, 0.0, 0.0], DEMO_BOX)
        ctx.midfield_enable(True, literal_d6)
        bdl, bdr = np.zeros(3), np.full(3, DEMO_BOX)
        ctx.tree_build(pos, maxleaf, bdl, bdr, 0)
        ctx.clear_tasks()
        ctx.tree_walk(theta, rcut, DEMO_BOX if periodic else 0.0, 0.5 * (bdr + bdl), bdr - bdl)
        nm2l = ctx.midfield_compute()
        acc = ctx.download_acc_original()                    # no P2P launched: accelerations are zero + mid-field
        ctx.build_csr()                                      # only for the task count
        return nm2l, acc, ctx.midfield_download(), ctx.counts()[0]
    finally:
        ctx.close()


@needs_ref
def test_reference_parameters_have_an_empty_m2l_list(demo_pos):
    r = refrun.run(demo_pos, DEMO_BOX, 16, DEMO_NSIDE, 0.4, True, 1)[0]
    assert int(r["idxM2L_total"][0]) == 0 and not r["acc_mid"].any()
    nm2l, acc, _, _ = _device_mid(demo_pos, 16, 0.4, True)
    assert nm2l == 0 and not acc.any()


@needs_ref
@pytest.mark.parametrize("maxleaf,theta", [(2, 1.0), (4, 1.2)])
def test_midfield_matches_the_reference_operators(demo_pos, maxleaf, theta):
    pos = demo_pos[::4].copy()                               # 8192 particles, still a periodic 32-cell box for r_s
    r = refrun.run(pos, DEMO_BOX, maxleaf, DEMO_NSIDE, theta, True, 1)[0]
    nm2l_ref = int(r["idxM2L_local"][0])
    assert nm2l_ref > 1000
    nm2l_local = _device_mid(pos, maxleaf, theta, literal_d6=False, periodic=False)[0]
    assert nm2l_local == nm2l_ref                            # the reference counts the local M2L list only (fmm.c:994)
    nm2l, acc, mid, ntask = _device_mid(pos, maxleaf, theta, literal_d6=True)
    assert nm2l > nm2l_ref                                   # + zero-shift self exchange + 26 images
    assert ntask == int(r["idxP2P_local"][0]) + sum(len(r[k]) // 2 for k in r if k.startswith("remote") and k.endswith("_tasks_ts"))
    nl = mid["leaf_M"].shape[0]
    for got, want in ((mid["leaf_M"], r["leaf_M"].reshape(-1, 20)[:nl]), (mid["node_M"], r["node_M"].reshape(-1, 20)),
                      (mid["leaf_L"], r["leaf_L"].reshape(-1, 20)[:nl])):
        scale = np.abs(want).max(axis=0) + 1e-300            # per coefficient order
        assert (np.abs(got - want) / scale).max() < 1e-11
    want = np.zeros_like(acc)
    want[r["part_orig_index"]] = r["acc_mid"].reshape(-1, 3)
    assert np.abs(want).max() > 0
    err = np.linalg.norm(acc - want, axis=1).max() / np.linalg.norm(want, axis=1).mean()
    assert err < 1e-10, err


@pytest.mark.parametrize("literal_d6", [False, True])
def test_midfield_matches_the_oracle_restatement(demo_pos, literal_d6):
    """the same comparison against oracle/p2p_oracle.c:oracle_midfield (itself pinned to the reference's operators in
    tests/test_oracle_vs_ref.py), which needs no prebuilt reference binary; literal_d6 False is the intended semantics"""
    pos = demo_pos[::4].copy()
    maxleaf, theta = 3, 1.1
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(pos))
    T = oracle.Tree(pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0)
    m = oracle.midfield(T, theta, rcut, rs, 1.0, DEMO_BOX, literal_d6=literal_d6)
    nm2l, acc, mid, _ = _device_mid(pos, maxleaf, theta, literal_d6)
    assert nm2l == m["nm2l_total"] > 1000
    for k in ("leaf_M", "node_M", "leaf_L"):
        scale = np.abs(m[k]).max(axis=0) + 1e-300
        assert (np.abs(mid[k] - m[k]) / scale).max() < 1e-11, k
    want = np.zeros_like(acc)
    want[T.perm] = m["acc"]
    assert np.linalg.norm(acc - want, axis=1).max() < 1e-10 * np.linalg.norm(want, axis=1).mean()
