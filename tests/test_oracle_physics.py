"""Known-answer and invariant tests of the oracle's pair arithmetic (there is no CPU p2p_kernel in
the reference tree to run, SURVEY fact 3, so the arithmetic is pinned by closed forms)."""
import math

import numpy as np
from conftest import DEMO_BOX, DEMO_MASS, DEMO_NSIDE, THETA

import oracle


def _one_pair(r, eps, rs, m=2.5):
    tpos = np.array([[0.0, 0.0, 0.0]])
    spos = np.array([[r, 0.0, 0.0]])
    one, zero = np.array([1], np.int32), np.array([0], np.int32)
    acc, n = oracle.p2p(tpos, one, zero, spos, one, zero, zero, zero, m, eps, rs)
    assert n == 1
    return acc[0, 0]


def test_two_body_closed_form():
    m = 2.5
    for r in (0.5, 3.0, 40.0):
        assert math.isclose(_one_pair(r, 1.0, 0.0, m), m * r / max(r, 1.0) ** 3, rel_tol=1e-15)   # I cu:346-354
    rs = 10.0
    for r in (0.5, 3.0, 40.0, 80.0):
        u = 0.5 * r / rs
        g = math.erfc(u) + 2 / math.sqrt(math.pi) * u * math.exp(-u * u)                              # R cu:443-446
        assert math.isclose(_one_pair(r, 1.0, rs, m), m * r / max(r, 1.0) ** 3 * g, rel_tol=1e-14)


def test_truncation_limits():
    rs = 3906.25
    assert abs(_one_pair(1e-3 * rs, 0.0, rs, 1.0) * (1e-3 * rs) ** 2 - 1.0) < 1e-8          # g -> 1 as r -> 0
    rcut = 4.5 * rs
    g_cut = _one_pair(rcut, 0.0, rs, 1.0) * rcut ** 2
    assert abs(g_cut - 0.0176) < 2e-4                                                       # ~0.0176 at r_cut


def test_self_pair_contributes_zero_and_list_is_antisymmetric(demo_pos):
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    T = oracle.Tree(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, 0)
    tt, ts = T.walk_p2p(THETA, rcut)
    fwd = set(zip(tt.tolist(), ts.tolist()))
    assert len(fwd) == len(tt)                                       # duplicate-free
    assert all((s, t) in fwd for (t, s) in fwd)                      # equal to its own transpose
    acc, _ = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, DEMO_MASS, eps, rs)
    # Newton's third law over a closed symmetric list: sum_i m a_i = 0
    tot = np.abs(acc.sum(axis=0)).max() / np.abs(acc).sum(axis=0).max()
    assert tot < 1e-12


def test_direct_summation_small_n():
    rng = np.random.default_rng(5)
    n = 300
    pos = rng.uniform(0, 100.0, (n, 3))
    T = oracle.Tree(pos, 8, [0, 0, 0], [100.0] * 3, 0)
    # all-pairs list over leaves
    tt, ts = np.meshgrid(np.arange(T.nleaf, dtype=np.int32), np.arange(T.nleaf, dtype=np.int32), indexing="ij")
    acc, npairs = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt.ravel(), ts.ravel(),
                             1.0, 0.5, 20.0)
    assert npairs == n * n
    d = T.pos[None, :, :] - T.pos[:, None, :]
    r = np.sqrt((d ** 2).sum(-1))
    u = 0.5 * r / 20.0
    from scipy.special import erfc
    g = erfc(u) + 2 / np.sqrt(np.pi) * u * np.exp(-u * u)
    ref = (d * (g / np.maximum(r, 0.5) ** 3)[:, :, None]).sum(1)
    assert np.abs(acc - ref).max() / np.abs(ref).max() < 1e-12
