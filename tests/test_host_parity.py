"""The product's own host code (libp2p_host.so: subtree-parallel build, frontier-parallel walks,
halo pruning, domain routing) against the oracle: bit-exact trees, identical task SEQUENCES."""
import numpy as np
import pytest
from conftest import DEMO_BOX, DEMO_NSIDE, THETA

import flow
import oracle
from p2p_b200 import host, step, synth

TREE_ATTRS = ("leaf_npart", "leaf_ipart", "leaf_center", "leaf_width", "node_npart", "node_son", "node_split",
              "node_center", "node_width")


def _same_tree(T, O):
    assert (T.nleaf, T.nnode, T.first_leaf, T.first_node) == (O.nleaf, O.nnode, O.first_leaf, O.first_node)
    for a in TREE_ATTRS:
        assert np.array_equal(getattr(T, a), getattr(O, a)[: len(getattr(T, a))]), a
    assert np.array_equal(T.pos, O.pos) and np.array_equal(T.perm, O.perm)


def _inputs():
    rng = np.random.default_rng(11)
    box = 1000.0
    c = rng.uniform(0, box, (9, 3))
    clumpy = (c[rng.integers(0, 9, 40000)] + rng.normal(0, 20.0, (40000, 3))) % box
    clumpy[:500] = clumpy[500:1000]
    return {
        "clumpy": (clumpy.astype(np.float32).astype(np.float64), box, 16),
        "zeldovich": synth.zeldovich_like(24, seed=3, box=synth.BOX * 24 / 32)[:2] + (24,),
        "tiny": (rng.uniform(0, box, (40, 3)), box, 4),
    }


@pytest.mark.parametrize("maxleaf", [8, 16, 32])
@pytest.mark.parametrize("nthreads", [1, 4])
def test_demo_tree_and_walk(demo_pos, maxleaf, nthreads):
    rs, rcut, eps = host.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    assert (rs, rcut, eps) == oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    O = oracle.Tree(demo_pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0)
    T = host.LocalTree(demo_pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0, nthreads)
    _same_tree(T, O)
    ott, ots = O.walk_p2p(THETA, rcut)
    tt, ts = T.walk_task_p2p(THETA, rcut, nthreads)
    assert np.array_equal(tt, ott) and np.array_equal(ts, ots)


@pytest.mark.parametrize("name", ["clumpy", "zeldovich", "tiny"])
@pytest.mark.parametrize("maxleaf", [2, 8, 32])
def test_other_inputs(name, maxleaf):
    pos, box, nside = _inputs()[name]
    rs, rcut, eps = host.derived_params(box, nside, len(pos))
    for direct in (0, 2):
        O = oracle.Tree(pos, maxleaf, [0, 0, 0], [box] * 3, direct)
        T = host.LocalTree(pos, maxleaf, [0, 0, 0], [box] * 3, direct)
        _same_tree(T, O)
        a, b = O.walk_p2p(THETA, rcut)
        c, d = T.walk_task_p2p(THETA, rcut)
        assert np.array_equal(a, c) and np.array_equal(b, d)


def test_halo_images_and_ext_walk(demo_pos):
    rs, rcut, eps = host.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    O = oracle.Tree(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, 0)
    T = host.LocalTree(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, 0)
    for sh in flow.SHIFTS:
        d = np.array(sh, float) * DEMO_BOX
        io = O.prune(O.node_center[0], O.node_width[0], d, THETA, rcut)
        ih = T.prepare_sendtree(T.node_center[0], T.node_width[0], d, THETA, rcut)
        for k, a in (("npart", ih.npart), ("son", ih.son), ("center", ih.center), ("width", ih.width), ("body", ih.body)):
            assert np.array_equal(io[k], a), (sh, k)
        a, b = O.walk_p2p_ext(io, THETA, rcut)
        c, e = T.walk_task_p2p_ext(ih, THETA, rcut)
        assert np.array_equal(a, c) and np.array_equal(b, e)


@pytest.mark.parametrize("nproc", [1, 2, 3, 4, 8])
def test_domain_setup_and_route(demo_pos, nproc):
    so, co, wo, do = oracle.domain_setup(nproc, DEMO_BOX)
    sh, ch, wh, dh = host.domain_setup(nproc, DEMO_BOX)
    assert np.array_equal(so, sh) and np.array_equal(co, ch) and np.array_equal(wo, wh) and np.array_equal(do, dh)
    assert [oracle.domain_of_rank(nproc, r) for r in range(nproc)] == [host.domain_of_rank(nproc, r) for r in range(nproc)]
    p1, p2 = demo_pos.copy(), demo_pos.copy()
    i1 = np.arange(len(p1), dtype=np.int64)
    i2 = i1.copy()
    s1 = oracle.domain_partition(nproc, so, p1, i1)
    s2 = host.domain_route(nproc, sh, p2, i2)
    assert np.array_equal(s1, s2) and np.array_equal(p1, p2) and np.array_equal(i1, i2)
    assert s2.sum() == len(p1)
    work = np.random.default_rng(nproc).integers(50000, 150000, nproc).astype(float)
    assert np.array_equal(oracle.domain_relax(nproc, DEMO_BOX, so, work), host.domain_relax(nproc, DEMO_BOX, sh, work))


def test_step_lists_equal_oracle_flow(demo_pos):
    """build_lists (local list + 26 periodic-image ghost lists, D6 skipped) == oracle flow, task for task."""
    L = step.build_lists(demo_pos, DEMO_BOX, 16, DEMO_NSIDE, THETA, periodic=True)
    rk = flow.short_range_lists(demo_pos, DEMO_BOX, 16, DEMO_NSIDE, THETA, 1, True, literal_d6=False)[0]
    assert np.array_equal(L.tt, rk["local"][0]) and np.array_equal(L.ts, rk["local"][1])
    assert len(L.gtt) == sum(len(r["tt"]) for r in rk["remote"]) == 205240
    # ghost tasks reference the same displaced bodies
    o = 0
    for rem in rk["remote"]:
        n = len(rem["tt"])
        if n == 0:
            continue
        assert np.array_equal(L.gtt[o:o + n], rem["tt"])
        k = np.random.default_rng(0).integers(0, n, 50)
        for i in k:
            g = L.gts[o + i]
            a = L.ghost_pos[L.ghost_start[g]: L.ghost_start[g] + L.ghost_count[g]]
            s = rem["ts"][i]
            b = rem["image"]["body"][rem["image"]["son"][s, 0]: rem["image"]["son"][s, 0] + rem["image"]["npart"][s]]
            assert np.array_equal(a, b)
        o += n


def test_domain_boxes_of_given_splits():
    """center_toptree for arbitrary splits: reproduces domain_setup's boxes for its own splits, and follows relaxed splits"""
    from p2p_b200 import host
    for P in (1, 2, 3, 4, 8):
        split, center, width, direct = host.domain_setup(P, 100.0)
        c2, w2, d2 = host.domain_boxes(P, 100.0, split)
        assert np.array_equal(center, c2) and np.array_equal(width, w2) and np.array_equal(direct, d2)
    split, center, width, direct = host.domain_setup(4, 100.0)
    relaxed = host.domain_relax(4, 100.0, split, [4.0, 1.0, 1.0, 1.0])
    c3, w3, _ = host.domain_boxes(4, 100.0, relaxed)
    assert relaxed[0] != split[0] and c3[1][0] == 0.5 * relaxed[0] and w3[1][0] == relaxed[0]
    # leaf boxes tile the volume
    assert abs(sum(np.prod(w3[n]) for n in range(3, 7)) - 100.0 ** 3) < 1e-6
