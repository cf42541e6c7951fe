"""Force parity at the BENCH workload and on the clustered box ("parity holes"): the device-resident step
(tree build + dual-tree walk incl. the 26 periodic images + packing + forces, p2p_step_device) against the fp64 oracle on
COMPLETE rows of the list the device produced -- every source leaf a sampled target leaf interacts with, image sources
included -- at sizes where the oracle cannot run the whole list.

Tolerance (north_star: 1e-5 relative), as everywhere:
  e1 = max_i |da_i| / max(|a_i|, mean|a|) <= 1e-5,   e2 = max_i |da_i| / sum_pairs |term|_i <= 1e-5."""
import numpy as np
import pytest
from conftest import THETA

import oracle
import p2p_b200
from p2p_b200 import host, synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rows_against_oracle(ctx, pos_tree, leaf_npart, leaf_ipart, rows, box, mass, eps, rs):
    """oracle forces (and sums of |terms|) of the particles of the target leaves `rows`, from the device's own CSR rows.
    Image / wrapped sources are displaced to the image nearest the target leaf, which is what the fixed-point coordinates
    of the device do pair by pair.  Returns (particle indices [tree order], a_ref, abs_ref, number of pairs)."""
    row, col = ctx.download_csr()
    tt = np.concatenate([np.full(row[r + 1] - row[r], r, np.int32) for r in rows])
    src = np.concatenate([col[row[r]:row[r + 1]] for r in rows]).astype(np.int64)
    cnt = leaf_npart[src].astype(np.int64)
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    pidx = np.repeat(leaf_ipart[src].astype(np.int64) - start, cnt) + np.arange(cnt.sum())
    spos = pos_tree[pidx].copy()
    ref = np.repeat(pos_tree[leaf_ipart[tt]], cnt, axis=0)                    # first particle of the task's target leaf
    if box > 0:
        spos -= box * np.round((spos - ref) / box)
    ts = np.arange(len(src), dtype=np.int32)
    a, npairs = oracle.p2p(pos_tree, leaf_npart, leaf_ipart, spos, cnt.astype(np.int32), start.astype(np.int32), tt, ts, mass, eps, rs)
    b, _ = oracle.p2p(pos_tree, leaf_npart, leaf_ipart, spos, cnt.astype(np.int32), start.astype(np.int32), tt, ts, mass, eps, rs,
                      absterms=True)
    sel = np.concatenate([np.arange(leaf_ipart[r], leaf_ipart[r] + leaf_npart[r]) for r in rows])
    return sel, a[sel], b[sel], npairs


def _errors(got, ref, absr):
    d = np.linalg.norm(got - ref, axis=1)
    na = np.linalg.norm(ref, axis=1)
    return float((d / np.maximum(na, na.mean())).max()), float((d / np.maximum(np.linalg.norm(absr, axis=1), 1e-300)).max())


def test_bench_workload_rows_match_the_oracle():
    """256^3, MAXLEAF 32, periodic, truncated kernel -- the workload bench.py times: 600 complete rows."""
    ns = 256
    pos, box = synth.zeldovich_like(ns)
    rs, rcut, eps = host.derived_params(box, ns, len(pos))
    ctx = p2p_b200.P2PContext(0)
    try:
        ctx.set_physics(synth.DEMO_MASS, eps, rs)
        ctx.set_box([0.0, 0.0, 0.0], box)
        ctx.step_device(pos, 32, [0.0] * 3, [box] * 3, THETA, rcut, box)
        assert ctx.csr_duplicates() == 0
        D = ctx.tree_download()
        got = ctx.download_acc()                                             # tree order
        rng = np.random.default_rng(11)
        rows = np.sort(rng.choice(D["nleaf"], 600, replace=False))
        sel, a, b, npairs = rows_against_oracle(ctx, D["pos"], D["leaf_npart"], D["leaf_ipart"], rows, box, synth.DEMO_MASS, eps, rs)
        assert npairs > 5e7
        e1, e2 = _errors(got[sel], a, b)
        assert e1 < TOL and e2 < TOL, (e1, e2)
    finally:
        ctx.close()


@pytest.mark.parametrize("ns,nrows", [(64, 0), (128, 800)])
def test_clustered_box_with_midfield_matches_the_oracle(ns, nrows):
    """Zel'dovich + NFW-like clumps: softened pairs, rows of a few targets, long rows, and a non-empty M2L list.
    ns 64: every particle against the oracle's whole step (P2P + mid-field); ns 128: sampled complete rows."""
    pos, box = synth.clustered(ns)
    rs, rcut, eps = host.derived_params(box, ns, len(pos))
    mass = synth.DEMO_MASS
    T = oracle.Tree(pos, 32, [0, 0, 0], [box] * 3, 0)
    mid = oracle.midfield(T, THETA, rcut, rs, mass, box)
    ctx = p2p_b200.P2PContext(0)
    try:
        ctx.set_physics(mass, eps, rs)
        ctx.set_box([0.0, 0.0, 0.0], box)
        ctx.midfield_enable(True, False)
        bdl, bdr = np.zeros(3), np.full(3, box)
        ctx.tree_build(pos, 32, bdl, bdr, 0)
        ctx.clear_tasks()
        ctx.tree_walk(THETA, rcut, box, 0.5 * (bdl + bdr), bdr - bdl)
        ctx.build_csr()
        ctx.compute()
        nm2l = ctx.midfield_compute()
        assert nm2l == mid["nm2l_total"] > 0
        assert ctx.csr_duplicates() == 0
        D = ctx.tree_download()
        assert np.array_equal(D["pos"], T.pos) and np.array_equal(D["leaf_ipart"], T.leaf_ipart[:T.nleaf])
        got = ctx.download_acc_original()[T.perm]                            # P2P + mid-field, tree order
        rows = np.arange(T.nleaf) if nrows == 0 else np.sort(np.random.default_rng(12).choice(T.nleaf, nrows, replace=False))
        rows = rows[T.leaf_npart[rows] > 0]
        sel, a, b, npairs = rows_against_oracle(ctx, T.pos, T.leaf_npart[:T.nleaf], T.leaf_ipart[:T.nleaf], rows, box, mass, eps, rs)
        far, _ = ctx.download_csr_class()
        assert 0.2 < far.mean() < 0.95
        e1, e2 = _errors(got[sel], a + mid["acc"][sel], b)
        assert e1 < TOL and e2 < TOL, (ns, e1, e2)
    finally:
        ctx.close()
