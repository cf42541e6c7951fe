"""Device-resident tree build and dual-tree walk against the oracle: the tree must be the
reference's bit for bit (permutation, leaves, kd cells, split values, son ids) and the packed list must be the
oracle's list (local walk + the 26 periodic images, image sources mapped to their local leaves)."""
import numpy as np
import pytest
from conftest import DEMO_BOX, DEMO_MASS, DEMO_NSIDE, THETA

import flow
import oracle
import p2p_b200
from p2p_b200 import host, synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ctx():
    c = p2p_b200.P2PContext(0)
    yield c
    c.close()


def _assert_same_tree(D, T):
    nl, nn = T.nleaf, T.nnode
    assert (D["nleaf"], D["nnode"]) == (nl, nn)
    assert np.array_equal(D["perm"], T.perm)
    assert np.array_equal(D["pos"], T.pos)
    for k in ("leaf_npart", "leaf_ipart", "leaf_center", "leaf_width"):
        assert np.array_equal(D[k], getattr(T, k)[:nl]), k
    for k in ("node_npart", "node_son", "node_split", "node_center", "node_width"):
        assert np.array_equal(D[k], getattr(T, k)[:nn]), k


def _sorted_csr(tt, ts, nleaf):
    order = np.lexsort((ts, tt))
    return np.searchsorted(tt[order], np.arange(nleaf + 1)), ts[order]


def _image_leaf_to_local(T, img, disp):
    """image node index -> local leaf id, by the exact displaced kd cell (prepare_sendtree2 ships centre + disp)."""
    nl = T.nleaf
    key = {}
    c = T.leaf_center[:nl] + disp
    w = T.leaf_width[:nl]
    for l in range(nl):
        key[c[l].tobytes() + w[l].tobytes()] = l
    out = np.full(len(img["npart"]), -1, np.int64)
    for j in range(len(out)):
        if img["son"][j, 0] >= 0 and img["npart"][j] <= T.maxleaf:
            out[j] = key.get(img["center"][j].tobytes() + img["width"][j].tobytes(), -1)
    return out


def _oracle_periodic_list(T, box, theta, rcut):
    """local list + image lists with image sources renamed to local leaf ids"""
    tt, ts = T.walk_p2p(theta, rcut)
    TT, TS = [tt], [ts]
    tc, tw = T.node_center[0], T.node_width[0]
    for sh in flow.SHIFTS[1:]:
        disp = np.array(sh, np.float64) * box
        img = T.prune(tc, tw, disp, theta, rcut)
        a, b = T.walk_p2p_ext(img, theta, rcut)
        if len(a) == 0:
            continue
        m = _image_leaf_to_local(T, img, disp)
        assert (m[b] >= 0).all()
        TT.append(a)
        TS.append(m[b].astype(np.int32))
    return np.concatenate(TT), np.concatenate(TS)


@pytest.mark.parametrize("maxleaf", [8, 16, 32])
def test_device_walk_on_host_tree(ctx, demo_pos, golden, maxleaf):
    """walk kernels alone: tree from the product's host builder, lists from the device"""
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    T = host.LocalTree(demo_pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0)
    O = oracle.Tree(demo_pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0)
    ctx.set_physics(DEMO_MASS, eps, rs)
    ctx.set_box([0.0, 0.0, 0.0], DEMO_BOX)
    ctx.upload_particles(T.pos)
    ctx.upload_leaves(T.leaf_npart, T.leaf_ipart)
    ctx.tree_upload(T)
    # local list
    ctx.clear_tasks()
    ctx.tree_walk(THETA, rcut)
    ctx.build_csr()
    g = next(c for c in golden["cases"] if c["maxleaf"] == maxleaf and c["nproc"] == 1)["ranks"][0]
    assert ctx.counts() == (g["local_tasks"], g["local_pairs"])
    row, col = ctx.download_csr()
    tt, ts = O.walk_p2p(THETA, rcut)
    r0, c0 = _sorted_csr(tt, ts, O.nleaf)
    assert np.array_equal(row, r0) and np.array_equal(col, c0)
    # local + periodic images
    ctx.clear_tasks()
    ctx.tree_walk(THETA, rcut, DEMO_BOX, T.node_center[0], T.node_width[0])
    ctx.build_csr()
    row, col = ctx.download_csr()
    tt, ts = _oracle_periodic_list(O, DEMO_BOX, THETA, rcut)
    r0, c0 = _sorted_csr(tt, ts, O.nleaf)
    assert np.array_equal(row, r0) and np.array_equal(col, c0)
    assert ctx.csr_duplicates() == 0
    if maxleaf == 16:
        assert ctx.counts() == (381377 + 205240, 83354950 + 36821870)


def _inputs(kind, n, seed, box):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.random((n, 3)) * box
    if kind == "f32":
        return (rng.random((n, 3)) * box).astype(np.float32).astype(np.float64)
    if kind == "grid":                                   # many exactly equal coordinates: ties at every level
        return np.floor(rng.random((n, 3)) * 16.0) / 16.0 * box + rng.random((n, 3)) * 1e-9 * box * (rng.random((n, 1)) < 0.5)
    if kind == "clumpy":
        c = rng.random((12, 3)) * box
        return np.mod(c[rng.integers(0, 12, n)] + rng.normal(0.0, 0.01 * box, (n, 3)), box)
    raise ValueError(kind)


@pytest.mark.parametrize("kind,n,maxleaf,direct", [
    ("uniform", 5000, 8, 0), ("uniform", 40000, 16, 1), ("f32", 30000, 32, 2), ("clumpy", 30000, 16, 0),
    ("grid", 20000, 32, 1), ("uniform", 3, 1, 0), ("uniform", 2, 1, 0), ("uniform", 2, 8, 0), ("uniform", 1, 8, 0), ("uniform", 37, 1, 2),
    ("f32", 300000, 32, 0)])
@pytest.mark.parametrize("plain_max", [-1, -2, -3, -4, -5, -6, -7, 0, 1 << 30])
def test_device_build_is_the_reference_tree(ctx, kind, n, maxleaf, direct, plain_max):
    """plain_max -1: defaults (speculative chunks / block per long node, warp per node, in-order fold for short runs, node-centric
    deep levels); -2: no block variant; -3: no speculative chunks; -4: speculative chunks for every node above 2048 particles;
    -5: particle-wide kernels for all levels; -6: no block-centric middle levels; -7: block-centric levels from the root on;
    0: every split mean through the parallel transducer sum; 1<<30: every one through the in-order fold"""
    box = 1000.0
    pos = _inputs(kind, n, 1234 + n, box)
    bdl, bdr = [0.0, 0.0, 0.0], [box, box, box]
    try:
        O = oracle.Tree(pos, maxleaf, bdl, bdr, direct)
    except RuntimeError:
        O = None
    ctx.set_physics(1.0, 0.01, 10.0)
    ctx.set_box([0.0, 0.0, 0.0], box)
    ctx.tree_set_option(plain_max)
    try:
        if O is None:
            with pytest.raises(p2p_b200.P2PError):
                ctx.tree_build(pos, maxleaf, bdl, bdr, direct)
            return
        ctx.tree_build(pos, maxleaf, bdl, bdr, direct)
    finally:
        ctx.tree_set_option(-1)
    _assert_same_tree(ctx.tree_download(), O)


def test_device_build_demo_and_domain_box(ctx, demo_pos):
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    ctx.set_physics(DEMO_MASS, eps, rs)
    ctx.set_box([0.0, 0.0, 0.0], DEMO_BOX)
    for maxleaf in (8, 16, 32):
        O = oracle.Tree(demo_pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0)
        ctx.tree_build(demo_pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0)
        _assert_same_tree(ctx.tree_download(), O)
    # a rank's sub-box with a different first split direction (multi-rank local trees)
    sel = demo_pos[demo_pos[:, 0] < 0.5 * DEMO_BOX]
    bdl, bdr = [0.0, 0.0, 0.0], [0.5 * DEMO_BOX, DEMO_BOX, DEMO_BOX]
    O = oracle.Tree(sel, 16, bdl, bdr, 1)
    ctx.tree_build(sel, 16, bdl, bdr, 1)
    _assert_same_tree(ctx.tree_download(), O)


@pytest.mark.parametrize("maxleaf", [16, 32])
def test_device_full_step_demo(ctx, demo_pos, maxleaf):
    """positions in, accelerations out, everything in between on the device"""
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    ctx.set_physics(DEMO_MASS, eps, rs)
    ctx.set_box([0.0, 0.0, 0.0], DEMO_BOX)
    bdl, bdr = np.zeros(3), np.full(3, DEMO_BOX)
    ctx.tree_build(demo_pos, maxleaf, bdl, bdr, 0)
    ctx.clear_tasks()
    ctx.tree_walk(THETA, rcut, DEMO_BOX, 0.5 * (bdr + bdl), bdr - bdl)
    ctx.build_csr()
    assert ctx.csr_duplicates() == 0
    ctx.compute()
    acc = ctx.download_acc_original()
    ref, ntask, npairs = flow.reference_forces(demo_pos, DEMO_BOX, maxleaf, DEMO_NSIDE, THETA, DEMO_MASS, 1, True)
    assert ctx.counts() == (ntask, npairs)
    absr, _, _ = flow.reference_forces(demo_pos, DEMO_BOX, maxleaf, DEMO_NSIDE, THETA, DEMO_MASS, 1, True, absterms=True)
    d = np.linalg.norm(acc - ref, axis=1)
    na = np.linalg.norm(ref, axis=1)
    e1 = (d / np.maximum(na, na.mean())).max()
    e2 = (d / np.maximum(np.linalg.norm(absr, axis=1), 1e-300)).max()
    assert e1 < TOL and e2 < TOL, (e1, e2)


def test_device_path_equals_host_path_at_scale(ctx):
    """128^3: the device-built tree equals the host library's (itself pinned to the oracle), the device walk
    yields the host lists' task and pair counts, and the forces agree with the host-list path"""
    from p2p_b200 import step
    nside, maxleaf = 128, 32
    pos, box = synth.zeldovich_like(nside)
    mass = 1.0
    rs, rcut, eps = host.derived_params(box, nside, pos.shape[0])
    L = step.build_lists(pos, box, maxleaf, nside, THETA, periodic=True)
    st = step.ShortRangeStep(0)
    acc_host = st.run(L, mass, truncated=True)                 # original order
    nt_host, np_host = st.ctx.counts()
    st.ctx.close()
    ctx.set_physics(mass, eps, rs)
    ctx.set_box([0.0, 0.0, 0.0], box)
    bdl, bdr = np.zeros(3), np.full(3, box)
    ctx.tree_build(pos, maxleaf, bdl, bdr, 0)
    _assert_same_tree(ctx.tree_download(), L.tree)
    ctx.clear_tasks()
    ctx.tree_walk(THETA, rcut, box, 0.5 * (bdr + bdl), bdr - bdl)
    ctx.build_csr()
    assert ctx.counts() == (nt_host, np_host)
    assert ctx.csr_duplicates() == 0
    ctx.compute()
    acc = ctx.download_acc_original()
    d = np.linalg.norm(acc - acc_host, axis=1)
    na = np.linalg.norm(acc_host, axis=1)
    assert (d / np.maximum(na, na.mean())).max() < 4e-6      # same arithmetic; summation order and near / far class of image sources differ
    info = ctx.tree_info()
    print("128^3 device tree: build %.2f ms, walk %.2f ms, %d levels, %d walk items" %
          (info["ms_build"], info["ms_walk"], info["nlevel"], info["walk_items"]))


def test_walk_argument_errors(ctx, demo_pos):
    c = p2p_b200.P2PContext(0)
    try:
        with pytest.raises(p2p_b200.P2PError):
            c.tree_walk(THETA, 100.0)                              # no tree
        with pytest.raises(p2p_b200.P2PError):
            c.tree_build(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, 0)   # no box set
        c.set_box([0.0, 0.0, 0.0], DEMO_BOX)
        with pytest.raises(p2p_b200.P2PError):
            c.tree_build(demo_pos, 64, [0, 0, 0], [DEMO_BOX] * 3, 0)   # maxleaf above P2P_MAX_LEAF
        with pytest.raises(p2p_b200.P2PError):
            c.tree_build(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, 3)   # bad direction
    finally:
        c.close()


def test_degenerate_inputs_fail_loudly(ctx, demo_pos):
    """more coincident particles than a leaf holds: the reference recurses until it overruns its arrays; here an error"""
    box = 1000.0
    pos = np.full((100, 3), 123.456)
    ctx.set_physics(1.0, 0.01, 10.0)
    ctx.set_box([0.0, 0.0, 0.0], box)
    with pytest.raises(p2p_b200.P2PError):
        ctx.tree_build(pos, 8, [0, 0, 0], [box] * 3, 0)
    # a periodic box too small for minimal-image sources is refused by the one-call step
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, 8, len(demo_pos))          # r_s of an 8-cell box: r_cut > box / 2
    ctx.set_physics(DEMO_MASS, eps, rs)
    ctx.set_box([0.0, 0.0, 0.0], DEMO_BOX)
    with pytest.raises(p2p_b200.P2PError):
        ctx.step_device(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, THETA, rcut, DEMO_BOX)
    # and the context is still usable afterwards
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    ctx.set_physics(DEMO_MASS, eps, rs)
    acc = ctx.step_device(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, THETA, rcut, DEMO_BOX)
    assert np.isfinite(acc).all() and ctx.counts() == (381377 + 205240, 83354950 + 36821870)


def test_non_periodic_step_and_repeatability(ctx, demo_pos):
    """period 0: local list only; two runs give bit-identical accelerations (sorted rows, no atomics on floats)"""
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    ctx.set_physics(DEMO_MASS, eps, rs)
    ctx.set_box([0.0, 0.0, 0.0], DEMO_BOX)
    a1 = ctx.step_device(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, THETA, rcut, 0.0).copy()
    assert ctx.counts() == (381377, 83354950)
    a2 = ctx.step_device(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, THETA, rcut, 0.0)
    assert np.array_equal(a1, a2)
    O = oracle.Tree(demo_pos, 16, [0, 0, 0], [DEMO_BOX] * 3, 0)
    tt, ts = O.walk_p2p(THETA, rcut)
    ref, _ = oracle.p2p(O.pos, O.leaf_npart, O.leaf_ipart, O.pos, O.leaf_npart, O.leaf_ipart, tt, ts, DEMO_MASS, eps, rs)
    want = np.zeros_like(ref)
    want[O.perm] = ref
    na = np.linalg.norm(want, axis=1)
    assert (np.linalg.norm(a1 - want, axis=1) / np.maximum(na, na.mean())).max() < TOL


def test_target_chunks_reproduce_the_unchunked_step(demo_pos):
    """p2p_forces_local in target chunks (bounded list memory: 1024^3 on one GPU lists 6.7e9 tasks) == one chunk: the same
    task and pair totals, the same M2L list, and -- every row lives in exactly one chunk -- bit-identical accelerations."""
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    bdl, bdr = np.zeros(3), np.full(3, DEMO_BOX)
    out = []
    for chunk_tasks in (1 << 29, 40000):
        ctx = p2p_b200.P2PContext(0)
        ctx.set_physics(DEMO_MASS, eps, rs)
        ctx.set_box([0.0, 0.0, 0.0], DEMO_BOX)
        ctx.set_chunk_tasks(chunk_tasks)
        ctx.midfield_enable(True, False)
        ctx.tree_build(demo_pos[::2].copy(), 4, bdl, bdr, 0)
        ctx.forces_local(1.0, rcut, DEMO_BOX, 0.5 * (bdl + bdr), bdr - bdl)
        nm2l = ctx.midfield_compute()
        st = ctx.step_timings()
        out.append((ctx.accumulated_counts(), nm2l, ctx.download_acc(), st["chunks"], ctx.download_acc_original()))
        ctx.close()
    assert out[0][3] == 1 and out[1][3] > 8
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1] > 0
    assert np.array_equal(out[0][2], out[1][2])                      # P2P part (FP32, tree order): bit for bit
    # with the mid-field (fp64 atomics: summation order differs)
    assert np.abs(out[0][4] - out[1][4]).max() < 1e-10 * np.abs(out[0][4]).max()


def test_pipelined_chunks_equal_the_unpipelined_step():
    """The chunk pipeline of p2p_forces_local (walk + packing of chunk k + 1 on a second stream and in the other list set beside
    the force kernel of chunk k, retiring warps) against the same step in one piece: identical counts, identical M2L list,
    bit-identical P2P accelerations -- every row lives in exactly one chunk and is summed in the same order."""
    from p2p_b200 import synth
    nside = 64
    pos, box = synth.clustered(nside)
    rs, rcut, eps = oracle.derived_params(box, nside, len(pos))
    bdl, bdr = np.zeros(3), np.full(3, box)
    out = []
    for pipe in (0, 4, 7):
        ctx = p2p_b200.P2PContext(0)
        ctx.set_physics(DEMO_MASS, eps, rs)
        ctx.set_box([0.0, 0.0, 0.0], box)
        ctx.set_chunk_pipeline(pipe)
        ctx.midfield_enable(True, False)
        for rep in range(2):                                   # the second step reuses buffers, events and the second stream
            ctx.tree_build(pos, 8, bdl, bdr, 0)
            ctx.forces_local(THETA, rcut, box, 0.5 * (bdl + bdr), bdr - bdl)
            nm2l = ctx.midfield_compute()
            st = ctx.step_timings()
            res = (ctx.accumulated_counts(), nm2l, ctx.download_acc(), st["chunks"])
        out.append(res)
        ctx.close()
    assert [o[3] for o in out] == [1, 4, 7]
    for o in out[1:]:
        assert o[0] == out[0][0] and o[1] == out[0][1]
        assert np.array_equal(o[2], out[0][2])

