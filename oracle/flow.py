"""Whole-step restatement of the reference's short-range P2P flow for P ranks, built from the C
pieces in p2p_oracle.c.  TEST INFRASTRUCTURE ONLY (see oracle.py).

Mirrors I/src/photoNs.c:83-123 minus PM and M2L:
  domain_initialize -> domain_decomposition (first call: equal-volume splits, route particles)
  -> fmm_construct/fmm_prepare (local kd-tree) -> fmm_task (local walk) -> fmm_ext (for every peer
  index n and every one of the 27 displacements: prune the sender's tree against my domain box,
  walk my tree against it).
"""
import numpy as np

import oracle

SHIFTS = [(0, 0, 0)] + [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) if (i, j, k) != (0, 0, 0)]
"""Order of the displacements as I/src/fmm.c:1056-1106 issues them: zero first, then mi, mj, mk loops."""


def decompose(pos, box, nproc):
    """Initial routing exactly as the harness + I/src/domains.c:298-377 do it: rank r starts with the
    slab [N r/P, N (r+1)/P) of the input order, partitions it in place by the equal-volume domain
    tree and ships blocks; the receiver concatenates blocks in source-rank order.
    Returns (list of (pos_r, orig_index_r), split, direct_of_node)."""
    n = pos.shape[0]
    split, center, width, direct = oracle.domain_setup(nproc, box)
    if nproc == 1:
        return [(np.array(pos, np.float64, copy=True), np.arange(n, dtype=np.int64))], split, direct
    blocks = [[None] * nproc for _ in range(nproc)]  # blocks[dst][src]
    for r in range(nproc):
        lo, hi = n * r // nproc, n * (r + 1) // nproc
        p = np.array(pos[lo:hi], np.float64, order="C", copy=True)
        idx = np.arange(lo, hi, dtype=np.int64)
        send = oracle.domain_partition(nproc, split, p, idx)
        off = np.concatenate([[0], np.cumsum(send)])
        for d in range(nproc):
            blocks[d][r] = (p[off[d]:off[d + 1]], idx[off[d]:off[d + 1]])
    out = []
    for d in range(nproc):
        out.append((np.concatenate([b[0] for b in blocks[d]]), np.concatenate([b[1] for b in blocks[d]])))
    return out, split, direct


def short_range_lists(pos, box, maxleaf, nside, theta, nproc=1, do_ext=True, literal_d6=True):
    """Returns one dict per rank: tree (oracle.Tree), orig_index, local (tt, ts), and `remote`: list of
    dicts {shift, src_rank, tree (pruned image), tt, ts} in the reference's call order.
    literal_d6=True reproduces the reference's zero-shift self exchange (defect D6: the n=0 call
    regenerates the local list); False skips it (the intended semantics)."""
    rs, rcut, eps = oracle.derived_params(box, nside, pos.shape[0])
    parts, split, direct = decompose(pos, box, nproc)
    _, tcenter, twidth, _ = oracle.domain_setup(nproc, box)
    ranks = []
    for r in range(nproc):
        dom = oracle.domain_of_rank(nproc, r)
        c, w = tcenter[dom], twidth[dom]
        # I/src/fmm.c:194-197: the local box is rebuilt from the toptree centre and width
        T = oracle.Tree(parts[r][0], maxleaf, c - 0.5 * w, c + 0.5 * w, int(direct[dom]))
        ranks.append(dict(tree=T, orig_index=parts[r][1][T.perm], domain=dom))
    # I/src/toptree.c:18-45: every domain's toptree box is overwritten by its local root box
    for r in range(nproc):
        T = ranks[r]["tree"]
        tcenter[ranks[r]["domain"]] = T.node_center[0]
        twidth[ranks[r]["domain"]] = T.node_width[0]
    for r in range(nproc):
        T = ranks[r]["tree"]
        ranks[r]["local"] = T.walk_p2p(theta, rcut)
        ranks[r]["remote"] = []
        ranks[r]["params"] = dict(rs=rs, rcut=rcut, eps=eps)
        if not do_ext:
            continue
        for si, sh in enumerate(SHIFTS):
            for n in range(nproc):
                if si == 0 and n == 0 and not literal_d6:
                    continue
                src = (r - n + nproc) % nproc  # I/src/remotes.c:747 (rrank); sender pruned against MY box
                disp = np.array(sh, np.float64) * box
                img = ranks[src]["tree"].prune(tcenter[ranks[r]["domain"]], twidth[ranks[r]["domain"]], disp, theta, rcut)
                tt, ts = T.walk_p2p_ext(img, theta, rcut)
                ranks[r]["remote"].append(dict(shift=disp, src_rank=src, image=img, tt=tt, ts=ts))
    return ranks


def reference_forces(pos, box, maxleaf, nside, theta, mass, nproc=1, truncated=True, nthreads=0, absterms=False):
    """Intended short-range accelerations (D6 fixed: no zero-shift self exchange) in the ORIGINAL
    particle order, plus total task and pair counts."""
    ranks = short_range_lists(pos, box, maxleaf, nside, theta, nproc, True, literal_d6=False)
    acc = np.zeros((pos.shape[0], 3))
    ntask = npairs = 0
    for rk in ranks:
        T = rk["tree"]
        prm = rk["params"]
        rs = prm["rs"] if truncated else 0.0
        a = np.zeros((T.npart, 3))
        tt, ts = rk["local"]
        _, npr = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, mass,
                            prm["eps"], rs, acc=a, nthreads=nthreads, absterms=absterms)
        ntask += len(tt); npairs += npr
        for rem in rk["remote"]:
            img = rem["image"]
            if len(rem["tt"]) == 0:
                continue
            # remote leaf bodies: [son[0], son[0]+npart) of the received body array (R/src/remotes.c:63-64)
            _, npr = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, img["body"], img["npart"], img["son"][:, 0].copy(),
                                rem["tt"], rem["ts"], mass, prm["eps"], rs, acc=a, nthreads=nthreads, absterms=absterms)
            ntask += len(rem["tt"]); npairs += npr
        acc[rk["orig_index"]] = a
    return acc, ntask, npairs
