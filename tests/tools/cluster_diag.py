"""Diagnostic (development aid): the worst particles of the clustered 64^3 box against the fp64 oracle, for several
kernel variants, with the oracle's force of the worst particle split by u = r / 2 r_s and by near / far class.
usage: python tests/tools/cluster_diag.py [nside]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"), os.path.join(ROOT, "tests")]
import oracle  # noqa: E402
import p2p_b200  # noqa: E402
from p2p_b200 import host, synth  # noqa: E402
from test_gpu_parity_scale import rows_against_oracle  # noqa: E402
from scipy.special import erfc  # noqa: E402

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pos, box = synth.clustered(ns)
rs, rcut, eps = host.derived_params(box, ns, len(pos))
mass = synth.DEMO_MASS
T = oracle.Tree(pos, 32, [0, 0, 0], [box] * 3, 0)
ctx = p2p_b200.P2PContext(0)
ctx.set_physics(mass, eps, rs)
ctx.set_box([0.0, 0.0, 0.0], box)
bdl, bdr = np.zeros(3), np.full(3, box)
res = {}
ref = None
for name, (variant, tt, far) in {"v2 default": (2, 0, -1.0), "v2 no far class": (2, 32, 0.0), "v1 packed": (2, 16, -1.0), "v1 scalar": (1, 16, -1.0),
                                "v2 far from 2.0": (2, 32, 2.0)}.items():
    ctx.set_kernel_variant(variant); ctx.set_tuning(tt, 0, 0); ctx.set_far_threshold(far)
    ctx.tree_build(pos, 32, bdl, bdr, 0)
    ctx.clear_tasks(); ctx.tree_walk(0.4, rcut, box, 0.5 * (bdl + bdr), bdr - bdl); ctx.build_csr(); ctx.compute()
    got = ctx.download_acc()
    if ref is None:
        rows = np.arange(T.nleaf)[T.leaf_npart[:T.nleaf] > 0]
        sel, a, b, npairs = rows_against_oracle(ctx, T.pos, T.leaf_npart[:T.nleaf], T.leaf_ipart[:T.nleaf], rows, box, mass, eps, rs)
        ref = (sel, a, b)
        row_csr, col_csr = ctx.download_csr(raw=True)
        far_csr, near_csr = ctx.download_csr_class()
    sel, a, b = ref
    d = np.linalg.norm(got[sel] - a, axis=1)
    na = np.linalg.norm(a, axis=1)
    e1 = d / np.maximum(na, na.mean())
    e2 = d / np.maximum(np.linalg.norm(b, axis=1), 1e-300)
    k = int(np.argmax(e1))
    res[name] = (e1, e2)
    print(f"{name:18s} e1 max {e1.max():.2e} (particle {sel[k]}, |a|/mean {na[k] / na.mean():.2f}, e2 there {e2[k]:.2e})  e1 p99.9 {np.percentile(e1, 99.9):.2e} median {np.median(e1):.2e}", flush=True)
# the worst particle of the default variant: where does its force come from?
e1 = res["v2 default"][0]
for k in np.argsort(-e1)[:3]:
    i = int(sel[k])
    r = int(np.searchsorted(T.leaf_ipart[:T.nleaf], i, side="right") - 1)
    srcs, cls = col_csr[row_csr[r]:row_csr[r + 1]], far_csr[row_csr[r]:row_csr[r + 1]]
    print(f"particle {i} row {r} nt {T.leaf_npart[r]} sources: {len(srcs)} leaves ({int(cls.sum())} far), errors by variant: " +
          ", ".join(f"{n}: {v[0][k]:.2e}" for n, v in res.items()))
    for c, cname in ((0, "near"), (1, "far")):
        sl = srcs[cls == c]
        if len(sl) == 0:
            continue
        pidx = np.concatenate([np.arange(T.leaf_ipart[s], T.leaf_ipart[s] + T.leaf_npart[s]) for s in sl])
        dx = T.pos[pidx] - T.pos[i]
        dx -= box * np.round(dx / box)
        rr = np.sqrt((dx ** 2).sum(1)); u = rr / (2 * rs)
        f = (erfc(u) + 2 / np.sqrt(np.pi) * u * np.exp(-u * u)) / np.maximum(rr, eps) ** 3
        tot = np.linalg.norm((dx * f[:, None]).sum(0))
        line = f"   {cname:4s}: {len(pidx)} sources, |sum| {tot * mass:.3e}; by u-bin |sum| (count): "
        for lo, hi in ((0, .25), (.25, .5), (.5, 1), (1, 1.5), (1.5, 2), (2, 2.5), (2.5, 3), (3, 4), (4, 9)):
            m = (u >= lo) & (u < hi)
            line += f"[{lo},{hi}) {np.linalg.norm((dx[m] * f[m, None]).sum(0)) * mass:.2e} ({int(m.sum())}) "
        print(line)
    print(f"   oracle |a| {np.linalg.norm(a[k]):.3e}  mean |a| {na.mean():.3e}  sum|terms| {np.linalg.norm(b[k]):.3e}  softened pairs (r < eps): "
          f"{int(((np.linalg.norm((T.pos - T.pos[i]) - box * np.round((T.pos - T.pos[i]) / box), axis=1)) < eps).sum() - 1)}")
