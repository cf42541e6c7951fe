#!/usr/bin/env python
"""Golden FORCE vectors: the fp64 oracle's accelerations on the reference's demo IC (tests/golden/demo_lcdm_pos_f32.npy),
MAXLEAF 16, local list of one rank (381 377 tasks, 83 354 950 pairs -- the counts of demo_lists.json), for a fixed sample of
4096 particles given by their ORIGINAL index in the snapshot:

  demo_forces.npz   index[4096]            original particle indices (seeded sample)
                    acc_plain[4096, 3]     plain Newtonian kernel  (1_Indexing/src/photoNs_CUDA.cu:342-354)
                    acc_trunc[4096, 3]     erfc-truncated kernel   (2_Redundant/src/photoNs_CUDA.cu:432-450)
                    abs_plain, abs_trunc   sum over pairs of |term| (the e2 denominator of the parity tests)

The plain vector is what the reference's OWN GPU kernel produces on this list to 1.2e-15 of the mean force
(tests/tools/ref_gpu_compare.py on a B200, profiles/r2_reference_kernel_r64_vs_oracle.json), which is what pins it.
Needs only the oracle (make -C oracle)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

BOX, NSIDE, MASS, THETA, MAXLEAF = 100000.0, 32, 211.75382579190332, 0.4, 16


def compute():
    pos = np.load(os.path.join(HERE, "demo_lcdm_pos_f32.npy")).astype(np.float64)
    rs, rcut, eps = oracle.derived_params(BOX, NSIDE, len(pos))
    T = oracle.Tree(pos, MAXLEAF, [0, 0, 0], [BOX] * 3, 0)
    tt, ts = T.walk_p2p(THETA, rcut)
    index = np.sort(np.random.default_rng(20250101).choice(len(pos), 4096, replace=False))
    inv = np.empty(len(pos), np.int64)
    inv[T.perm] = np.arange(len(pos))                    # original index -> tree position
    out = {"index": index}
    for name, r in (("plain", 0.0), ("trunc", rs)):
        a, npairs = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, MASS, eps, r)
        b, _ = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, MASS, eps, r, absterms=True)
        assert (len(tt), npairs) == (381377, 83354950)
        out["acc_" + name], out["abs_" + name] = a[inv[index]], b[inv[index]]
    return out


if __name__ == "__main__":
    np.savez(os.path.join(HERE, "demo_forces.npz"), **compute())
