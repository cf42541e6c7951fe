"""Kernel tuning sweep (development aid): 128^3 Zel'dovich-like box, MAXLEAF 32 and 16, local list,
all (targets per pass, sources per lane, min blocks, scalar/packed) combinations."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
import p2p_b200
from p2p_b200 import step, synth
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pos, box = synth.zeldovich_like(ns)
for maxleaf in (32, 16):
    L = step.build_lists(pos, box, maxleaf, ns, periodic=False)
    st = step.ShortRangeStep(0)
    st.upload(L, synth.DEMO_MASS, True)
    nt, npairs = st.ctx.counts()
    print(f"maxleaf {maxleaf}: leaves {L.tree.nleaf} tasks {nt} pairs {npairs}", flush=True)
    base = None
    for variant, tt, nsrc, minb in [(1, 16, 2, 4), (2, 16, 1, 3), (2, 16, 1, 4), (2, 16, 2, 3), (2, 16, 2, 4), (2, 16, 4, 3), (2, 16, 4, 4),
                                    (2, 16, 2, 19), (2, 16, 2, 20), (2, 16, 4, 19), (2, 16, 4, 20), (2, 8, 2, 4), (2, 8, 4, 4), (2, 8, 4, 20), (2, 16, 2, 51), (2, 16, 2, 52)]:
        st.ctx.set_kernel_variant(variant); st.ctx.set_tuning(tt, nsrc, minb)
        for _ in range(2):
            st.ctx.zero_acc(); st.ctx.compute()
        ts = []
        for _ in range(3):
            st.ctx.zero_acc(); st.ctx.compute(); ts.append(st.ctx.last_timings()[0])
        a = st.ctx.download_acc()
        if base is None: base = a
        nb = np.linalg.norm(base, axis=1)
        err = (np.linalg.norm(a - base, axis=1) / np.maximum(nb, nb.mean())).max()
        ms = min(ts)
        print(f"  {'packed' if variant==2 else 'scalar'} tt {tt:2d} nsrc {nsrc} minb {minb % 16} poly {minb // 16}: {ms:8.3f} ms  {npairs/ms/1e6:7.1f} Gpair/s  "
              f"{npairs/ms/1e6*38/74449.92*100:5.1f}% of FP32 peak  (dev vs first cfg {err:.1e})", flush=True)
