"""Kernel tuning sweep (development aid): Zel'dovich-like or clustered box, MAXLEAF 32 and 16, local list; the first-generation
kernel (A/B baseline) against the production kernel's tunings (sources per lane, blocks per SM, far class threshold).
usage: python tools/sweep.py [nside] [--clustered]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
import p2p_b200
from p2p_b200 import step, synth
args = [a for a in sys.argv[1:] if not a.startswith("--")]
ns = int(args[0]) if args else 128
pos, box = synth.clustered(ns) if "--clustered" in sys.argv else synth.zeldovich_like(ns)
PEAK = 148 * 128 * 2 * 1965e6
for maxleaf in (32, 16):
    L = step.build_lists(pos, box, maxleaf, ns, periodic=False)
    st = step.ShortRangeStep(0)
    st.upload(L, synth.DEMO_MASS, True)
    nt, npairs = st.ctx.counts()
    print(f"maxleaf {maxleaf}: leaves {L.tree.nleaf} tasks {nt} pairs {npairs} far columns {st.ctx.download_csr_class()[0].mean():.3f}", flush=True)
    base = None
    #          variant, tt, nsrc, blocks / SM, far threshold
    for cfg in [(1, 16, 0, 0, -1.0), (2, 16, 0, 0, -1.0), (2, 32, 1, 3, 0.0), (2, 32, 1, 3, -1.0), (2, 32, 2, 3, -1.0), (2, 32, 1, 4, -1.0),
                (2, 32, 1, 3, 1.5), (2, 32, 1, 3, 2.0)]:
        variant, tt, nsrc, minb, far = cfg
        st.ctx.set_kernel_variant(variant); st.ctx.set_tuning(tt, nsrc, minb); st.ctx.set_far_threshold(far)
        st.ctx.build_csr()
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
        name = "scalar v1" if variant == 1 else ("packed v1" if tt != 32 else "packed v2")
        print(f"  {name} tt {tt:2d} nsrc {nsrc} minb {minb} far {far:4.1f}: {ms:8.3f} ms  {npairs/ms/1e6:7.1f} Gpair/s  "
              f"{npairs/ms*1e3*38/PEAK*100:5.1f}% of FP32 peak  (dev vs scalar {err:.1e})", flush=True)
