"""Small fixed workload for ncu captures: 96^3 Zel'dovich-like box, MAXLEAF 32, local list, default tuning.
usage: profile_run.py [nside] [maxleaf] [tt nsrc minb variant]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
from p2p_b200 import step, synth
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 96
maxleaf = int(sys.argv[2]) if len(sys.argv) > 2 else 32
pos, box = synth.clustered(ns) if 'clustered' in sys.argv else synth.zeldovich_like(ns)
L = step.build_lists(pos, box, maxleaf, ns, periodic=False)
st = step.ShortRangeStep(0)
if len(sys.argv) > 6 and sys.argv[3].isdigit():
    st.ctx.set_tuning(int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])); st.ctx.set_kernel_variant(int(sys.argv[6]))
st.upload(L, synth.DEMO_MASS, True)
for _ in range(3):
    st.ctx.zero_acc(); st.ctx.compute()
st.ctx.synchronize()
nt, npairs = st.ctx.counts()
ms = st.ctx.last_timings()[0]
print(f"nside {ns} maxleaf {maxleaf} tasks {nt} pairs {npairs} kernel {ms:.3f} ms {npairs/ms/1e6:.1f} Gpair/s")
