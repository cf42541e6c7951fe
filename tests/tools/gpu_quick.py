"""Quick GPU sanity run (development aid): demo IC, oracle lists, CUDA forces vs fp64 oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
import oracle, p2p_b200
pos = np.load(os.path.join(ROOT, "tests/golden/demo_lcdm_pos_f32.npy")).astype(np.float64)
box, mass, nside = 100000.0, 211.75382579190332, 32
rs, rcut, eps = oracle.derived_params(box, nside, len(pos))
ctx = p2p_b200.P2PContext(0)
for ml in (8, 16, 32):
    T = oracle.Tree(pos, ml, [0, 0, 0], [box] * 3, 0)
    tt, ts = T.walk_p2p(0.4, rcut)
    for trunc in (True, False):
        ref, npairs = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, mass, eps, rs if trunc else 0.0)
        absr, _ = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, mass, eps, rs if trunc else 0.0, absterms=True)
        amean = np.linalg.norm(ref, axis=1).mean()
        for var in (1, 2):
            ctx.set_physics(mass, eps, rs if trunc else 0.0); ctx.set_kernel_variant(var)
            ctx.upload_particles(T.pos); ctx.upload_leaves(T.leaf_npart[:T.nleaf], T.leaf_ipart[:T.nleaf])
            ctx.clear_tasks(); ctx.append_tasks(tt, ts); ctx.build_csr()
            nt, npg = ctx.counts()
            ctx.compute(); ctx.synchronize()
            for _ in range(3):
                ctx.zero_acc(); ctx.compute()
            acc = ctx.download_acc()
            ms, mscsr = ctx.last_timings()
            d = np.linalg.norm(acc - ref, axis=1)
            print(f"maxleaf {ml:2d} trunc {int(trunc)} variant {var}: tasks {nt} pairs {npg} (oracle {npairs}) "
                  f"err/mean|a| {d.max()/amean:.2e} err/sum|terms| {(d/np.linalg.norm(absr,axis=1)).max():.2e} "
                  f"kernel {ms:.3f} ms -> {npg/ms/1e6:.1f} Gpair/s  csr {mscsr:.3f} ms", flush=True)
