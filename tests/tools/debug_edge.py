import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"))
import oracle, p2p_b200
rng = np.random.default_rng(2)
box = 64.0
pos = rng.uniform(0, box, (3000, 3))
pos[100:140] = pos[100]
pos[200:260] = pos[200] + rng.normal(0, 0.01, (60, 3))
pos = pos.astype(np.float32).astype(np.float64)
eps, rs, mass = 0.05, 2.5, 3.0
ctx = p2p_b200.P2PContext(0)
for maxleaf in (1, 3, 8, 32):
    T = oracle.Tree(pos, maxleaf, [0, 0, 0], [box] * 3, 0)
    tt, ts = T.walk_p2p(0.4, 4.5 * rs)
    keep = np.ones(len(tt), bool); keep[rng.integers(0, len(tt), len(tt) // 3)] = False
    tt, ts = tt[keep], ts[keep]
    ref, npairs = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, mass, eps, rs)
    absr, _ = oracle.p2p(T.pos, T.leaf_npart, T.leaf_ipart, T.pos, T.leaf_npart, T.leaf_ipart, tt, ts, mass, eps, rs, absterms=True)
    for variant in (1, 2):
        for tune in ((0,0,0),(8,1,3),(8,2,4),(16,1,4),(16,2,4),(32,1,3)):
            ctx.set_kernel_variant(variant); ctx.set_tuning(*tune)
            ctx.set_physics(mass, eps, rs); ctx.set_box([0,0,0], box)
            ctx.upload_particles(T.pos); ctx.upload_leaves(T.leaf_npart[:T.nleaf], T.leaf_ipart[:T.nleaf])
            ctx.clear_tasks(); ctx.append_tasks(tt, ts); ctx.build_csr(); ctx.compute()
            acc = ctx.download_acc()
            d = np.linalg.norm(acc - ref, axis=1); na = np.linalg.norm(absr, axis=1)
            rel = d / np.maximum(na, 1e-300)
            i = int(np.argmax(rel))
            print(f"maxleaf {maxleaf} variant {variant} tune {tune}: max rel {rel.max():.2e} at tree idx {i} (orig {T.perm[i]}) d {d[i]:.3e} abs {na[i]:.3e} nbad(>1e-5) {(rel>1e-5).sum()} counts {ctx.counts()} vs {len(tt)},{npairs}", flush=True)
