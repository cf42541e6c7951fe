"""Seeded synthetic particle sets of the BASELINE.json configurations.

Positions are generated in float32 and widened to float64, like the Gadget-2 snapshots the
reference reads (float32 on disk, 1_Indexing/src/snapshot.c:243-259), so that the fp64 oracle and
the FP32 device path see exactly the same coordinates."""
import numpy as np

BOX = 100000.0          # h^-1 kpc, the demo box (1_Indexing/demo/lcdm_g2.run)
DEMO_NSIDE = 32
DEMO_MASS = 211.75382579190332
SEED = 20250101


def box_for(nside_particles):
    """Box scaled so that the mean spacing equals the demo's (BOX / 32)."""
    return BOX * nside_particles / DEMO_NSIDE


def zeldovich_like(nside, sigma=0.2, seed=SEED, box=None):
    """Grid + Gaussian displacement of rms `sigma` mean spacings per axis (z ~ 49 like the demo), wrapped."""
    box = box_for(nside) if box is None else box
    d = box / nside
    rng = np.random.default_rng(seed)
    g = (np.arange(nside, dtype=np.float64) + 0.5) * d
    pos = np.empty((nside ** 3, 3), np.float64)
    pos[:, 0] = np.repeat(g, nside * nside)
    pos[:, 1] = np.tile(np.repeat(g, nside), nside)
    pos[:, 2] = np.tile(g, nside * nside)
    pos += rng.normal(0.0, sigma * d, size=pos.shape)
    pos %= box
    p32 = pos.astype(np.float32)
    p32[p32 >= np.float32(box)] = np.nextafter(np.float32(box), np.float32(0))
    return p32.astype(np.float64), box


def clustered(nside, frac_in_halos=0.3, seed=SEED, box=None):
    """Zel'dovich-like background plus NFW-ish clumps holding `frac_in_halos` of the particles
    (power-law halo occupation), the load-imbalance configuration."""
    pos, box = zeldovich_like(nside, 0.2, seed, box)
    n = pos.shape[0]
    rng = np.random.default_rng(seed + 1)
    nh_part = int(frac_in_halos * n)
    nhalo = max(1, nh_part // 2000)
    occ = rng.pareto(1.0, nhalo) + 1.0
    occ = np.maximum(1, np.floor(occ / occ.sum() * nh_part)).astype(np.int64)
    occ[0] += nh_part - occ.sum()
    centers = rng.uniform(0, box, size=(nhalo, 3))
    d = box / nside
    idx = rng.choice(n, nh_part, replace=False)
    o = 0
    for h in range(nhalo):
        m = int(occ[h])
        if m <= 0:
            continue
        rvir = d * 0.6 * m ** (1.0 / 3.0)
        u = rng.uniform(0, 1, m)
        r = rvir * (0.05 + u ** 1.5)             # centrally concentrated profile, never closer than 0.05 rvir
        v = rng.normal(size=(m, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pos[idx[o:o + m]] = centers[h] + v * r[:, None]
        o += m
    pos %= box
    p32 = pos.astype(np.float32)
    p32[p32 >= np.float32(box)] = np.nextafter(np.float32(box), np.float32(0))
    return p32.astype(np.float64), box


def uniform(n, box=BOX, seed=SEED):
    rng = np.random.default_rng(seed)
    p32 = rng.uniform(0, box, size=(n, 3)).astype(np.float32)
    p32[p32 >= np.float32(box)] = np.nextafter(np.float32(box), np.float32(0))
    return p32.astype(np.float64), box
