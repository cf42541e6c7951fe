"""Seeded synthetic particle sets of the BASELINE.json configurations.

Positions are generated in float32 and widened to float64, like the Gadget-2 snapshots the
reference reads (float32 on disk, 1_Indexing/src/snapshot.c:243-259), so that the fp64 oracle and
the FP32 device path see exactly the same coordinates.

Every generator is SLAB-DECOMPOSABLE: particle i of the global array depends only on (seed, i) -- the random numbers come
from one counter-based stream (Philox) per block of 2^20 particles -- so a rank can generate exactly the slab
[N r / P, N (r + 1) / P) it starts from (1_Indexing/src/initial.c:648-692 reads the same slab of the snapshot) without any
rank ever holding the whole box: 1024^3 is 26 GB of fp64 positions."""
import numpy as np

BOX = 100000.0          # h^-1 kpc, the demo box (1_Indexing/demo/lcdm_g2.run)
DEMO_NSIDE = 32
DEMO_MASS = 211.75382579190332
SEED = 20250101
BLOCK = 1 << 20         # particles per random stream
MAX_CLUMP = 1 << 21     # particles in the largest clump of the clustered box (binds from 1024^3 on)


def box_for(nside_particles):
    """Box scaled so that the mean spacing equals the demo's (BOX / 32)."""
    return BOX * nside_particles / DEMO_NSIDE


def _rng(seed, block, stream):
    return np.random.Generator(np.random.Philox(key=[(int(seed) << 8) | stream, int(block)]))


def _wrap32(pos, box):
    pos %= box
    p32 = pos.astype(np.float32)
    p32[p32 >= np.float32(box)] = np.nextafter(np.float32(box), np.float32(0))
    return p32.astype(np.float64)


def _grid(nside, box, i):
    """grid point of global particle index i (x slowest), cell-centred"""
    d = box / nside
    out = np.empty((len(i), 3), np.float64)
    out[:, 0] = (i // (nside * nside) + 0.5) * d
    out[:, 1] = ((i // nside) % nside + 0.5) * d
    out[:, 2] = (i % nside + 0.5) * d
    return out


def _parallel(fn, jobs):
    """blocks are independent streams: generate them on all host threads (numpy releases the GIL in the bulk generators)"""
    import os
    from concurrent.futures import ThreadPoolExecutor
    nthr = int(os.environ.get("P2P_SYNTH_THREADS", "0")) or min(16, os.cpu_count() or 1)
    if len(jobs) <= 1 or nthr <= 1:
        for j in jobs:
            fn(j)
        return
    with ThreadPoolExecutor(nthr) as ex:
        list(ex.map(fn, jobs))


def _blocks(lo, hi):
    for b in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
        b0 = b * BLOCK
        yield b, max(lo, b0) - b0, min(hi, b0 + BLOCK) - b0


def zeldovich_slab(nside, lo, hi, sigma=0.2, seed=SEED, box=None):
    """Particles [lo, hi) of the nside^3 box: grid + Gaussian displacement of rms `sigma` mean spacings per axis
    (z ~ 49 like the demo), wrapped.  Returns (positions float64 [hi - lo, 3], box)."""
    box = box_for(nside) if box is None else box
    d = box / nside
    out = np.empty((hi - lo, 3), np.float64)

    def fill(job):
        b, a, e = job
        o = b * BLOCK + a - lo
        disp = _rng(seed, b, 0).normal(0.0, sigma * d, size=(BLOCK, 3))[a:e]
        i = b * BLOCK + np.arange(a, e, dtype=np.int64)
        out[o:o + e - a] = _wrap32(_grid(nside, box, i) + disp, box)

    _parallel(fill, list(_blocks(lo, hi)))
    return out, box


def zeldovich_like(nside, sigma=0.2, seed=SEED, box=None):
    """the whole box (small sizes only)"""
    return zeldovich_slab(nside, 0, nside ** 3, sigma, seed, box)


class HaloCatalogue:
    """NFW-ish clumps holding `frac` of the particles: power-law occupation, uniform centres (a few 10^5 entries at most)"""

    def __init__(self, nside, box, frac, seed):
        n = nside ** 3
        rng = _rng(seed, 0, 7)
        nh_part = int(frac * n)
        self.nhalo = max(1, nh_part // 2000)
        occ = rng.pareto(1.0, self.nhalo) + 1.0
        self.weight = occ / occ.sum()
        # the Pareto tail can hand one clump most of the box (the 1024^3 draw: 71 %); no clump holds more than 2^21 particles
        cap = MAX_CLUMP / max(nh_part, 1)
        while self.weight.max() > cap * (1 + 1e-9) and cap * self.nhalo > 1.0:
            w = np.minimum(self.weight, cap)
            free = w < cap
            w[free] *= (1.0 - cap * np.count_nonzero(~free)) / w[free].sum()
            self.weight = w
        self.cdf = np.cumsum(self.weight)
        self.cdf[-1] = 1.0
        self.centers = rng.uniform(0, box, size=(self.nhalo, 3))
        d = box / nside
        self.rvir = d * 0.6 * np.maximum(self.weight * nh_part, 1.0) ** (1.0 / 3.0)
        self.frac = frac


def clustered_slab(nside, lo, hi, frac_in_halos=0.3, seed=SEED, box=None, catalogue=None):
    """Particles [lo, hi) of the clustered box: the Zel'dovich-like background, of which every particle moves into a clump
    with probability `frac_in_halos` (clump drawn by occupation, radius from a centrally concentrated profile that never
    comes closer than 0.05 r_vir) -- the load-imbalance configuration."""
    box = box_for(nside) if box is None else box
    pos, _ = zeldovich_slab(nside, lo, hi, 0.2, seed, box)
    H = catalogue or HaloCatalogue(nside, box, frac_in_halos, seed)

    def fill(job):
        b, a, e = job
        o = b * BLOCK + a - lo
        rng = _rng(seed, b, 1)
        u = rng.uniform(0, 1, size=(BLOCK, 3))[a:e]
        v = rng.normal(size=(BLOCK, 3))[a:e]
        m = u[:, 0] < H.frac
        h = np.searchsorted(H.cdf, u[m, 1], side="right").clip(0, H.nhalo - 1)
        r = H.rvir[h] * (0.05 + u[m, 2] ** 1.5)
        vv = v[m] / np.linalg.norm(v[m], axis=1, keepdims=True)
        seg = pos[o:o + e - a]
        seg[m] = _wrap32(H.centers[h] + vv * r[:, None], box)

    _parallel(fill, list(_blocks(lo, hi)))
    return pos, box


def clustered(nside, frac_in_halos=0.3, seed=SEED, box=None):
    return clustered_slab(nside, 0, nside ** 3, frac_in_halos, seed, box)


def uniform(n, box=BOX, seed=SEED):
    rng = np.random.default_rng(seed)
    p32 = rng.uniform(0, box, size=(n, 3)).astype(np.float32)
    p32[p32 >= np.float32(box)] = np.nextafter(np.float32(box), np.float32(0))
    return p32.astype(np.float64), box
