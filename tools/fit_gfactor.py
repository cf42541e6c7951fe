#!/usr/bin/env python
"""Minimax fit of the PM-split truncation factor used by the P2P kernel.

  g(u) = erfc(u) + 2/sqrt(pi) u exp(-u^2),  u = r / (2 r_s)      (2_Redundant/src/photoNs_CUDA.cu:443-446)

is evaluated on the device as  g(u) = exp(-u^2) * Q(u),  Q(u) = erfcx(u) + 2/sqrt(pi) u, with Q a
degree-N polynomial with q0 = 1, q1 = 0, q2 = 1 pinned (exact Taylor terms: Newtonian limit to O(u^3)).  One MUFU.EX2 plus N FFMA per pair,
no cancellation at large u (g -> 0 with the exponential), no upper clamp needed.
The fit minimises max (u + 0.05) |dg| over u in [0, UMAX]: measured on the demo lists (120 target
leaves, MAXLEAF 16/32) this weighting gives the smallest realised force error per degree, because
the error of a target is a coherent sum over the many far pairs, not over the few close ones:
  degree 8: 2.5e-6, degree 9: 5.2e-7, degree 10: 1.8e-7 of mean|a| (flat weighting: 5.4e-6 / 1.4e-6 / 4.0e-7).

Far field (leaf pairs whose particles are all at u >= U_FAR, csrc/csr_pack.cuh classifies them): in the kernel's
length unit (exp(-u^2) = 2^(-w), w = r'^2) the force factor is
    g(u) / r'^3 = 2^(-w) * H,   H = (erfcx(u) + 2 u / sqrt(pi)) / w^(3/2),
and H is a smooth function of the shifted reciprocal t = 1 / (w + 1/2) (the shift is the one of the continued fraction of
erfcx, which makes the expansion converge fast): a degree-6 polynomial in t = MUFU.RCP(w + 1/2) replaces the rsqrt, the
softening clamp and the degree-10 polynomial -- 17 FP32 instructions per pair instead of 22, to 1e-7 RELATIVE accuracy
(a massive clump a few r_s away contributes coherently, so the far tail needs relative, not absolute, accuracy).
Writes photons-2.0_gpu-p2p-redundancy_b200/csrc/p2p_gcoef.h.
"""
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 40
UMAX = 5.0
DEGREES = (8, 9, 10, 11)


def Q_exact(u):
    u = mp.mpf(u)
    return mp.erfc(u) * mp.e ** (u * u) + 2 / mp.sqrt(mp.pi) * u


def lawson(x, f, wt, deg, iters=200):
    """weighted minimax of f ~ 1 + x^2 + sum_{k=3..deg} c_k x^k via Lawson iteration.
    q0 = 1, q1 = 0, q2 = 1 are the exact Taylor coefficients of Q (g = 1 - 4/(3 sqrt(pi)) u^3 + ...),
    pinned so that close pairs keep the Newtonian force to O(u^3) whatever the far-field weighting."""
    t = x / x.max()
    V = np.stack([t ** k for k in range(3, deg + 1)], axis=1)
    rhs = f - 1.0 - x * x
    lw = np.ones_like(x)
    best = None
    for _ in range(iters):
        W = np.sqrt(lw) * wt
        c, *_ = np.linalg.lstsq(V * W[:, None], rhs * W, rcond=None)
        err = np.abs(wt * (V @ c - rhs))
        if best is None or err.max() < best[1]:
            best = (c.copy(), err.max())
        lw = lw * (err / err.max() + 1e-4)
        lw /= lw.sum()
    c, e = best
    coef = np.concatenate([[1.0, 0.0, 1.0], c / x.max() ** np.arange(3, deg + 1)])
    return coef, e


U_FAR = 1.25         # classification threshold (tight leaf bounds at least 2 r_s U_FAR apart)
U_FAR_FIT = 1.2      # the fit covers a little more than the classification admits
U_FAR_MAX = 6.0      # beyond: exp(-u^2) < 3e-16
U_FAR_REL = 4.3      # relative accuracy up to here (a massive clump at u ~ 4 can still dominate a force), absolute beyond
FAR_SHIFT = 0.5      # t = 1 / (w + FAR_SHIFT)
FAR_DEGREE = 6


def H_far_exact(w):
    """force factor without the exponential, kernel units: g(u) / r'^3 * 2^w = (erfcx(u) + 2 u / sqrt(pi)) / w^(3/2)"""
    w = mp.mpf(w)
    u = mp.sqrt(w / mp.log(mp.e, 2))
    return (mp.erfc(u) * mp.e ** (u * u) + 2 / mp.sqrt(mp.pi) * u) / (w * mp.sqrt(w))


def fit_far(deg=FAR_DEGREE, shift=FAR_SHIFT, u_lo=U_FAR_FIT):
    """weighted minimax of H(t), t = 1 / (w + shift): RELATIVE error for u <= U_FAR_REL, tapering with g(u) beyond.
    Returns (coefficients h[k] of t^k, max weighted relative error)."""
    c2 = float(mp.log(mp.e, 2))
    u = np.linspace(u_lo, U_FAR_MAX, 3001)
    w = c2 * u * u
    H = np.array([float(H_far_exact(x)) for x in w])
    gfun = lambda x: float(mp.erfc(x) + 2 / mp.sqrt(mp.pi) * x * mp.e ** (-x * x))
    g = np.array([gfun(x) for x in u])
    wt = np.minimum(1.0, g / gfun(U_FAR_REL)) / H
    t = 1.0 / (w + shift)
    tmax = t.max()
    V = np.stack([(t / tmax) ** k for k in range(deg + 1)], axis=1)
    lw = np.ones_like(u)
    best = None
    for _ in range(800):
        W = np.sqrt(lw) * wt
        cf, *_ = np.linalg.lstsq(V * W[:, None], H * W, rcond=None)
        err = np.abs(wt * (V @ cf - H))
        if best is None or err.max() < best[1]:
            best = (cf.copy(), err.max())
        lw = lw * (err / err.max() + 1e-4)
        lw /= lw.sum()
    cf, _ = best
    h = cf / tmax ** np.arange(deg + 1)
    h32 = h.astype(np.float32).astype(np.float64)
    e = np.abs(wt * (np.polyval(h32[::-1], t) - H)).max()
    return h, e


def main():
    us = np.linspace(0.0, UMAX, 6001)
    Q = np.array([float(Q_exact(u)) for u in us])
    wt = np.exp(-us * us) * (us + 0.05)
    out = []
    for deg in DEGREES:
        coef, e = lawson(us, Q, wt, deg)
        c32 = coef.astype(np.float32).astype(np.float64)
        dg = np.exp(-us * us) * (np.polyval(c32[::-1], us) - Q)
        print(f"deg {deg}: weighted err {e:.3e}  max|dg| {np.abs(dg).max():.3e}  max|dg|/(u^2+0.02) "
              f"{np.abs(dg / (us * us + 0.02)).max():.3e}", file=sys.stderr)
        out.append((deg, coef, np.abs(dg).max()))
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "photons-2.0_gpu-p2p-redundancy_b200", "csrc", "p2p_gcoef.h")
    with open(dst, "w") as f:
        f.write("// GENERATED by tools/fit_gfactor.py -- do not edit.\n")
        f.write("// g(u) = exp(-u^2) * Q(u), Q(u) = sum_k q[k] u^k ~ erfcx(u) + 2/sqrt(pi) u on [0, %.1f], q[0] = 1.\n" % UMAX)
        f.write("#pragma once\n")
        for deg, coef, e in out:
            f.write(f"// degree {deg}: max |dg| = {e:.3e}\n")
            f.write(f"static const double P2P_GCOEF_{deg}[{deg + 1}] = {{\n")
            f.write(",\n".join(f"    {c:.17e}" for c in coef))
            f.write("\n};\n")
        h, e = fit_far()
        print(f"far field: degree {FAR_DEGREE} in t = 1 / (w + {FAR_SHIFT}), u >= {U_FAR_FIT}: max weighted relative error {e:.3e}", file=sys.stderr)
        f.write("// far field (leaf pairs classified at u >= P2P_U_FAR): g(u) / r'^3 = 2^(-w) H(t), t = 1 / (w + P2P_FAR_SHIFT), w = r'^2 = u^2 log2(e);\n")
        f.write("// fitted for u in [%.2f, %.1f]: relative error <= %.2e for u <= %.1f, tapering with g(u) beyond\n" % (U_FAR_FIT, U_FAR_MAX, e, U_FAR_REL))
        f.write("#define P2P_U_FAR %.17g\n" % U_FAR)
        f.write("#define P2P_FAR_SHIFT %.17g\n" % FAR_SHIFT)
        f.write("#define P2P_FAR_DEGREE %d\n" % FAR_DEGREE)
        f.write(f"static const double P2P_GFAR[{FAR_DEGREE + 1}] = {{\n")
        f.write(",\n".join(f"    {c:.17e}" for c in h))
        f.write("\n};\n")

if __name__ == "__main__":
    main()
