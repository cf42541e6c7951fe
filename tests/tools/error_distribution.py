"""Distribution of the per-particle force error of the device path against the fp64 oracle on the identical list
("additionally report the distribution of per-particle relative error"): demo IC, MAXLEAF 16 and 32,
local list + 26 periodic images, truncated kernel, for the production kernel and its variants.  Prints one JSON object.
usage: python tests/tools/error_distribution.py [--variants]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200")]
import flow  # noqa: E402
import p2p_b200  # noqa: E402
from p2p_b200 import step  # noqa: E402

BOX, NSIDE, THETA, MASS = 100000.0, 32, 0.4, 211.75382579190332
pos = np.load(os.path.join(ROOT, "tests", "golden", "demo_lcdm_pos_f32.npy")).astype(np.float64)
out = {"input": "demo IC 32^3, theta 0.4, local list + 26 periodic images, erfc-truncated kernel; reference = fp64 oracle on the same lists"}
ctx = p2p_b200.P2PContext(0)
#           name: (tt, nsrc, blocks / SM, far threshold)
VARIANTS = {"production": (0, 0, 0, -1.0)}
if "--variants" in sys.argv:
    VARIANTS.update({"first_generation_kernel": (16, 0, 0, -1.0), "no_far_class": (32, 1, 3, 0.0), "two_sources_per_lane": (32, 2, 3, -1.0),
                     "far_class_from_u_2": (32, 1, 3, 2.0), "fp64_force_factor_same_fp32_separations": (32, 9, 0, -1.0)})
for maxleaf in (16, 32):
    ref, rt, rp = flow.reference_forces(pos, BOX, maxleaf, NSIDE, THETA, MASS, 1, True)
    absr, _, _ = flow.reference_forces(pos, BOX, maxleaf, NSIDE, THETA, MASS, 1, True, absterms=True)
    for name, (tt, nsrc, minb, far) in VARIANTS.items():
        ctx.set_tuning(tt, nsrc, minb)
        ctx.set_far_threshold(far)
        acc, _, ntask, npairs = step.run_device_step(ctx, pos, BOX, maxleaf, NSIDE, MASS, THETA, periodic=True)
        assert (ntask, npairs) == (rt, rp)
        d = np.linalg.norm(acc - ref, axis=1)
        na = np.linalg.norm(ref, axis=1)
        rel = d / np.maximum(na, 1e-300)
        q = [50, 90, 99, 99.9, 100]
        out[f"maxleaf_{maxleaf}" + ("" if name == "production" else "_" + name)] = {
            "pairs": npairs,
            "rel_err_percentiles": dict(zip(map(str, q), [float(x) for x in np.percentile(rel, q)])),
            "err_over_mean_force_percentiles": dict(zip(map(str, q), [float(x) for x in np.percentile(d / na.mean(), q)])),
            "err_over_sum_abs_terms_percentiles": dict(zip(map(str, q), [float(x) for x in np.percentile(d / np.linalg.norm(absr, axis=1), q)])),
            "e1_max_over_max_of_own_and_mean_force": float((d / np.maximum(na, na.mean())).max()),
            "fraction_above_1e-5_relative_to_own_force": float((rel > 1e-5).mean()),
            "median_ratio_net_force_to_sum_abs_terms": float(np.median(na / np.linalg.norm(absr, axis=1)))}
print(json.dumps(out))
