#!/usr/bin/env python
"""Generates the committed golden fixtures from the REFERENCE ITSELF, run in this container:

  demo_lcdm_pos_f32.npy  positions block of /root/reference/1_Indexing/demo/ic_lcdm.gdt2
  demo_lists.json        for MAXLEAF 8/16/32 and P = 1, 2, 4, 8 ranks: what the reference's own
                         tree build + walk_task_p2p + fmm_ext produce on that input (leaf counts, task
                         and pair counts, per-call remote task counts, fingerprints of the sorted lists),
                         captured by oracle/_ref/ref_lists (see oracle/ref_harness/).

Needs /root/reference and a built oracle/_ref (make -C oracle).  The fingerprints use the hash defined
in SURVEY.md section 8c (oracle_fingerprint); the counts reproduce SURVEY.md section 6 / BASELINE.md.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
import refrun  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    pos32, box, mass, npart = refrun.read_gadget2_positions("/root/reference/1_Indexing/demo/ic_lcdm.gdt2")
    np.save(os.path.join(HERE, "demo_lcdm_pos_f32.npy"), pos32)
    pos = pos32.astype(np.float64)
    out = {"input": "1_Indexing/demo/ic_lcdm.gdt2", "npart": int(len(pos)), "box": box, "mass": mass[1], "nside": 32,
           "theta": 0.4, "cases": []}
    for maxleaf in (8, 16, 32):
        for nproc in (1, 2, 4, 8):
            if nproc > 1 and maxleaf != 16:
                continue
            ranks = refrun.run(pos, box, maxleaf, 32, 0.4, do_ext=True, nproc=nproc)
            case = {"maxleaf": maxleaf, "nproc": nproc, "ranks": []}
            for r in ranks:
                nl = r["leaf_npart_ipart"].reshape(-1, 2)
                ts = r["local_tasks_ts"].reshape(-1, 2)
                pairs = int((nl[ts[:, 0], 0].astype(np.int64) * nl[ts[:, 1], 0]).sum())
                rk = {"npart": int(r["npart"][0]), "nleaf": int(len(nl)), "nnode": int(len(r["node_npart_son"]) // 3),
                      "local_tasks": int(len(ts)), "local_pairs": pairs,
                      "local_fingerprint": "%016x" % oracle.fingerprint_sorted(ts[:, 0], ts[:, 1]),
                      "remote_tasks": [], "remote_pairs": []}
                for c in range(1, int(r["n_remote_calls"][0]) + 1):
                    rts = r[f"remote{c}_tasks_ts"].reshape(-1, 2)
                    rn = r[f"remote{c}_node_npart_son"].reshape(-1, 3)
                    first_leaf = int(r["first_leaf"][0])
                    rk["remote_tasks"].append(int(len(rts)))
                    rk["remote_pairs"].append(int((nl[rts[:, 0], 0].astype(np.int64) * rn[rts[:, 1] + first_leaf, 0]).sum()))
                case["ranks"].append(rk)
            out["cases"].append(case)
            print(maxleaf, nproc, [(k["local_tasks"], sum(k["remote_tasks"])) for k in case["ranks"]])
    with open(os.path.join(HERE, "demo_lists.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
