"""The oracle (oracle/p2p_oracle.c) against the golden fixtures the REFERENCE produced
(tests/golden/demo_lists.json, made by tests/golden/make_golden.py from oracle/_ref/ref_lists)."""
import numpy as np
import pytest
from conftest import DEMO_BOX, DEMO_NSIDE, THETA

import flow
import oracle


@pytest.mark.parametrize("maxleaf", [8, 16, 32])
def test_local_list_matches_reference_golden(demo_pos, golden, maxleaf):
    case = next(c for c in golden["cases"] if c["maxleaf"] == maxleaf and c["nproc"] == 1)["ranks"][0]
    rs, rcut, eps = oracle.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    T = oracle.Tree(demo_pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0)
    tt, ts = T.walk_p2p(THETA, rcut)
    assert T.nleaf == case["nleaf"] and T.nnode == case["nnode"]
    assert len(tt) == case["local_tasks"]
    assert int((T.leaf_npart[tt].astype(np.int64) * T.leaf_npart[ts]).sum()) == case["local_pairs"]
    assert "%016x" % oracle.fingerprint_sorted(tt, ts) == case["local_fingerprint"]


def test_survey_counts(golden):
    """SURVEY.md section 6 / BASELINE.md section 2 numbers, reproduced by the reference run."""
    want = {8: (4452, 1105228, 63282334), 16: (2303, 381377, 83354950), 32: (1195, 133155, 109733162)}
    for ml, (nleaf, ntask, npairs) in want.items():
        rk = next(c for c in golden["cases"] if c["maxleaf"] == ml and c["nproc"] == 1)["ranks"][0]
        assert (rk["nleaf"], rk["local_tasks"], rk["local_pairs"]) == (nleaf, ntask, npairs)
    rk = next(c for c in golden["cases"] if c["maxleaf"] == 16 and c["nproc"] == 1)["ranks"][0]
    assert rk["remote_tasks"][0] == 381377                      # zero-shift self exchange duplicates the local list (D6)
    assert sum(rk["remote_tasks"][1:]) == 205240 and sum(rk["remote_pairs"][1:]) == 36821870
    assert rk["remote_tasks"][1:14] == [126, 1664, 112, 2384, 26922, 2137, 98, 1683, 142, 2048, 29227, 2424, 33653]


@pytest.mark.parametrize("nproc", [1, 2, 4, 8])
def test_multirank_flow_matches_reference_golden(demo_pos, golden, nproc):
    case = next(c for c in golden["cases"] if c["maxleaf"] == 16 and c["nproc"] == nproc)
    ranks = flow.short_range_lists(demo_pos, DEMO_BOX, 16, DEMO_NSIDE, THETA, nproc, do_ext=True, literal_d6=True)
    for rk, g in zip(ranks, case["ranks"]):
        T = rk["tree"]
        assert (T.npart, T.nleaf, T.nnode) == (g["npart"], g["nleaf"], g["nnode"])
        tt, ts = rk["local"]
        assert len(tt) == g["local_tasks"]
        assert "%016x" % oracle.fingerprint_sorted(tt, ts) == g["local_fingerprint"]
        assert [len(r["tt"]) for r in rk["remote"]] == g["remote_tasks"]
        pairs = [int((T.leaf_npart[r["tt"]].astype(np.int64) * r["image"]["npart"][r["ts"]]).sum()) for r in rk["remote"]]
        assert pairs == g["remote_pairs"]


def test_oracle_reproduces_the_golden_force_vectors():
    """tests/golden/demo_forces.npz (made by tests/golden/make_golden_forces.py; its plain vector equals what the reference's
    own GPU kernel returns on this list to 1.2e-15, profiles/r2_reference_kernel_r64_vs_oracle.json): the oracle still
    produces it -- same tree, same list, same fp64 arithmetic (libm erfc / exp may differ in the last bits between builds)."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("make_golden_forces", os.path.join(ROOT, "tests", "golden", "make_golden_forces.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    got = m.compute()
    want = np.load(os.path.join(ROOT, "tests", "golden", "demo_forces.npz"))
    assert np.array_equal(got["index"], want["index"])
    assert np.array_equal(got["acc_plain"], want["acc_plain"])                       # sqrt and division only: bit for bit
    scale = np.linalg.norm(want["acc_trunc"], axis=1).mean()
    assert np.abs(got["acc_trunc"] - want["acc_trunc"]).max() < 1e-13 * scale
