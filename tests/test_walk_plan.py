"""Chunked walk plan (walk/compute pipeline): the union of the chunks is the task multiset of the
one-shot walk, and every chunk only holds targets of its own row range (CPU test)."""
import numpy as np
import pytest
from conftest import DEMO_BOX, DEMO_NSIDE, THETA

from p2p_b200 import host


@pytest.mark.parametrize("maxleaf,nchunks", [(8, 5), (16, 8), (32, 1), (32, 64)])
def test_chunks_partition_the_list(demo_pos, maxleaf, nchunks):
    rs, rcut, eps = host.derived_params(DEMO_BOX, DEMO_NSIDE, len(demo_pos))
    T = host.LocalTree(demo_pos, maxleaf, [0, 0, 0], [DEMO_BOX] * 3, 0)
    tt, ts = T.walk_task_p2p(THETA, rcut)
    plan = T.walk_plan(THETA, rcut, nchunks)
    assert plan.nchunks >= 1
    bounds = [plan.rows(c) for c in range(plan.nchunks)]
    assert bounds[0][0] == 0 and bounds[-1][1] == T.nleaf and all(bounds[i][1] == bounds[i + 1][0] for i in range(len(bounds) - 1))
    at, as_ = [], []
    for c, (b, e) in enumerate(bounds):
        a, s = plan.run(c, nthreads=3)
        assert len(a) == 0 or (a.min() >= b and a.max() < e)
        at.append(a)
        as_.append(s)
    at, as_ = np.concatenate(at), np.concatenate(as_)
    k1 = np.sort(tt.astype(np.int64) * T.nleaf + ts)
    k2 = np.sort(at.astype(np.int64) * T.nleaf + as_)
    assert np.array_equal(k1, k2)
