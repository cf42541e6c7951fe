"""CPU models of the two non-obvious steps of the device tree build (csrc/device_tree.cuh), checked against the
reference's sequential loops as restated in oracle/p2p_oracle.c (mean_split, 1_Indexing/src/fmm.c:29-77):

  * the closed form of the Hoare-like partition (which elements change place, and the left count with its quirk);
  * the exact parallel evaluation of the SEQUENTIAL fp64 sum (parity transducers composed by a scan).

The CUDA kernels implement exactly these models; their bit-for-bit agreement with the oracle's tree is checked on
the GPU in tests/test_gpu_device_tree.py.
"""
import math
import struct

import numpy as np
import pytest


def ref_partition(x):
    """The reference loop on one coordinate array; returns (permutation applied, np0, mean)."""
    x = list(x)
    idx = list(range(len(x)))
    n = len(x)
    mean = 0.0
    for v in x:
        mean += v
    mean /= float(n)
    but = n - 1
    i = 0
    while i < but:
        if x[i] > mean:
            while x[but] > mean and but > i:
                but -= 1
            x[i], x[but] = x[but], x[i]
            idx[i], idx[but] = idx[but], idx[i]
        i += 1
    return idx, but, mean


def model_partition(x, mean):
    """Closed form used by flag_kernel / split_kernel / slot_kernel / swap_kernel."""
    x = np.asarray(x)
    n = len(x)
    big = x > mean
    G = np.concatenate([[0], np.cumsum(big)])
    nbig = int(G[n])
    np0 = n - 1 if nbig == 0 else n - nbig
    slot = {}
    for i in range(n):
        if i >= np0 and not big[i]:
            cb = int(G[i])
            r = (n - nbig) - (i - cb) - 1
            slot[r] = i
    idx = list(range(n))
    for i in range(np0):
        if big[i]:
            j = slot[int(G[i])]
            idx[i], idx[j] = idx[j], idx[i]
    return idx, np0


@pytest.mark.parametrize("seed", range(6))
def test_partition_closed_form(seed):
    rng = np.random.default_rng(seed)
    for trial in range(400):
        n = int(rng.integers(3, 120))
        kind = trial % 5
        if kind == 0:
            x = rng.random(n)
        elif kind == 1:
            x = rng.integers(0, 4, n).astype(np.float64)          # many ties
        elif kind == 2:
            x = np.full(n, 3.25)                                    # all equal: the unexamined-last-element quirk
        elif kind == 3:
            x = np.sort(rng.random(n))[::-1].copy()                # descending
        else:
            x = np.sort(rng.random(n))
            x[-1] = x[0]                                            # small last element
        ref_idx, ref_np0, mean = ref_partition(x)
        idx, np0 = model_partition(x, mean)
        assert np0 == ref_np0
        assert idx == ref_idx


def seq_sum(x):
    s = 0.0
    for v in x:
        s += float(v)
    return s


def dbits(x):
    return struct.unpack("<q", struct.pack("<d", x))[0]


def from_bits(b):
    return struct.unpack("<d", struct.pack("<q", b))[0]


def model_seq_sum(x, lanes=32, per_lane=8):
    """seq_sum_warp of device_tree.cuh, lane by lane, with Python integers standing in for the int64 registers."""
    x = [float(v) for v in x]
    n = len(x)
    S = 0.0
    pos = 0
    rounds = 0
    while pos < n:
        rounds += 1
        e = ((dbits(S) >> 52) & 0x7FF) - 1023
        okS = S > 0.0 and -900 <= e <= 900
        scale = 2.0 ** (52 - e) if okS else 1.0
        top = 2.0 ** (e + 1) if okS else 0.0
        lane_t = []
        vals = []
        for lane in range(lanes):
            v = [x[pos + lane * per_lane + k] if pos + lane * per_lane + k < n else 0.0 for k in range(per_lane)]
            vals.append(v)
            d = [0, 0]
            p = [0, 1]
            hard = not okS
            for xv in v:
                if not (xv >= 0.0 and xv < top):
                    hard = True
                y = xv * scale
                if not math.isfinite(y):
                    hard = True
                    continue
                qf = math.floor(y)
                r = y - qf
                q = int(qf)
                gt, tie = r > 0.5, r == 0.5
                for s in range(2):
                    t = p[s] ^ (q & 1)
                    c = 1 if (gt or (tie and t)) else 0
                    d[s] += q + c
                    p[s] = t ^ c
            lane_t.append((d, p, hard))
        kS = ((dbits(S) & 0xFFFFFFFFFFFFF) | (1 << 52)) if okS else 0
        par = kS & 1
        # inclusive composition lane by lane (the warp scan computes the same thing in log steps)
        inc = []
        cur_d, cur_p = 0, par
        first_bad = None
        for lane in range(lanes):
            d, p, hard = lane_t[lane]
            cur_d += d[cur_p]
            cur_p = p[cur_p]
            inc.append(cur_d)
            if first_bad is None and (hard or kS + cur_d >= (1 << 53)):
                first_bad = lane
        u = 2.0 ** (e - 52) if okS else 0.0
        if first_bad is None:
            S = u * float(kS + inc[-1])
            pos += lanes * per_lane
            continue
        if first_bad > 0:
            S = u * float(kS + inc[first_bad - 1])
        for xv in vals[first_bad]:
            S += xv
        pos += per_lane * (first_bad + 1)
    return S, rounds


@pytest.mark.parametrize("seed", range(4))
def test_sequential_sum_transducer(seed):
    rng = np.random.default_rng(100 + seed)
    cases = []
    for n in (1, 7, 255, 256, 257, 3000, 20000):
        cases.append(rng.random(n) * 8.0e5)                                             # generic doubles in a box
        cases.append((rng.random(n) * 8.0e5).astype(np.float32).astype(np.float64))    # float32-exact (ties are common)
        cases.append(np.round(rng.random(n) * 64.0) / 64.0)                             # coarse grid: exact ties everywhere
        cases.append(rng.random(n) * 1e-3 + (rng.random(n) < 0.01) * 1e5)               # wide dynamic range
    cases.append(np.zeros(100))
    cases.append(np.concatenate([np.zeros(20), rng.random(600)]))
    cases.append(np.array([5e-324, 1e-310, 3.0, 1e-300] * 100))                         # denormals mixed in
    cases.append(np.array([-1.0, 2.0, 3.0] * 50))                                       # negative values: native path
    for x in cases:
        ref = seq_sum(x)
        got, rounds = model_seq_sum(x)
        assert dbits(got) == dbits(ref), (len(x), got, ref)


def test_transducer_is_parallel_for_long_runs():
    """long runs must mostly take the 256-per-round path, not the 8-per-round fallback"""
    rng = np.random.default_rng(7)
    x = (rng.random(100000) * 8.0e5).astype(np.float32).astype(np.float64)
    got, rounds = model_seq_sum(x)
    assert dbits(got) == dbits(seq_sum(x))
    assert rounds < 100000 / 256 + 80


def ref_route_partition(x, split):
    """bksort_body_inplace as restated in oracle/p2p_oracle.c:split_at (1_Indexing/src/domains.c:163-270), runs >= 3"""
    x = list(x)
    idx = list(range(len(x)))
    n = len(x)
    top = 0
    while top < n and x[top] <= split:
        top += 1
    but = n - 1
    while but >= 0 and x[but] > split:
        but -= 1
    if top == n:
        return idx, n
    if but == -1:
        return idx, 0
    i = top
    while i <= but:
        if x[i] > split:
            x[i], x[but] = x[but], x[i]
            idx[i], idx[but] = idx[but], idx[i]
            while x[but] > split:
                but -= 1
        i += 1
    return idx, (but + 1 if i == but else i)


def model_route_partition(x, split):
    """closed form used by route_flag_kernel / route_split_kernel / slot_kernel / swap_kernel (csrc/device_tree.cuh)"""
    x = np.asarray(x)
    n = len(x)
    big = x > split
    G = np.concatenate([[0], np.cumsum(big)])
    nbig = int(G[n])
    np0 = n - nbig
    slot = {}
    for i in range(n):
        if i >= np0 and not big[i]:
            slot[(n - nbig) - (i - int(G[i])) - 1] = i
    idx = list(range(n))
    for i in range(np0):
        if big[i]:
            j = slot[int(G[i])]
            idx[i], idx[j] = idx[j], idx[i]
    return idx, np0


def test_routing_partition_closed_form():
    rng = np.random.default_rng(3)
    for trial in range(6000):
        n = int(rng.integers(3, 60))
        kind = trial % 4
        if kind == 0:
            x = rng.random(n)
        elif kind == 1:
            x = rng.integers(0, 4, n).astype(np.float64)
        elif kind == 2:
            x = np.full(n, 0.5)
        else:
            x = np.sort(rng.random(n))[::-1].copy()
        split = float(rng.choice([0.5, rng.random(), -1.0, 2.0, 1.0, 0.0]))
        assert ref_route_partition(x, split) == model_route_partition(x, split)


# ---- register bitonic network of the CSR row sort (csrc/csr_pack.cuh: warp_sort_keys) --------------------------------
def model_warp_sort(v):
    """v[e][lane], key i = e * 32 + lane: exchanges at distance j < 32 are a shuffle (lane ^ j), at j >= 32 they pair the
    registers e and e | (j / 32) of one lane -- the index arithmetic of warp_sort_keys, statement for statement"""
    E = len(v)
    lane = np.arange(32)
    k = 2
    while k <= 32 * E:
        j = k >> 1
        while j > 0:
            if j >= 32:
                je = j >> 5
                for e in range(E):
                    if (e & je) == 0:
                        up = ((e << 5) & k) == 0
                        lo, hi = np.minimum(v[e], v[e | je]), np.maximum(v[e], v[e | je])
                        v[e], v[e | je] = (lo, hi) if up else (hi, lo)
            else:
                lower = (lane & j) == 0
                for e in range(E):
                    y = v[e][lane ^ j]
                    up = (((e << 5) | lane) & k) == 0
                    v[e] = np.where(lower == up, np.minimum(v[e], y), np.maximum(v[e], y))
            j >>= 1
        k <<= 1
    return v


@pytest.mark.parametrize("E", [1, 2, 4, 8, 16])
def test_register_bitonic_network_sorts(E):
    rng = np.random.default_rng(E)
    n = 32 * E
    for trial in range(150):
        ln = int(rng.integers(2, n + 1))
        keys = (rng.integers(0, 50, ln) if trial % 3 == 0 else rng.integers(0, 2 ** 32 - 1, ln, dtype=np.uint64)).astype(np.uint32)
        if trial % 5 == 0:
            keys |= np.uint32(0x80000000) * (rng.random(ln) < 0.4).astype(np.uint32)      # class bit: near columns sort last
        a = np.full(n, 0xffffffff, np.uint32)
        a[:ln] = keys
        out = np.concatenate(model_warp_sort([a[e * 32:(e + 1) * 32].copy() for e in range(E)]))
        assert np.array_equal(out[:ln], np.sort(keys))


# ---- node-centric level (csrc/device_tree.cuh: node_level_kernel): the same closed form from ballot masks ---------------
def model_node_level(x, mean):
    """one warp, flags as 32-bit words: the counting and filing of node_level_kernel, statement for statement"""
    x = np.asarray(x)
    n = len(x)
    nword = (n + 31) >> 5
    mask = []
    for c in range(nword):
        m = 0
        for lane in range(32):
            i = c * 32 + lane
            if i < n and x[i] > mean:
                m |= 1 << lane
        mask.append(m)
    nbig = sum(bin(m).count("1") for m in mask)
    np0 = n - 1 if nbig == 0 else n - nbig
    idx = list(range(n))
    if nbig == 0:
        return idx, np0
    nsmall = n - nbig
    big, small_r = {}, {}
    big_before = 0
    for c in range(nword):
        m = mask[c]
        for lane in range(32):
            i = c * 32 + lane
            if i < n:
                bl = big_before + bin(m & ((1 << lane) - 1)).count("1")
                f = (m >> lane) & 1
                if i < np0:
                    if f:
                        big[bl] = i
                elif not f:
                    small_r[nsmall - (i - bl) - 1] = i
        big_before += bin(m).count("1")
    big_left = 0
    c = 0
    while c * 32 < np0:
        rem = np0 - c * 32
        big_left += bin(mask[c] if rem >= 32 else mask[c] & ((1 << rem) - 1)).count("1")
        c += 1
    assert sorted(big) == list(range(big_left)) and sorted(small_r) == list(range(big_left))
    for k in range(big_left):
        i, j = big[k], small_r[k]
        idx[i], idx[j] = idx[j], idx[i]
    return idx, np0


@pytest.mark.parametrize("seed", range(4))
def test_node_level_model_equals_reference_partition(seed):
    rng = np.random.default_rng(100 + seed)
    for trial in range(500):
        n = int(rng.integers(3, 200))
        kind = trial % 5
        if kind == 0:
            x = rng.random(n)
        elif kind == 1:
            x = rng.integers(0, 4, n).astype(np.float64)
        elif kind == 2:
            x = np.full(n, 3.25)
        elif kind == 3:
            x = np.sort(rng.random(n))[::-1].copy()
        else:
            x = np.sort(rng.random(n))
        idx, but, mean = ref_partition(x)
        midx, np0 = model_node_level(x, mean)
        assert (midx, np0) == (idx, but)


# ---- block-centric level (csrc/device_tree.cuh: block_level_kernel): warps own contiguous word ranges -------------------
def model_block_level(x, mean, nwarp):
    """the two sweeps of block_level_kernel with `nwarp` warps: per-warp counts, exclusive scan, ranks from running counts"""
    x = np.asarray(x)
    n = len(x)
    nword = (n + 31) >> 5
    popc = lambda v: bin(v).count("1")
    mask = [0] * nword
    wcnt = [0] * nwarp
    rng_of = lambda w: (nword * w // nwarp, nword * (w + 1) // nwarp)
    for w in range(nwarp):
        for c in range(*rng_of(w)):
            m = 0
            for lane in range(32):
                i = c * 32 + lane
                if i < n and x[i] > mean:
                    m |= 1 << lane
            mask[c] = m
            wcnt[w] += popc(m)
    nbig = sum(wcnt)
    np0 = n - 1 if nbig == 0 else n - nbig
    idx = list(range(n))
    if nbig == 0:
        return idx, np0
    nsmall = n - nbig
    big, small, big_left = {}, {}, 0
    for w in range(nwarp):
        run = sum(wcnt[:w])
        for c in range(*rng_of(w)):
            m = mask[c]
            for lane in range(32):
                i = c * 32 + lane
                if i < n:
                    bl = run + popc(m & ((1 << lane) - 1))
                    f = (m >> lane) & 1
                    if i < np0:
                        if f:
                            big[bl] = i
                    elif not f:
                        small[nsmall - (i - bl) - 1] = i
            rem = np0 - c * 32
            big_left += popc(m if rem >= 32 else (m & ((1 << rem) - 1) if rem > 0 else 0))
            run += popc(m)
    assert sorted(big) == list(range(big_left)) and sorted(small) == list(range(big_left))
    for k in range(big_left):
        i, j = big[k], small[k]
        idx[i], idx[j] = idx[j], idx[i]
    return idx, np0


@pytest.mark.parametrize("nwarp", [1, 3, 16])
def test_block_level_model_equals_reference_partition(nwarp):
    rng = np.random.default_rng(200 + nwarp)
    for trial in range(400):
        n = int(rng.integers(3, 700))
        kind = trial % 5
        if kind == 0:
            x = rng.random(n)
        elif kind == 1:
            x = rng.integers(0, 4, n).astype(np.float64)
        elif kind == 2:
            x = np.full(n, 3.25)
        elif kind == 3:
            x = np.sort(rng.random(n))[::-1].copy()
        else:
            x = np.sort(rng.random(n))
        idx, but, mean = ref_partition(x)
        assert model_block_level(x, mean, nwarp) == (idx, but)
