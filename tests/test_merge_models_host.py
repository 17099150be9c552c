"""Host models of the merge algorithms (csrc/pg_boxes.cu) restated in numpy and checked against plain sorts and the
sequential greedy loop: the chunked bitonic network and the cluster counting-sort pass of the large-page kernels,
and the Jacobi rounds of the resolve kernels.  They pin the algorithms themselves (which comparators run where,
which counts go where, why rounds reach the greedy result); the kernels are checked on the GPU in
tests/test_gpu_parity.py."""
import numpy as np
import pytest

CH = 8192  # EMIT_SMEM_ELEMS


def _cswap(keys, idx, i, pr):
    """emit_cswap_global / the shared-memory comparator: ascending by (key, idx), vectorised over disjoint pairs."""
    a, b, ia, ib = keys[i], keys[pr], idx[i], idx[pr]
    sw = (b < a) | ((b == a) & (ib < ia))
    keys[i], keys[pr] = np.where(sw, b, a), np.where(sw, a, b)
    idx[i], idx[pr] = np.where(sw, ib, ia), np.where(sw, ia, ib)


def _pairs(n_half, j, k, flip_allowed):
    t = np.arange(n_half)
    i = ((t & ~(j - 1)) << 1) | (t & (j - 1))
    pr = np.where(j == (k >> 1), i ^ (k - 1), i ^ j) if flip_allowed else i ^ j
    return i, pr


def emit_cluster_model(keys, idx, csize, ch=CH):
    """nms_emit_cluster_kernel: chunk-local stages in 'shared memory' (a copy of the chunk), cross-chunk steps on
    the global arrays with the comparators dealt to csize CTAs; elements >= K are virtual +inf (pr >= K: skip)."""
    K = len(keys)
    n2 = 1
    while n2 < K:
        n2 <<= 1
    assert n2 > ch
    nch = n2 // ch
    keys, idx = keys.copy(), idx.copy()

    def local(cb, steps):
        kc = min(ch, K - cb)
        ks, xs = keys[cb: cb + kc].copy(), idx[cb: cb + kc].copy()
        for (k, j, flip) in steps:
            i, pr = _pairs(ch >> 1, j, k, flip)
            m = pr < kc
            _cswap(ks, xs, i[m], pr[m])
        keys[cb: cb + kc], idx[cb: cb + kc] = ks, xs

    full = [(k, j, True) for k in (2 ** e for e in range(1, ch.bit_length())) for j in
            (2 ** f for f in range(k.bit_length() - 2, -1, -1))]
    for rank in range(csize):            # stage 0
        for c in range(rank, nch, csize):
            if c * ch >= K:
                break
            local(c * ch, full)
    k = 2 * ch
    while k <= n2:
        j = k >> 1
        while j >= ch:                   # cross-chunk steps
            i, pr = _pairs(n2 >> 1, j, k, True)
            for rank in range(csize):    # CTA `rank` takes t = rank*1024 + tid, striding by csize*1024
                t = np.arange(n2 >> 1)
                mine = ((t // 1024) % csize) == rank
                m = mine & (pr < K)
                _cswap(keys, idx, i[m], pr[m])
            j >>= 1
        tail = [(k, jj, False) for jj in (2 ** f for f in range(ch.bit_length() - 2, -1, -1))]
        for rank in range(csize):
            for c in range(rank, nch, csize):
                if c * ch >= K:
                    break
                local(c * ch, tail)
        k <<= 1
    return keys, idx


@pytest.mark.parametrize("K,csize,ch", [(300, 8, 64), (257, 4, 64), (1000, 8, 64), (129, 2, 128), (4096 + 17, 8, 256),
                                        (20000, 8, CH)])
def test_chunked_bitonic_network_sorts_by_key_then_index(K, csize, ch):
    rng = np.random.default_rng(K)
    keys = rng.integers(0, max(2, K // 3), K).astype(np.uint64)  # many ties: the index breaks them
    idx = rng.permutation(K).astype(np.int32)
    k2, i2 = emit_cluster_model(keys, idx, csize, ch)
    order = np.lexsort((idx, keys))
    assert np.array_equal(k2, keys[order]) and np.array_equal(i2, idx[order])


def cluster_stable_pass_model(elems, digit_of, ndig, csize):
    """cluster_stable_pass: global warp g = rank*32 + warp owns a contiguous chunk; bases = digit base over the
    cluster + totals of lower CTAs + counts of lower warps of the same CTA."""
    m = len(elems)
    tw = csize * 32
    chunk = (((m + tw - 1) // tw) + 31) & ~31
    hist = np.zeros((csize, 32, ndig), np.int64)
    spans = {}
    for rank in range(csize):
        for warp in range(32):
            lo = min(m, (rank * 32 + warp) * chunk)
            hi = min(m, lo + chunk)
            spans[rank, warp] = (lo, hi)
            np.add.at(hist[rank, warp], digit_of(elems[lo:hi]), 1)
    all_tot = hist.sum(1)                       # [csize, ndig], what every CTA receives through DSMEM
    digit_total = all_tot.sum(0)
    digit_base = np.concatenate([[0], np.cumsum(digit_total)[:-1]])
    dst = np.full(m, -1, np.int64)
    for rank in range(csize):
        before = all_tot[:rank].sum(0)
        running = digit_base + before
        for warp in range(32):
            base = running.copy()
            running = running + hist[rank, warp]
            lo, hi = spans[rank, warp]
            for c in range(lo, hi, 32):         # one warp iteration: ranks inside the match group
                e = elems[c: min(c + 32, hi)]
                d = digit_of(e)
                for lane in range(len(e)):
                    rk = int((d[:lane] == d[lane]).sum())
                    dst[base[d[lane]] + rk] = e[lane]
                np.add.at(base, d, 1)
    return dst


@pytest.mark.parametrize("m,csize", [(0, 8), (1, 8), (31, 2), (1000, 4), (5000, 8), (12345, 8)])
def test_cluster_counting_sort_pass_is_a_stable_sort(m, csize):
    rng = np.random.default_rng(m + csize)
    cell = rng.integers(0, 64 * 128, max(m, 1)).astype(np.int64)[:m]
    ident = np.arange(m)
    tmp = cluster_stable_pass_model(ident, lambda e: cell[e] & 127, 128, csize)
    srt = cluster_stable_pass_model(tmp, lambda e: cell[e] >> 7, 64, csize)
    assert np.array_equal(srt, np.argsort(cell, kind="stable"))


def jacobi_resolve_model(sup):
    """nms_resolve: sup[i] = the boxes that outrank i, share its class and overlap it beyond the threshold.
    Every round reads the state of the previous round only (double-buffered kept/undecided words):
    undecided -> suppressed if a suppressor is kept, -> kept if none of its suppressors is still undecided."""
    n = len(sup)
    kept = np.zeros(n, bool)
    undec = np.ones(n, bool)
    rounds = 0
    while undec.any():
        k2, u2 = kept.copy(), undec.copy()
        for i in np.nonzero(undec)[0]:
            s = sup[i]
            if kept[s].any():
                u2[i] = False
            elif not undec[s].any():
                k2[i], u2[i] = True, False
        kept, undec = k2, u2
        rounds += 1
        assert rounds <= n
    return kept, rounds


@pytest.mark.parametrize("seed", range(5))
def test_jacobi_rounds_reach_the_greedy_fixed_point(seed):
    """The sequential loop of apply_non_max_suppression (3_combine_grids.py:104-136) against the round-based
    resolution, on random boxes with score ties and several classes."""
    from oracle import boxes as ob
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    xy = rng.uniform(0, 300, (n, 2))
    wh = rng.uniform(5, 80, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1)
    scores = rng.choice(np.round(rng.uniform(0.1, 0.9, 40), 2), n)  # heavy ties
    classes = rng.integers(0, 3, n).astype(np.float64)
    thr = float(rng.choice([0.5, 0.2, 0.0]))
    picks = ob.nms_pick_order(boxes.tolist(), scores.tolist(), classes.tolist(), thr)
    sup = []
    for i in range(n):
        s = [j for j in range(n) if j != i and classes[j] == classes[i]
             and (scores[j] > scores[i] or (scores[j] == scores[i] and j < i))
             and ob.iou(boxes[j].tolist(), boxes[i].tolist()) > thr]
        sup.append(np.asarray(s, np.int64))
    kept, rounds = jacobi_resolve_model(sup)
    assert sorted(picks) == np.nonzero(kept)[0].tolist()
    # emit: survivors ordered by (score descending, earlier position first) == the reference's pick order
    order = sorted(np.nonzero(kept)[0].tolist(), key=lambda i: (-scores[i], i))
    assert order == picks
