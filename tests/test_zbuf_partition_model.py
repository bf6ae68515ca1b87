"""CPU model of the work split inside k_zb_faces (csrc/nr_raster_zbuf.cu, "my share"): the 256 threads of a CTA
take consecutive whole rows of the CTA's pixel boxes, laid end to end in width order, by equal COST
(rows x (width + ZB_ROW_COST)).  Whatever the boxes, every row must be walked by exactly one thread.
(The kernel itself is checked bit for bit on the GPU; this pins the arithmetic of the split, which no
parity case exercises at its corners: empty CTAs, one giant box, costs that divide evenly, ...)"""
import numpy as np
import pytest

THREADS, ROW_COST = 256, 4


def split(widths, heights):
    """-> list of (thread, face position, row) in the order the kernel walks them."""
    order = np.argsort(widths, kind="stable")                  # counting sort by width (order inside a width: any)
    w = [int(widths[i]) for i in order]
    h = [int(heights[i]) for i in order]
    n = len(w)
    pre = [0] * (THREADS + 1)
    acc = 0
    for t in range(THREADS):
        pre[t] = acc
        if t < n:
            acc += (w[t] + ROW_COST) * h[t]
    pre[THREADS] = acc
    total = acc
    chunk = (total + THREADS - 1) // THREADS
    start, first = [0] * (THREADS + 1), [None] * THREADS
    for t in range(THREADS):
        t0 = min(t * chunk, total)
        st, p, row = total, 0, 0
        if t0 < total:
            lo, hi = 0, n                                      # first j with pre[j] > t0
            while lo < hi:
                mid = (lo + hi) >> 1
                if pre[mid] > t0:
                    hi = mid
                else:
                    lo = mid + 1
            p = lo - 1
            fw = w[p] + ROW_COST
            row = (t0 - pre[p] + fw - 1) // fw
            if row == h[p]:
                p, row = p + 1, 0
            st = pre[p] + row * ((w[p] + ROW_COST) if p < n else 0)
        start[t], first[t] = st, (p, row)
    start[THREADS] = total
    walked = []
    for t in range(THREADS):
        rem = start[t + 1] - start[t]
        assert rem >= 0
        p, row = first[t]
        while rem > 0:
            assert p < n
            walked.append((t, p, row))
            rem -= w[p] + ROW_COST
            row += 1
            if row == h[p]:
                p, row = p + 1, 0
        assert rem == 0, "a thread's share is a whole number of rows"
    return walked, w, h


CASES = {
    "empty": ([], []),
    "one pixel": ([1], [1]),
    "one tall box": ([1], [4096]),
    "one wide box": ([32], [128]),
    "all equal": ([4] * 256, [4] * 256),
    "evenly divisible": ([4] * 256, [8] * 256),
    "two sizes": ([1] * 128 + [32] * 128, [1] * 128 + [7] * 128),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_every_row_is_walked_once(name):
    widths, heights = CASES[name]
    walked, w, h = split(np.array(widths, dtype=int), np.array(heights, dtype=int))
    rows = [(p, r) for _, p, r in walked]
    want = [(p, r) for p in range(len(w)) for r in range(h[p])]
    assert rows == want


def test_random_boxes():
    rng = np.random.RandomState(0)
    for trial in range(200):
        n = int(rng.randint(0, 257))
        widths = rng.randint(1, 33, size=n)
        heights = np.minimum(rng.geometric(0.2, size=n), 4096 // np.maximum(widths, 1))
        heights = np.maximum(heights, 1)
        walked, w, h = split(widths, heights)
        rows = [(p, r) for _, p, r in walked]
        assert rows == [(p, r) for p in range(len(w)) for r in range(h[p])]
        # the split is by cost: no thread gets more than its share plus one row
        cost = {}
        for t, p, r in walked:
            cost[t] = cost.get(t, 0) + w[p] + ROW_COST
        if cost:
            total = sum(cost.values())
            chunk = (total + THREADS - 1) // THREADS
            assert max(cost.values()) <= chunk + 32 + ROW_COST
