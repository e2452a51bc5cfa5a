"""numpy's AVX-512 and AVX2 argsort (x86-simd-sort `avx512_argsort` / `avx2_argsort` + libstdc++ std::sort fallback),
restated in plain Python.

This is the specification the CUDA kernel `k_xss_argsort` (wgdi_b200/csrc/icl_b200.cu) follows; it was written from
the library's algorithm and checked step by step against numpy 2.3.5 on an AVX-512 host:

    python tools/xss_emulation.py          # compares with numpy.argsort(kind='quicksort') on this machine

The comparison only means something where numpy dispatches int64 argsort to its AVX-512 kernels (`is_avx512_numpy`) or,
for the 4-lane flavour, to its AVX2 kernels (`is_avx2_numpy`; on an AVX-512 host: NPY_DISABLE_CPU_FEATURES with every
AVX512* feature).

  keys already in non-decreasing order: the identity (found by experiment: numpy returns arange(n) for every such array,
  and only at the top level — a sorted sub-range is still permuted by its partitions)
  argsort_(left, right, iters = 2 * floor(log2 n))
    iters <= 0          std::sort(arg + left, arg + right + 1, arr[a] < arr[b])          (libstdc++ introsort)
    <= 256 elements     key-index bitonic network over 8-lane registers: padding key = max, padding index = 0;
                        a compare-exchange moves the index with its key and moves nothing on equal keys
    else                pivot = 5th smallest of arr[arg[left + k * ((right - left) // 8)]], k = 1..8
                        partition_unrolled (4 vectors of 8 lanes), then left part if pivot != smallest key seen,
                        right part if pivot != largest key seen, both with iters - 1

AVX2 (`lanes=4`): the same on 4-lane vectors — 4 samples at left + k * ((right - left) // 4), pivot = the 3rd smallest;
groups of 4 x 4 = 16 elements in the partition — with one difference that shows in the tie order: AVX2 has no
compress-store, the library permutes the vector (keys < pivot first, keys >= pivot last) and stores it whole at both
ends, and its permutation table puts the keys >= pivot in REVERSED lane order.
"""
import math
import sys

import numpy as np

MAX = np.iinfo(np.int64).max


def network(keys, idx):
    """Bitonic sort with flip merges over P = len(keys) slots (P a power of two >= 8), in place."""
    P = len(keys)

    def pair(i, j):
        if keys[j] < keys[i]:
            keys[i], keys[j] = keys[j], keys[i]
            idx[i], idx[j] = idx[j], idx[i]
    S = 2
    while S <= P:
        for g in range(0, P, S):
            for k in range(S // 2):
                pair(g + k, g + S - 1 - k)
        m = S // 4
        while m >= 1:
            for i in range(P):
                if not i & m:
                    pair(i, i | m)
            m //= 2
        S *= 2


def argsort_n(arr, arg, left, n, lanes=8):
    N = 256
    while N > lanes and n * 2 <= N:
        N //= 2
    keys = [arr[arg[left + i]] if i < n else MAX for i in range(N)]
    idx = [arg[left + i] if i < n else 0 for i in range(N)]
    network(keys, idx)
    arg[left:left + n] = idx[:n]


def get_pivot(arr, arg, left, right, lanes=8):
    size = (right - left) // lanes
    return sorted(arr[arg[left + (j + 1) * size]] for j in range(lanes))[lanes // 2]


def partition_unrolled(arr, arg, left, right, pivot, unroll=4, lanes=8):
    """`right` exclusive; returns (pivot index, smallest key, largest key)."""
    W = lanes
    smallest, biggest = MAX, -MAX - 1
    for _ in range((right - left) % (unroll * W)):
        v = arr[arg[left]]
        smallest, biggest = min(smallest, v), max(biggest, v)
        if not v < pivot:
            right -= 1
            arg[left], arg[right] = arg[right], arg[left]
        else:
            left += 1
    if left == right:
        return left, smallest, biggest
    state = [left, right - W, smallest, biggest]          # l_store, r_store, min, max

    def store(vec):
        keys = [arr[a] for a in vec]
        ge = [a for a, k in zip(vec, keys) if k >= pivot]
        lt = [a for a, k in zip(vec, keys) if not k >= pivot]
        arg[state[0]:state[0] + len(lt)] = lt
        arg[state[1] + W - len(ge):state[1] + W] = ge[::-1] if W == 4 else ge
        state[0] += len(lt)
        state[1] -= len(ge)
        state[2], state[3] = min(state[2], min(keys)), max(state[3], max(keys))
    vec_left = [list(arg[left + W * i:left + W * i + W]) for i in range(unroll)]
    vec_right = [list(arg[right - W * (unroll - i):right - W * (unroll - i) + W]) for i in range(unroll)]
    left += unroll * W
    right -= unroll * W
    while right != left:
        if (state[1] + W) - right < left - state[0]:
            right -= unroll * W
            cur = [list(arg[right + W * i:right + W * i + W]) for i in range(unroll)]
        else:
            cur = [list(arg[left + W * i:left + W * i + W]) for i in range(unroll)]
            left += unroll * W
        for vec in cur:
            store(vec)
    for vec in vec_left + vec_right:
        store(vec)
    return state[0], state[2], state[3]


def partition_merge_path(arr, arg, left, right, pivot, unroll=4, lanes=8):
    """`partition_unrolled` without the sequential dependence between groups — what the CTA-cooperative phase of
    `k_xss_argsort` computes.  With lt(g) = keys < pivot in group g, the rule "take the next group from the right iff
    (r_store + W) - right < left - l_store" reads: sum of lt over the groups taken from the right so far < sum of (G - lt)
    over the groups taken from the left so far.  A[i] / B[j] = those sums after i groups from the left / j from the right;
    both only grow, so the groups are taken in the order of the MERGE of A and B (A first on ties), and the two store
    pointers at a group's turn follow from its place in the merge.  Same result as `partition_unrolled`
    (`check_merge_path`)."""
    import bisect
    W = lanes
    G = unroll * W
    smallest, biggest = MAX, -MAX - 1
    for _ in range((right - left) % G):
        v = arr[arg[left]]
        smallest, biggest = min(smallest, v), max(biggest, v)
        if not v < pivot:
            right -= 1
            arg[left], arg[right] = arg[right], arg[left]
        else:
            left += 1
    if left == right:
        return left, smallest, biggest
    L0, R0 = left, right
    M = (R0 - L0) // G
    src = list(arg[L0:R0])
    keys = [arr[a] for a in src]
    smallest, biggest = min(smallest, min(keys)), max(biggest, max(keys))
    lt = [sum(1 for k in keys[g * G:(g + 1) * G] if k < pivot) for g in range(M)]
    out = [None] * (R0 - L0)

    def put(g, l_store, r_store):                       # store pointers relative to L0
        for vv in range(unroll):
            vec = src[g * G + vv * W:g * G + (vv + 1) * W]
            ks = keys[g * G + vv * W:g * G + (vv + 1) * W]
            ge = [a for a, k in zip(vec, ks) if k >= pivot]
            lo = [a for a, k in zip(vec, ks) if not k >= pivot]
            out[l_store:l_store + len(lo)] = lo
            out[r_store + W - len(ge):r_store + W] = ge[::-1] if W == 4 else ge
            l_store += len(lo)
            r_store -= len(ge)
    m = M - 2                                           # the first and the last group are set aside and stored last
    A, B = [0] * (m + 1), [0] * (m + 1)
    for i in range(m):
        A[i + 1] = A[i] + (G - lt[1 + i])               # left candidate i is group 1 + i
    for j in range(m):
        B[j + 1] = B[j] + lt[M - 2 - j]                 # right candidate j is group M - 2 - j
    for g in range(1, M - 1):
        i = g - 1
        j = bisect.bisect_left(B, A[i], 0, m + 1)       # right groups taken before left candidate i: B[j'] < A[i]
        if i + j >= m:                                  # the merge ends before it gets here from the left
            j = M - 2 - g
            i = bisect.bisect_right(A, B[j], 0, m + 1)  # left groups taken before right candidate j: A[i'] <= B[j]
        LT = (i * G - A[i]) + B[j]
        GE = A[i] + (j * G - B[j])
        put(g, LT, (R0 - L0) - W - GE)
    LTm = sum(lt[1:M - 1])
    GEm = m * G - LTm
    put(0, LTm, (R0 - L0) - W - GEm)
    put(M - 1, LTm + lt[0], (R0 - L0) - W - GEm - (G - lt[0]))
    arg[L0:R0] = out
    return L0 + sum(lt), smallest, biggest


def check_merge_path(trials=300, seed=1):
    """`partition_merge_path` against the sequential `partition_unrolled`, both vector widths; returns the mismatches."""
    rng = np.random.default_rng(seed)
    bad = 0
    for trial in range(trials):
        n = int(rng.integers(64, 3000))
        kind = trial % 4
        x = (rng.integers(0, 5, n) if kind == 0 else rng.integers(0, 1000, n) if kind == 1 else
             np.sort(rng.integers(0, 50, n)) if kind == 2 else rng.choice([25, 30, 50], n))
        arr = [int(k) for k in x]
        for lanes in (8, 4):
            pivot = int(rng.choice(x))
            a1 = [int(k) for k in rng.permutation(n)]
            a2 = list(a1)
            r1 = partition_unrolled(arr, a1, 0, n, pivot, lanes=lanes)
            r2 = partition_merge_path(arr, a2, 0, n, pivot, lanes=lanes)
            bad += r1 != r2 or a1 != a2
    return bad


def std_sort(arr, arg, first, last):
    """libstdc++ std::sort(arg + first, arg + last, [arr](a, b) { return arr[a] < arr[b]; })."""
    def less(i, j):
        return arr[arg[i]] < arr[arg[j]]

    def unguarded_linear_insert(pos):
        val = arg[pos]
        while arr[val] < arr[arg[pos - 1]]:
            arg[pos] = arg[pos - 1]
            pos -= 1
        arg[pos] = val

    def insertion_sort(f, l):
        for i in range(f + 1, l):
            if less(i, f):
                val = arg[i]
                arg[f + 1:i + 1] = arg[f:i]
                arg[f] = val
            else:
                unguarded_linear_insert(i)

    def introsort_loop(f, l, depth):
        while l - f > 16:
            if depth == 0:
                raise NotImplementedError("std::sort depth limit (heap sort)")
            depth -= 1
            a, b, c = f + 1, f + (l - f) // 2, l - 1
            if less(a, b):
                r = b if less(b, c) else c if less(a, c) else a
            elif less(a, c):
                r = a
            elif less(b, c):
                r = c
            else:
                r = b
            arg[f], arg[r] = arg[r], arg[f]
            i, j = f + 1, l
            while True:
                while less(i, f):
                    i += 1
                j -= 1
                while less(f, j):
                    j -= 1
                if not i < j:
                    break
                arg[i], arg[j] = arg[j], arg[i]
                i += 1
            introsort_loop(i, l, depth)
            l = i
    if first == last:
        return
    introsort_loop(first, last, 2 * ((last - first).bit_length() - 1))
    if last - first > 16:
        insertion_sort(first, first + 16)
        for i in range(first + 16, last):
            unguarded_linear_insert(i)
    else:
        insertion_sort(first, last)


def std_sort_equal_pos(f, l, q):
    """Where std::sort(arg + f, arg + l, cmp) leaves the element found at position q when every key of the range is the
    same: no comparison is true, so each level of __introsort_loop moves the middle element to the front, reverses
    [f + 1, l) (the unguarded partition swaps (f + 1 + k, l - 1 - k) until the scans meet) and cuts at
    f + 1 + (l - f - 1) // 2; the final insertion sort moves nothing.  The kernel applies this in one step per element
    (`xs_std_equal_pos`) instead of partitioning level by level; `check_std_sort_equal` compares it with `std_sort`."""
    while l - f > 16:
        mid = f + ((l - f) >> 1)
        q = f + l - mid if q == f else (f if q == mid else f + l - q)
        cut = f + 1 + ((l - f - 1) >> 1)
        if q < cut:
            l = cut
        else:
            f = cut
    return q


def check_std_sort_equal(sizes=tuple(range(1, 300)) + (1000, 1023, 1024, 1025, 4097, 9999)):
    rng = np.random.default_rng(3)
    bad = 0
    for n in sizes:
        arr = [7] * n
        arg = [int(k) for k in rng.permutation(n)]
        a1 = list(arg)
        std_sort(arr, a1, 0, n)
        a2 = [None] * n
        for q in range(n):
            a2[std_sort_equal_pos(0, n, q)] = arg[q]
        bad += a1 != a2
    return bad


def argsort_(arr, arg, left, right, iters, lanes=8):
    if iters <= 0:
        std_sort(arr, arg, left, right + 1)
        return
    if right + 1 - left <= 256:
        argsort_n(arr, arg, left, right + 1 - left, lanes)
        return
    pivot = get_pivot(arr, arg, left, right, lanes)
    pi, smallest, biggest = partition_unrolled(arr, arg, left, right + 1, pivot, lanes=lanes)
    if pivot != smallest:
        argsort_(arr, arg, left, pi - 1, iters - 1, lanes)
    if pivot != biggest:
        argsort_(arr, arg, pi, right, iters - 1, lanes)


def avx512_argsort(x, lanes=8):
    arr = [int(k) for k in x]
    arg = list(range(len(arr)))
    if len(arr) > 1 and not all(a <= b for a, b in zip(arr, arr[1:])):   # already sorted (ties included): untouched
        argsort_(arr, arg, 0, len(arr) - 1, 2 * int(math.log2(len(arr))), lanes)
    return arg


def avx2_argsort(x):
    return avx512_argsort(x, lanes=4)


# ---- passes without progress, in closed form (what the CUDA kernel does instead of replaying them) -------------------
# pivot == smallest key of the range: every element goes to the "right" side, the pass is a fixed permutation of the
# positions [left, R) and argsort_ comes back to the same range with iters - 1.

def pass_forward(q, left, R, W):
    """Position after one such pass of the element at position q."""
    G = 4 * W
    pre = (R - left) % G
    rend = R - pre
    if pre:
        if q == left:
            q = R - 1
        elif q >= rend:
            q = left if q == rend else q - 1
    if q >= rend:
        return q
    M = (rend - left) // G
    g, r = divmod(q - left, G)
    j, t = divmod(r, W)
    k = g if M == 2 else 0 if g == 1 else M - 2 if g == 0 else M - 1 if g == M - 1 else M - 1 - g
    return rend - k * G - (j + 1) * W + (t if W == 8 else W - 1 - t)


def pass_forward_t(q, t, left, R, W):
    """Position after t passes: groups 2 .. M-2 stay where they are and only reverse their vectors (AVX2: and lanes), an
    involution; groups 0, 1, M-1 and the rotated elements move among themselves and are followed pass by pass."""
    G = 4 * W
    rend = R - (R - left) % G
    M = (rend - left) // G
    g, r = divmod(q - left, G)
    if q < rend and 2 <= g <= M - 2:
        if t % 2 == 0:
            return q
        j, l = divmod(r, W)
        return left + g * G + (3 - j) * W + (l if W == 8 else W - 1 - l)
    for _ in range(t):
        q = pass_forward(q, left, R, W)
    return q


def pass_inverse(d, left, R, W):
    """Position before the pass of the element now at d."""
    G = 4 * W
    pre = (R - left) % G
    rend = R - pre
    if d >= rend:
        return left if d == R - 1 else d + 1
    M = (rend - left) // G
    k, r = divmod(rend - 1 - d, G)
    j, s = r // W, W - 1 - r % W
    t = s if W == 8 else W - 1 - s
    g = k if M == 2 else 1 if k == 0 else 0 if k == M - 2 else M - 1 if k == M - 1 else M - 1 - k
    q = left + g * G + j * W + t
    return rend if pre and q == left else q


def argsort_ff(arr, arg, left, right, iters, lanes=8):
    """argsort_ with runs of no-progress passes applied at once: the sample positions are followed backwards through
    the permutation to get each pass's pivot, the run ends when the pivot moves, the budget runs out or all keys are equal."""
    if iters <= 0:
        std_sort(arr, arg, left, right + 1)
        return
    if right + 1 - left <= 256:
        argsort_n(arr, arg, left, right + 1 - left, lanes)
        return
    W, R = lanes, right + 1
    size = (right - left) // W
    q = [left + (j + 1) * size for j in range(W)]
    keys = sorted(arr[arg[p]] for p in q)
    pivot = keys[W // 2]
    if keys[0] == pivot:
        lo_k = min(arr[arg[i]] for i in range(left, R))
        hi_k = max(arr[arg[i]] for i in range(left, R))
        if pivot == lo_k:
            t = 0
            while True:
                t += 1
                if pivot == hi_k:
                    state = "all equal"
                    break
                if iters - t == 0:
                    state = "out of budget"
                    break
                q = [pass_inverse(p, left, R, W) for p in q]
                pivot = sorted(arr[arg[p]] for p in q)[W // 2]
                if pivot != lo_k:
                    state = "progress"
                    break
            old = arg[left:R]
            for i in range(left, R):
                arg[pass_forward_t(i, t, left, R, W)] = old[i - left]
            iters -= t
            if state == "all equal":
                return
            if state == "out of budget":
                std_sort(arr, arg, left, R)
                return
    pi, smallest, biggest = partition_unrolled(arr, arg, left, R, pivot, lanes=W)
    if pivot != smallest:
        argsort_ff(arr, arg, left, pi - 1, iters - 1, lanes)
    if pivot != biggest:
        argsort_ff(arr, arg, pi, right, iters - 1, lanes)


def check_fast_forward(cases=60, seed=3):
    """The closed forms against the sequential partition, and argsort_ff against argsort_ (no numpy involved)."""
    rng = np.random.default_rng(seed)
    for W in (8, 4):
        for n in list(range(8 * W, 8 * W + 70)) + [257, 300, 1000, 1023, 1024, 1025, 3333]:
            left = int(rng.integers(0, 50))
            R = left + n
            arr = [5] * (R + 2000)
            arg = list(range(1000, 1000 + R))
            old = arg[:]
            assert partition_unrolled(arr, arg, left, R, 5, lanes=W)[0] == left
            for p in range(left, R):
                d = pass_forward(p, left, R, W)
                assert arg[d] == old[p] and pass_inverse(d, left, R, W) == p, (W, n, p)
                for t in (2, 3, 7):
                    e = p
                    for _ in range(t):
                        e = pass_forward(e, left, R, W)
                    assert pass_forward_t(p, t, left, R, W) == e, (W, n, p, t)
        for c in range(cases):
            n = int(rng.integers(257, 4000))
            x = rng.choice([25, 30, 50], n, p=[.8, .15, .05] if c % 2 else [.5, .3, .2])
            m = rng.random(n) < (0.1 if c % 2 else 0.3)
            x[m] += rng.integers(20, 2000, int(m.sum()))
            if c % 5 == 0:
                x = rng.integers(0, int(rng.choice([1, 2, 3])) + 1, n)
            arr = [int(k) for k in x]
            a, b = list(range(n)), list(range(n))
            it = 2 * int(math.log2(n))
            argsort_(arr, a, 0, n - 1, it, W)
            argsort_ff(arr, b, 0, n - 1, it, W)
            assert a == b, (W, c, n)
    return True


def _numpy_is(fn):
    rng = np.random.default_rng(99)
    for n in (5, 16, 100, 300, 700, 1500):
        x = rng.integers(0, 4, n).astype(np.int64)
        if np.argsort(x, kind="quicksort").tolist() != fn(x):
            return False
    return True


def is_avx2_numpy():
    """numpy.argsort on this machine (in this environment) behaves like the 4-lane emulation on a few tie-heavy arrays."""
    return _numpy_is(avx2_argsort)


def is_avx512_numpy():
    """numpy.argsort on this machine behaves like the emulation on a few tie-heavy arrays."""
    return _numpy_is(avx512_argsort)


def self_check(cases=400, seed=0, verbose=True, fn=None):
    fn = fn or avx512_argsort
    rng = np.random.default_rng(seed)
    bad = 0
    for t in range(cases):
        n = int(rng.integers(1, 3000)) if t % 3 else int(rng.integers(1, 300))
        nv = int(rng.choice([2, 3, 5, 40, 3000]))
        x = rng.integers(0, nv, n).astype(np.int64)
        if t % 7 == 3:                                   # constant / sorted with ties / sorted but for one element
            x = np.sort(x) if t % 2 else np.full(n, 7, np.int64)
            if t % 4 == 1 and n > 2:
                x[n // 2] = x[0] - 1
        if t % 5 == 0:                                   # like DP scores: a mass of 25 / 30 / 50 plus a tail
            x = rng.choice([25, 30, 50], n, p=[.5, .3, .2]).astype(np.int64)
            m = rng.random(n) < 0.3
            x[m] += rng.integers(20, 2000, int(m.sum()))
        bad += np.argsort(x, kind="quicksort").tolist() != fn(x)
    if verbose:
        print(f"{cases} arrays, {bad} differ from numpy.argsort(kind='quicksort')")
    return bad


if __name__ == "__main__":
    if is_avx512_numpy():
        sys.exit(1 if self_check() else 0)
    if is_avx2_numpy():
        print("numpy sorts int64 with its AVX2 kernels here")
        sys.exit(1 if self_check(fn=avx2_argsort) else 0)
    print("numpy on this machine sorts int64 with neither its AVX-512 nor its AVX2 kernels: nothing to compare")
    sys.exit(0)
