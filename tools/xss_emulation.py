"""numpy's AVX-512 argsort (x86-simd-sort `avx512_argsort` + libstdc++ std::sort fallback), restated in plain Python.

This is the specification the CUDA kernel `k_xss_argsort` (wgdi_b200/csrc/icl_b200.cu) follows; it was written from
the library's algorithm and checked step by step against numpy 2.3.5 on an AVX-512 host:

    python tools/xss_emulation.py          # compares with numpy.argsort(kind='quicksort') on this machine

The comparison only means something where numpy dispatches int64 argsort to its AVX-512 kernels (`is_avx512_numpy`).

  argsort_(left, right, iters = 2 * floor(log2 n))
    iters <= 0          std::sort(arg + left, arg + right + 1, arr[a] < arr[b])          (libstdc++ introsort)
    <= 256 elements     key-index bitonic network over 8-lane registers: padding key = max, padding index = 0;
                        a compare-exchange moves the index with its key and moves nothing on equal keys
    else                pivot = 5th smallest of arr[arg[left + k * ((right - left) // 8)]], k = 1..8
                        partition_unrolled (4 vectors of 8 lanes), then left part if pivot != smallest key seen,
                        right part if pivot != largest key seen, both with iters - 1
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


def argsort_n(arr, arg, left, n):
    N = 256
    while N > 8 and n * 2 <= N:
        N //= 2
    keys = [arr[arg[left + i]] if i < n else MAX for i in range(N)]
    idx = [arg[left + i] if i < n else 0 for i in range(N)]
    network(keys, idx)
    arg[left:left + n] = idx[:n]


def get_pivot(arr, arg, left, right):
    size = (right - left) // 8
    return sorted(arr[arg[left + (j + 1) * size]] for j in range(8))[4]


def partition_unrolled(arr, arg, left, right, pivot, unroll=4):
    """`right` exclusive; returns (pivot index, smallest key, largest key)."""
    smallest, biggest = MAX, -MAX - 1
    for _ in range((right - left) % (unroll * 8)):
        v = arr[arg[left]]
        smallest, biggest = min(smallest, v), max(biggest, v)
        if not v < pivot:
            right -= 1
            arg[left], arg[right] = arg[right], arg[left]
        else:
            left += 1
    if left == right:
        return left, smallest, biggest
    state = [left, right - 8, smallest, biggest]          # l_store, r_store, min, max

    def store(vec):
        keys = [arr[a] for a in vec]
        ge = [a for a, k in zip(vec, keys) if k >= pivot]
        lt = [a for a, k in zip(vec, keys) if not k >= pivot]
        arg[state[0]:state[0] + len(lt)] = lt
        arg[state[1] + 8 - len(ge):state[1] + 8] = ge
        state[0] += len(lt)
        state[1] -= len(ge)
        state[2], state[3] = min(state[2], min(keys)), max(state[3], max(keys))
    vec_left = [list(arg[left + 8 * i:left + 8 * i + 8]) for i in range(unroll)]
    vec_right = [list(arg[right - 8 * (unroll - i):right - 8 * (unroll - i) + 8]) for i in range(unroll)]
    left += unroll * 8
    right -= unroll * 8
    while right != left:
        if (state[1] + 8) - right < left - state[0]:
            right -= unroll * 8
            cur = [list(arg[right + 8 * i:right + 8 * i + 8]) for i in range(unroll)]
        else:
            cur = [list(arg[left + 8 * i:left + 8 * i + 8]) for i in range(unroll)]
            left += unroll * 8
        for vec in cur:
            store(vec)
    for vec in vec_left + vec_right:
        store(vec)
    return state[0], state[2], state[3]


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


def argsort_(arr, arg, left, right, iters):
    if iters <= 0:
        std_sort(arr, arg, left, right + 1)
        return
    if right + 1 - left <= 256:
        argsort_n(arr, arg, left, right + 1 - left)
        return
    pivot = get_pivot(arr, arg, left, right)
    pi, smallest, biggest = partition_unrolled(arr, arg, left, right + 1, pivot)
    if pivot != smallest:
        argsort_(arr, arg, left, pi - 1, iters - 1)
    if pivot != biggest:
        argsort_(arr, arg, pi, right, iters - 1)


def avx512_argsort(x):
    arr = [int(k) for k in x]
    arg = list(range(len(arr)))
    if len(arr) > 1:
        argsort_(arr, arg, 0, len(arr) - 1, 2 * int(math.log2(len(arr))))
    return arg


def is_avx512_numpy():
    """numpy.argsort on this machine behaves like the emulation on a few tie-heavy arrays."""
    rng = np.random.default_rng(99)
    for n in (5, 16, 100, 300, 700):
        x = rng.integers(0, 4, n).astype(np.int64)
        if np.argsort(x, kind="quicksort").tolist() != avx512_argsort(x):
            return False
    return True


def self_check(cases=400, seed=0, verbose=True):
    rng = np.random.default_rng(seed)
    bad = 0
    for t in range(cases):
        n = int(rng.integers(1, 3000)) if t % 3 else int(rng.integers(1, 300))
        nv = int(rng.choice([2, 3, 5, 40, 3000]))
        x = rng.integers(0, nv, n).astype(np.int64)
        if t % 5 == 0:                                   # like DP scores: a mass of 25 / 30 / 50 plus a tail
            x = rng.choice([25, 30, 50], n, p=[.5, .3, .2]).astype(np.int64)
            m = rng.random(n) < 0.3
            x[m] += rng.integers(20, 2000, int(m.sum()))
        bad += np.argsort(x, kind="quicksort").tolist() != avx512_argsort(x)
    if verbose:
        print(f"{cases} arrays, {bad} differ from numpy.argsort(kind='quicksort')")
    return bad


if __name__ == "__main__":
    if not is_avx512_numpy():
        print("numpy on this machine does not sort int64 with its AVX-512 kernels: nothing to compare")
        sys.exit(0)
    sys.exit(1 if self_check() else 0)
