"""Prototype (numpy / pure Python) of the data-parallel formulation of libstdc++'s std::sort used by
lo_sort_segments: Hoare partition expressed with prefix counts, leaves (<=16) finished by a stable rank sort.
Checked against the real std::sort (oracle_std_sort_by_key) on tie-heavy inputs."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob


def median_to_first(A, r, a, b, c):
    lt = lambda i, j: A[i][0] < A[j][0]
    def sw(i, j): A[i], A[j] = A[j], A[i]
    if lt(a, b):
        if lt(b, c): sw(r, b)
        elif lt(a, c): sw(r, c)
        else: sw(r, a)
    elif lt(a, c): sw(r, a)
    elif lt(b, c): sw(r, c)
    else: sw(r, b)


def partition_parallel(A, f, l):
    """range [f, l), pivot at f; returns cut.  Formulation with counts only."""
    p = A[f][0]
    lo, hi = f + 1, l
    keys = np.array([A[t][0] for t in range(lo, hi)])
    isL = keys >= p
    isR = keys <= p
    cL = np.concatenate([[0], np.cumsum(isL)[:-1]])          # L's strictly before t
    totR = isR.sum()
    cR = totR - np.cumsum(isR)                               # R's strictly after t
    swapL = isL & (cR > cL)
    swapR = isR & (cL > cR)
    K = swapL.sum()
    assert K == swapR.sum()
    posL = np.zeros(K, int); posR = np.zeros(K, int)
    for t in range(hi - lo):
        if swapL[t]: posL[cL[t]] = lo + t
        if swapR[t]: posR[cR[t]] = lo + t
    for k in range(K):
        A[posL[k]], A[posR[k]] = A[posR[k]], A[posL[k]]
    nsl = np.nonzero(isL & ~swapL)[0]
    first_non_swap_L = lo + nsl[0] if len(nsl) else 1 << 30
    sr = np.nonzero(swapR)[0]
    min_swap_R = lo + sr[0] if len(sr) else hi
    return min(first_non_swap_L, min_swap_R)


def heap_sort(A, f, l):
    # not needed for the prototype inputs (depth limit never reached on random data) — flag it
    raise RuntimeError("depth limit reached")


def sort_parallel(keys):
    n = len(keys)
    A = [(keys[k], k) for k in range(n)]
    if n <= 1: return [a[1] for a in A]
    lg = n.bit_length() - 1
    stack = [(0, n, 2 * lg)]
    leaves = []
    while stack:
        f, l, depth = stack.pop()
        done = False
        while l - f > 16:
            if depth == 0:
                heap_sort(A, f, l); done = True; break
            depth -= 1
            mid = f + (l - f) // 2
            median_to_first(A, f, f + 1, mid, l - 1)
            cut = partition_parallel(A, f, l)
            stack.append((cut, l, depth))
            l = cut
        if not done:
            leaves.append((f, l))
    out = [None] * n
    for f, l in leaves:
        for t in range(f, l):
            r = f + sum(1 for u in range(f, l) if A[u][0] < A[t][0] or (A[u][0] == A[t][0] and u < t))
            out[r] = A[t][1]
    return out


def main():
    rng = np.random.default_rng(0)
    n_cases = 0
    for trial in range(3000):
        n = int(rng.integers(2, 700))
        mode = trial % 5
        if mode == 0: keys = rng.integers(0, 8, n).astype(np.float64)            # very heavy ties
        elif mode == 1: keys = np.round(rng.normal(size=n) ** 2, 2)              # quantised
        elif mode == 2: keys = rng.uniform(size=n)                               # no ties
        elif mode == 3: keys = np.sort(rng.integers(0, 50, n)).astype(np.float64)  # sorted with ties
        else: keys = np.sort(rng.integers(0, 50, n))[::-1].astype(np.float64)    # reversed
        ref = ob.std_sort_by_key(keys)
        try:
            got = sort_parallel(list(keys))
        except RuntimeError:
            continue
        assert list(ref) == got, (trial, n, mode)
        n_cases += 1
    print("ok", n_cases)


if __name__ == "__main__":
    main()
