"""Pure-Python (float64, plain loops) restatement of DPP from sort/dpp_sort.go:372-551 — an independent second
implementation used only to cross-check oracle/oracle.c on small cases.  Follows the same [UNVERIFIED-UPSTREAM]
gonum summation orders as the C oracle."""
import math


def norm2(x):
    scale, sumsq = 0.0, 1.0
    for v in x:
        if v == 0:
            continue
        a = abs(v)
        if scale < a:
            s = scale / a
            sumsq = 1 + sumsq * s * s
            scale = a
        else:
            s = a / scale
            sumsq += s * s
    return scale * math.sqrt(sumsq)


def dot_unitary(x, y):
    s = [0.0, 0.0, 0.0, 0.0]
    n = len(x)
    i = 0
    while i + 4 <= n:
        for j in range(4):
            s[j] += x[i + j] * y[i + j]
        i += 4
    while i < n:
        s[0] += x[i] * y[i]
        i += 1
    return (s[0] + s[2]) + (s[1] + s[3])


def gemm_nt(a, b, serial=False):
    """One element of gonum Dgemm(NoTrans, Trans): k in blocks of 64 (dgemmParallel), or — when the product has fewer than
    four 64 x 64 blocks of C, i.e. at most 64 items — one DotUnitary over all of k (dgemmSerial)."""
    if serial:
        return 0.0 + dot_unitary(a, b)
    c = 0.0
    for k0 in range(0, len(a), 64):
        c += dot_unitary(a[k0:k0 + 64], b[k0:k0 + 64])
    return c


def max_idx(s):
    mx, ind = math.nan, 0
    for i, v in enumerate(s):
        if math.isnan(v):
            continue
        if v > mx or math.isnan(mx):
            mx, ind = v, i
    return ind


def kernel_matrix(emb, rel, alpha, normalize=True):
    n = len(emb)
    F, r = [], []
    c = 1 / math.sqrt(2) if False else 0.70710678118654752440
    for i in range(n):
        f = list(emb[i])
        if normalize:
            s = 1 / norm2(f)
            f = [v * s for v in f]
        f.append(1.0)
        f = [v * c for v in f]
        F.append(f)
        r.append(math.exp(alpha * rel[i]))
    serial = ((n + 63) // 64) ** 2 < 4
    return [[(r[i] * gemm_nt(F[i], F[j], serial)) * r[j] for j in range(n)] for i in range(n)]


def dpp(L, top_n, existed):
    N = len(L)
    top_n = min(top_n, N)
    d2 = [L[i][i] if i not in existed else math.nan for i in range(N)]
    j = max_idx(d2)
    Y = [j]
    C = []
    while len(Y) < top_n:
        dj = d2[j]
        if dj < 1e-10:
            break
        dj = math.sqrt(dj)
        k = len(Y) - 1
        inv = 1 / dj
        if k == 0:
            e = [inv * L[j][i] for i in range(N)]
        else:
            ss = [0.0] * N
            for l in range(k):
                tmp = C[l][j]
                if tmp != 0:
                    for i in range(N):
                        ss[i] += tmp * C[l][i]
            e = [inv * (L[j][i] - ss[i]) for i in range(N)]
        C.append(e)
        d2 = [d2[i] - e[i] * e[i] for i in range(N)]
        d2[j] = math.nan
        j = max_idx(d2)
        Y.append(j)
    if len(Y) < top_n:
        for i in range(N):
            if i not in existed and i not in Y:
                Y.append(i)
                if len(Y) == top_n:
                    break
    return Y


def dpp_with_window(L, top_n, window):
    result = []
    if top_n <= window:
        return dpp(L, top_n, result)
    for _ in range(top_n // window):
        result = result + dpp(L, window, result)
    if top_n % window > 0:
        result = result + dpp(L, top_n % window, result)
    return result


# ---------------------------------------------------------------------------------------------------------------------
# Go's sort.Sort (pdqsort, go1.19+: src/sort/zsortinterface.go) restated a second time, independently of oracle.c and of
# csrc/sort.cu — plain Python over a `less(i, j)` / `swap(i, j)` pair, function for function.
def go_sort_perm(score, descending=True):
    """perm[i] = input index at output position i after sort.Sort(sort.Reverse(ItemScoreSlice)) (descending) or sort.Sort."""
    perm = list(range(len(score)))

    def less(i, j):
        a, b = score[perm[i]], score[perm[j]]
        return b < a if descending else a < b

    def swap(i, j):
        perm[i], perm[j] = perm[j], perm[i]

    def insertion_sort(a, b):
        for i in range(a + 1, b):
            j = i
            while j > a and less(j, j - 1):
                swap(j, j - 1)
                j -= 1

    def sift_down(lo, hi, first):
        root = lo
        while True:
            child = 2 * root + 1
            if child >= hi:
                return
            if child + 1 < hi and less(first + child, first + child + 1):
                child += 1
            if not less(first + root, first + child):
                return
            swap(first + root, first + child)
            root = child

    def heap_sort(a, b):
        first, lo, hi = a, 0, b - a
        for i in range((hi - 1) // 2, -1, -1):
            sift_down(i, hi, first)
        for i in range(hi - 1, -1, -1):
            swap(first, first + i)
            sift_down(lo, i, first)

    def partition(a, b, pivot):
        swap(a, pivot)
        i, j = a + 1, b - 1
        while i <= j and less(i, a):
            i += 1
        while i <= j and not less(j, a):
            j -= 1
        if i > j:
            swap(j, a)
            return j, True
        swap(i, j)
        i += 1
        j -= 1
        while True:
            while i <= j and less(i, a):
                i += 1
            while i <= j and not less(j, a):
                j -= 1
            if i > j:
                break
            swap(i, j)
            i += 1
            j -= 1
        swap(j, a)
        return j, False

    def partition_equal(a, b, pivot):
        swap(a, pivot)
        i, j = a + 1, b - 1
        while True:
            while i <= j and not less(a, i):
                i += 1
            while i <= j and less(a, j):
                j -= 1
            if i > j:
                break
            swap(i, j)
            i += 1
            j -= 1
        return i

    def partial_insertion_sort(a, b):
        max_steps, shortest_shifting = 5, 50
        i = a + 1
        for _ in range(max_steps):
            while i < b and not less(i, i - 1):
                i += 1
            if i == b:
                return True
            if b - a < shortest_shifting:
                return False
            swap(i, i - 1)
            if i - a >= 2:
                for j in range(i - 1, 0, -1):
                    if not less(j, j - 1):
                        break
                    swap(j, j - 1)
            if b - i >= 2:
                for j in range(i + 1, b):
                    if not less(j, j - 1):
                        break
                    swap(j, j - 1)
        return False

    def break_patterns(a, b):
        length = b - a
        if length >= 8:
            rnd = length                                       # xorshift seeded with the length
            modulus = 1 << length.bit_length()                 # nextPowerOfTwo
            idx = a + (length // 4) * 2 - 1
            for i in range(3):
                rnd ^= (rnd << 13) & 0xFFFFFFFFFFFFFFFF
                rnd ^= rnd >> 7
                rnd ^= (rnd << 17) & 0xFFFFFFFFFFFFFFFF
                other = rnd & (modulus - 1)
                if other >= length:
                    other -= length
                swap(idx - 1 + i, a + other)

    def choose_pivot(a, b):
        shortest_ninther, max_swaps = 50, 4 * 3
        n = b - a
        swaps = [0]
        i, j, k = a + n // 4 * 1, a + n // 4 * 2, a + n // 4 * 3

        def order2(x, y):
            if less(y, x):
                swaps[0] += 1
                return y, x
            return x, y

        def median(x, y, z):
            x, y = order2(x, y)
            y, z = order2(y, z)
            x, y = order2(x, y)
            return y

        if n >= 8:
            if n >= shortest_ninther:
                i, j, k = median(i - 1, i, i + 1), median(j - 1, j, j + 1), median(k - 1, k, k + 1)
            j = median(i, j, k)
        if swaps[0] == 0:
            return j, "increasing"
        if swaps[0] == max_swaps:
            return j, "decreasing"
        return j, "unknown"

    def pdqsort(a, b, limit):
        max_insertion = 12
        was_balanced = was_partitioned = True
        while True:
            length = b - a
            if length <= max_insertion:
                insertion_sort(a, b)
                return
            if limit == 0:
                heap_sort(a, b)
                return
            if not was_balanced:
                break_patterns(a, b)
                limit -= 1
            pivot, hint = choose_pivot(a, b)
            if hint == "decreasing":
                i, j = a, b - 1
                while i < j:                                   # reverseRange
                    swap(i, j)
                    i += 1
                    j -= 1
                pivot = (b - 1) - (pivot - a)
                hint = "increasing"
            if was_balanced and was_partitioned and hint == "increasing":
                if partial_insertion_sort(a, b):
                    return
            if a > 0 and not less(a - 1, pivot):
                a = partition_equal(a, b, pivot)
                continue
            mid, already = partition(a, b, pivot)
            was_partitioned = already
            left_len, right_len = mid - a, b - mid
            balance_threshold = length // 8
            if left_len < right_len:
                was_balanced = left_len >= balance_threshold
                pdqsort(a, mid, limit)
                a = mid + 1
            else:
                was_balanced = right_len >= balance_threshold
                pdqsort(mid + 1, b, limit)
                b = mid

    n = len(perm)
    if n > 1:
        pdqsort(0, n, n.bit_length())
    return perm
