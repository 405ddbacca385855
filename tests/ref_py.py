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
