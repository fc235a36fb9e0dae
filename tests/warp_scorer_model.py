"""Scalar model of `subflatten_score_warp_kernel` (csrc/pairs.cu): the same arithmetic, lane by lane, in numpy floats.
Test infrastructure: it documents the kernel's algorithm and lets the CPU suite check its numerics against LAPACK."""
import numpy as np


def score_warp_model(M):
    """M: k x L with 4 < k <= 32, k <= L.  Returns (score, [4 largest eigenvalues of M M^T])."""
    k, L = M.shape
    assert 4 < k <= 32 and k <= L
    G = np.zeros((32, 33))
    for r in range(k):
        for c in range(k):
            acc = 0.0
            for l in range(L):
                acc = acc + M[r, l] * M[c, l]
            G[r, c] = acc
    trace = sum(G[i, i] for i in range(k))
    e, d = np.zeros(32), np.zeros(32)
    for j in range(0, k - 2):  # Householder step j: lane i owns row i
        x = np.array([G[i, j] if j < i < k else 0.0 for i in range(32)])
        s2 = float((x * x).sum())
        aj = x[j + 1]
        alpha = 0.0
        if s2 > 0.0:
            alpha = -np.sqrt(s2) if aj > 0.0 else np.sqrt(s2)
            v = np.array([(aj - alpha) if i == j + 1 else (x[i] if j + 1 < i < k else 0.0) for i in range(32)])
            vn2 = float((v * v).sum())
            if vn2 > 0.0:
                beta = 2.0 / vn2
                p = np.zeros(32)
                for i in range(j + 1, k):
                    acc = 0.0
                    for l in range(j + 1, k):
                        acc += G[i, l] * v[l]
                    p[i] = acc * beta
                K = 0.5 * beta * float((v * p).sum())
                q = p - K * v
                for i in range(j + 1, k):
                    for l in range(j + 1, k):
                        G[i, l] -= v[i] * q[l] + q[i] * v[l]
        e[j] = alpha
    for i in range(k):
        d[i] = G[i, i]
    e[k - 2] = G[k - 1, k - 2]
    off = lambda i: abs(e[i]) if 0 <= i < k - 1 else 0.0  # noqa: E731
    lo = min(d[i] - off(i) - off(i - 1) for i in range(k))
    hi = max(d[i] + off(i) + off(i - 1) for i in range(k))
    lams = []
    for grp in range(4):  # 8 lanes per wanted eigenvalue: 9-section with Sturm counts
        want = k - 1 - grp
        glo, ghi = lo, hi
        for _ in range(18):
            w = (ghi - glo) / 9.0
            t = 0
            for m in range(8):
                xm = glo + w * (m + 1)
                piv = d[0] - xm
                cnt = int(piv < 0.0)
                for i in range(1, k):
                    if piv == 0.0:
                        piv = -1.0e-300
                    piv = d[i] - xm - e[i - 1] * e[i - 1] / piv
                    cnt += int(piv < 0.0)
                t += int(cnt <= want)
            nlo = glo + w * t
            if t < 8:
                ghi = glo + w * (t + 1)
            glo = nlo
        lams.append(0.5 * (glo + ghi))
    top = ((max(lams[0], 0.0) + max(lams[1], 0.0)) + max(lams[2], 0.0)) + max(lams[3], 0.0)
    return (float(np.sqrt(max(trace - top, 0.0) / trace)) if trace > 0 else float("nan")), lams


# ---------------------------------------------------------------------------------------------
# Round-2 scorer (`subflatten_score_warp2_kernel`): Gram matrix from TRIPLE tables, two rows per lane (k <= 64), and a
# division-free Sturm count.  Same arithmetic as the kernel, modelled with numpy floats.
# ---------------------------------------------------------------------------------------------
def triple_tables(T, total):
    """T: [n, n, 4, 4] (= H N_xy H^T).  Returns the tables the kernel gathers from:
       P[x, x', y, c, c'] = sum_{d<3} T[x, y, c, d] T[x', y, c', d]     (3 x 3 block of  S_x S_x'^T  contributed by column taxon y)
       R[x, y, c]         = sum_{d<3} T[x, y, c, d] T[y, y, 3, d]       (last-row / last-column terms)
       D[y]               = sum_{d<3} T[y, y, 3, d]^2
       m[x, c]            = T[x, x, c, 3]"""
    n = T.shape[0]
    P = np.zeros((n, n, n, 3, 3))
    for x in range(n):
        for xp in range(x, n):
            for y in range(n):
                blk = np.zeros((3, 3))
                for c in range(3):
                    for cp in range(3):
                        acc = 0.0
                        for d in range(3):
                            acc = acc + T[x, y, c, d] * T[xp, y, cp, d]
                        blk[c, cp] = acc
                P[x, xp, y] = blk
                P[xp, x, y] = blk.T  # mirrored copy: bitwise symmetric Gram matrices
    R = np.zeros((n, n, 3))
    D = np.zeros(n)
    m = np.zeros((n, 3))
    for x in range(n):
        for c in range(3):
            m[x, c] = T[x, x, c, 3]
        for y in range(n):
            for c in range(3):
                acc = 0.0
                for d in range(3):
                    acc = acc + T[x, y, c, d] * T[y, y, 3, d]
                R[x, y, c] = acc
    for y in range(n):
        acc = 0.0
        for d in range(3):
            acc = acc + T[y, y, 3, d] * T[y, y, 3, d]
        D[y] = acc
    return P, R, D, m


def gram_from_triples(tabs, total, la, lb):
    """G = S S^T of the subflattening with row taxa la (|la| <= |lb| after the caller's orientation), k = 3 |la| + 1."""
    P, R, D, m = tabs
    a = len(la)
    k = 3 * a + 1
    G = np.zeros((k, k))
    for i in range(a):
        for c in range(3):
            r = 3 * i + c
            for ip in range(a):
                acc = np.zeros(3)
                for y in lb:
                    acc = acc + P[la[i], la[ip], y, c]
                for cp in range(3):
                    G[r, 3 * ip + cp] = acc[cp] + m[la[i], c] * m[la[ip], cp]
            acc = 0.0
            for y in lb:
                acc = acc + R[la[i], y, c]
            G[r, 3 * a] = acc + m[la[i], c] * total
            G[3 * a, r] = G[r, 3 * a]
    acc = 0.0
    for y in lb:
        acc = acc + D[y]
    G[3 * a, 3 * a] = acc + total * total
    return G


def sturm_count_division_free(d, e2, k, x):
    """Number of eigenvalues of the symmetric tridiagonal (d, e) below x from the determinant recurrence
    p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}: one sign change per eigenvalue below x.  No division.  (d, e2) are
    scaled so that |d_i - x| <= 2 and e2 <= 1 (no overflow for k <= 64); every 4 steps the pair (p_i, p_{i-1}) is
    multiplied by the power of two that brings the larger one back to [1, 2), as the kernel does with exponent
    arithmetic.  An exact zero counts as positive."""
    import math
    pm, p = 1.0, d[0] - x
    cnt = int(math.copysign(1.0, p) < 0)
    i = 1
    while i + 3 < k:
        for u in range(4):
            pn = (d[i + u] - x) * p - e2[i + u - 1] * pm
            cnt += int((math.copysign(1.0, pn) < 0) != (math.copysign(1.0, p) < 0))
            pm, p = p, pn
        big = max(abs(p), abs(pm))
        if big > 0.0 and big >= 2.2250738585072014e-308:
            sc = 2.0 ** (-math.frexp(big)[1] + 1)
            p, pm = p * sc, pm * sc
        i += 4
    while i < k:
        pn = (d[i] - x) * p - e2[i - 1] * pm
        cnt += int((math.copysign(1.0, pn) < 0) != (math.copysign(1.0, p) < 0))
        pm, p = p, pn
        i += 1
    return cnt


def score_warp2_model(G):
    """G: k x k symmetric PSD, 4 < k <= 64.  Householder tridiagonalisation (a lane owns rows i and i + 32), the four
    largest eigenvalues by 9-section with the division-free Sturm count, score = sqrt((trace - top4) / trace)."""
    G = np.array(G, dtype=np.float64)
    k = G.shape[0]
    assert 4 < k <= 64
    trace = 0.0
    for i in range(k):
        trace += G[i, i]
    e, d = np.zeros(k), np.zeros(k)
    for j in range(0, k - 2):
        x = np.where((np.arange(k) > j), G[:, j], 0.0)
        s2 = float((x * x).sum())
        aj = x[j + 1]
        alpha = 0.0
        if s2 > 0.0:
            alpha = -np.sqrt(s2) if aj > 0.0 else np.sqrt(s2)
            v = x.copy()
            v[j + 1] = aj - alpha
            vn2 = float((v * v).sum())
            if vn2 > 0.0:
                beta = 2.0 / vn2
                p = np.zeros(k)
                p[j + 1:] = (G[j + 1:, j + 1:] @ v[j + 1:]) * beta
                K = 0.5 * beta * float((v * p).sum())
                q = p - K * v
                G[j + 1:, j + 1:] -= np.outer(v[j + 1:], q[j + 1:]) + np.outer(q[j + 1:], v[j + 1:])
        e[j] = alpha
    for i in range(k):
        d[i] = G[i, i]
    e[k - 2] = G[k - 1, k - 2]
    off = lambda i: abs(e[i]) if 0 <= i < k - 1 else 0.0  # noqa: E731
    lo = min(d[i] - off(i) - off(i - 1) for i in range(k))
    hi = max(d[i] + off(i) + off(i - 1) for i in range(k))
    nrm = max(abs(lo), abs(hi))
    inv = 1.0 / nrm if nrm > 0.0 else 1.0
    ds = d * inv
    e2 = (e * inv) * (e * inv)
    lams = []
    for grp in range(4):
        want = k - 1 - grp
        glo, ghi = lo * inv, hi * inv
        for _ in range(18):
            w = (ghi - glo) / 9.0
            t = 0
            for mm in range(8):
                t += int(sturm_count_division_free(ds, e2, k, glo + w * (mm + 1)) <= want)
            nlo = glo + w * t
            if t < 8:
                ghi = glo + w * (t + 1)
            glo = nlo
        lams.append(0.5 * (glo * nrm + ghi * nrm))
    top = ((max(lams[0], 0.0) + max(lams[1], 0.0)) + max(lams[2], 0.0)) + max(lams[3], 0.0)
    return (float(np.sqrt(max(trace - top, 0.0) / trace)) if trace > 0 else float("nan")), lams
