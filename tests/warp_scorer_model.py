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
