"""CPU prototype for the next subflattening scorer: Gram -> Householder
tridiagonalisation -> bisection for the 4 largest eigenvalues only, score = sqrt(1 - top4 / trace).
Question answered here: does that meet the parity tolerance max(1e-9, 64 eps / score^2) against the reference's
LAPACK SVD route, including true splits whose score^2 is 1e-5 .. 1e-7?   (numpy / scipy only; not product code)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.linalg
from oracle import splitp_oracle as O
from splitp_b200 import simulation, splits as splits_mod, trees

EPS = np.finfo(float).eps
n, N = 20, 1_000_000
tree = trees.balanced_tree(n, 0.05)
codes = simulation.simulate_codes(tree, simulation.GTR((0.1, 0.2, 0.3, 0.4), (1, 2, 3, 4, 5, 6)), N, seed=3, device="cpu").numpy()
tables, total = O.pair_tables_from_codes(codes)
pos = {t: i for i, t in enumerate(tree.taxa)}
true = set(tree.splits())
allsp = list(splits_mod.all_splits(tree))
rng = np.random.default_rng(0)
sample = [s for s in allsp if s in true] + [allsp[i] for i in rng.choice(len(allsp), 300, replace=False)]
worst = 0.0
rows = []
for s in sample:
    ia, ib = [pos[t] for t in s[0]], [pos[t] for t in s[1]]
    M = O.subflattening_from_tables(tables, total, ia, ib)
    if M.shape[0] > M.shape[1]:
        M = M.T
    if M.shape[0] <= 4:
        continue
    ref = O.split_score(M)
    G = M @ M.T
    T = scipy.linalg.hessenberg(G)           # Householder reduction; symmetric input -> tridiagonal (dsytrd-equivalent arithmetic)
    d, e = np.diag(T).copy(), np.diag(T, -1).copy()
    k = len(d)
    top = scipy.linalg.eigvalsh_tridiagonal(d, e, select="i", select_range=(k - 4, k - 1), lapack_driver="stebz")
    got = np.sqrt(max(1.0 - top.sum() / np.trace(G), 0.0))
    tol = max(1e-9, 64 * EPS / max(ref * ref, 1e-300))
    err = abs(got - ref) / ref if ref > 0 else abs(got - ref)
    rows.append((err / tol, err, tol, ref, s in true, k))
    worst = max(worst, err / tol)
rows.sort(reverse=True)
print(f"{len(rows)} splits; worst error / tolerance = {worst:.3f}")
for r in rows[:8]:
    print("err/tol %.3f  err %.2e  tol %.2e  score %.3e  true=%s k=%d" % r)
