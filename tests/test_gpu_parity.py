"""GPU parity tests: the CUDA path (through the C-ABI library) against the oracle and the golden
fixtures produced by the unmodified reference.  Run on the B200 box with `-m gpu`.

Tolerances (stated per path):
  * counts, flattenings (sparse / reduced / dense), pair tables from counts: BIT-EXACT.
  * subflattening entries from probabilities: rtol = atol = 1e-14 (the reference's own test uses 1e-15 on a
    4-pattern table; atomics change the summation order of up to P terms).
  * scores, fp64 path: rel <= max(1e-9, 64 * eps / score^2): score^2 = 1 - top4/total cancels, so an fp64
    relative rounding error eps on the eigenvalues becomes eps / score^2 on the radicand.
  * scores, exact-integer tensor-core Gram path: same bound (the Gram itself is exact).
  * rank-1 marginals / divergence: rel <= 1e-12 (tree-ordered fp64 sums against the reference's sequential ones; the
    terms v log(v / (r c)) have both signs, but |sum| is of the order of the largest terms on phylogenetic data).
"""
import ctypes as C
import hashlib
import json

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def score_tol(ref):
    return max(1e-9, 64 * EPS / max(ref * ref, 1e-300))


def assert_score(got, ref, extra=1.0):
    got, ref = float(got), float(ref)
    if ref == 0.0 or not np.isfinite(ref):
        assert got == pytest.approx(ref, abs=1e-7, nan_ok=True)
        return
    assert abs(got - ref) / abs(ref) <= extra * score_tol(ref), (got, ref, abs(got - ref) / abs(ref))


@pytest.fixture(scope="module")
def sp():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import splitp_b200
    return splitp_b200


@pytest.fixture(scope="module")
def eng(sp):
    return sp.engine


def random_codes(n, N, seed, p_bad=0.0):
    rng = np.random.default_rng(seed)
    # a few frequent patterns plus noise, so that counts span 1 .. thousands
    base = rng.integers(0, 4, size=(n, 64))
    pick = rng.integers(0, 64, size=N)
    codes = base[:, pick]
    mut = rng.random((n, N)) < 0.08
    codes = np.where(mut, rng.integers(0, 4, size=(n, N)), codes).astype(np.uint8)
    if p_bad > 0:
        codes[rng.random((n, N)) < p_bad] = 255
    return codes


# ---------------------------------------------------------------------------------------------
# kernel 1: pack + count
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,N,seed,p_bad,force_hash", [
    (4, 5, 0, 0.0, False), (7, 400, 1, 0.02, False), (10, 10_000, 2, 0.0, False), (12, 100_003, 3, 0.001, False),
    (12, 8192 * 3, 4, 0.0, True), (16, 50_001, 5, 0.0005, False), (20, 200_000, 6, 0.0, False), (31, 4097, 7, 0.0, False),
    (1, 33, 8, 0.1, False), (5, 1, 9, 0.0, False), (9, 31, 10, 0.5, True),
])
def test_count_patterns(eng, oracle, n, N, seed, p_bad, force_hash):
    codes = random_codes(n, N, seed, p_bad)
    aln = eng.pack(codes, is_ascii=False)
    tab = eng.count_patterns(aln, want_first=True, force_hash=force_hash)
    keys, counts, usable = oracle.get_pattern_counts_arrays(codes)
    assert int(tab.divisor) == usable
    order = np.argsort(keys)
    got_keys = tab.keys.cpu().numpy().view(np.uint64)
    np.testing.assert_array_equal(got_keys, keys[order])  # ascending = lexicographic order
    np.testing.assert_array_equal(tab.counts.cpu().numpy().view(np.uint32), counts[order])
    # first-occurrence order (dict insertion order of the reference)
    first = tab.first.cpu().numpy().view(np.uint32)
    np.testing.assert_array_equal(got_keys[np.argsort(first, kind="stable")], keys)


def test_count_site_ranges_accumulate(eng, oracle):
    codes = random_codes(11, 70_001, 21, 0.001)
    aln = eng.pack(codes)
    full = eng.count_patterns(aln)
    # two ranges accumulated into one direct table via the C ABI
    lib = eng.lib
    table = torch.zeros(4 ** 11, dtype=torch.int32, device="cuda")
    usable = torch.zeros(1, dtype=torch.int64, device="cuda")
    for b, e in ((0, 33_333), (33_333, 70_001)):
        eng.call("spb_count_direct", eng._p(aln.sm), eng._p(aln.valid), 11, b, e, eng._p(table), None, eng._p(usable), eng._st())
    keys, counts, us = oracle.get_pattern_counts_arrays(codes)
    dense = np.zeros(4 ** 11, dtype=np.int64)
    dense[keys.astype(np.int64)] = counts
    np.testing.assert_array_equal(table.cpu().numpy(), dense)
    assert int(usable.item()) == us == int(full.divisor)


def test_empty_alignment(eng):
    aln = eng.pack(np.zeros((4, 0), np.uint8))
    tab = eng.count_patterns(aln)
    assert tab.num == 0 and tab.divisor == 0.0


def test_fasta_golden(sp, golden_small, tmp_path):
    from splitp_b200.parsers import fasta
    g = golden_small["fasta_small"]
    path = tmp_path / "a.fa"
    path.write_text("".join(f">{k}\n{v}\n" for k, v in g["sequences"].items()))
    probs, L = fasta.pattern_probs_from_alignment(str(path), return_sequence_length=True)
    assert L == g["usable"] and probs == g["probs"] and list(probs) == list(g["probs"])
    assert fasta.pattern_probs_from_alignment(str(path)) == g["probs"]
    # reference tests/test_parsers.py:16-26 expected values
    assert probs == {"ATCG": 2 / 5, "GATC": 1 / 5, "CGAT": 1 / 5, "TCGA": 1 / 5}
    d = golden_small["fasta_dirty"]
    path2 = tmp_path / "b.fa"
    with open(path2, "w") as f:
        for i, s in enumerate(d["rows"]):
            f.write(f">t{i}\n")
            for k in range(0, len(s), 61):
                f.write(s[k:k + 61] + "\n")
    aln = fasta.read_alignment_from_file(str(path2))
    assert list(aln.keys()) == d["names"]
    counts, L = fasta.get_pattern_counts(aln)
    assert L == d["usable"]
    assert list(counts.items()) == [tuple(x) for x in d["counts_items"]]
    assert list(fasta.pattern_counts_to_probs(counts, L).items()) == [tuple(x) for x in d["probs_items"]]


# ---------------------------------------------------------------------------------------------
# kernel 2: flattenings, drop-in surface against the reference's golden vectors
# ---------------------------------------------------------------------------------------------
def test_four_pattern_golden(sp, golden_small):
    g = golden_small["four_pattern"]
    for k, (a, b) in enumerate(g["splits"]):
        split = (set(a), set(b))  # the reference tests pass sets of ints (tests/test_constructions.py:12)
        # NB: with int taxa the default taxa order is sorted(union) = [0,1,2,3]
        F = sp.flattening(split, g["probs"])
        assert type(F).__name__ == "dok_matrix"
        np.testing.assert_array_equal(np.asarray(F.todense()), np.array(g["sparse"][k]))
        np.testing.assert_array_equal(sp.flattening(split, g["probs"], sp.FlatFormat.dense), np.array(g["sparse"][k]))
        R = sp.flattening(split, g["probs"], sp.FlatFormat.reduced)
        np.testing.assert_array_equal(R, np.array(g["reduced"][k]))
        S = sp.subflattening(split, g["probs"])
        np.testing.assert_allclose(S, np.array(g["subflat"][k]), rtol=1e-15, atol=1e-15)
        assert_score(sp.split_score(R), g["scores_reduced"][k])
        assert_score(sp.split_score(S), g["scores_subflat"][k])


def test_subflattening_identity(sp, golden_small):
    """tests/test_constructions.py:90-107 of the reference, run against this engine."""
    g = golden_small["four_pattern"]
    S = np.array([[1, -1], [1, 1]])
    S4 = np.kron(np.kron(S, S), np.kron(S, S))
    idx = [3, 7, 11, 12, 13, 14, 15]
    for a, b in g["splits"]:
        split = (set(a), set(b))
        F = np.asarray(sp.flattening(split, g["probs"]).todense())
        ref = (S4 @ F @ S4.T)[np.ix_(idx, idx)]
        np.testing.assert_allclose(sp.subflattening(split, g["probs"]), ref, rtol=1e-15, atol=1e-15)


def test_quirks(sp, golden_small):
    q = golden_small["quirks"]
    assert sp.flattening("01|23", {"ACGT": 1.0}, "nope") is None
    with pytest.raises(KeyError):
        sp.subflattening("01|23", {"ACGT": 1.0})
    assert float(sp.split_score(np.arange(15.0).reshape(3, 5) + 1)) == q["score_3x5"] == 0.0
    assert float(sp.split_score(np.arange(16.0).reshape(4, 4) ** 1.5 + 1)) == q["score_4x4"] == 0.0
    tab = {"AACG": 0.25, "AATG": 0.5, "CATG": 0.25}
    al3 = sp.Alignment(tab, "0123")
    np.testing.assert_array_equal(np.asarray(sp.flattening((["0"], ["1", "3"]), al3).todense()), np.array(q["partial_cover_sparse"]))
    np.testing.assert_array_equal(sp.flattening((["0"], ["1", "3"]), al3, sp.FlatFormat.reduced), np.array(q["partial_cover_reduced"]))
    np.testing.assert_array_equal(sp.flattening((["0"], ["1", "3"]), al3, sp.FlatFormat.dense), np.array(q["partial_cover_sparse"]))
    with pytest.raises(KeyError):
        sp.flattening("01|23", {"ACGX": 1.0})
    with pytest.raises(KeyError):
        sp.flattening("01|29", sp.Alignment({"ACGT": 1.0}, "0123"))


def test_random_tables_golden(sp, golden_random):
    z, meta = golden_random
    alns = {}
    for rec in meta:
        c = rec["case"]
        if c not in alns:
            pats = [str(p) for p in z[f"c{c}_patterns"]]
            alns[c] = sp.Alignment(dict(zip(pats, z[f"c{c}_values"].tolist())), "0123456789"[:len(pats[0])])
        aln = alns[c]
        a, b = rec["split"].split("|")
        split = (tuple(a), tuple(b))
        D = sp.flattening(split, aln, sp.FlatFormat.dense)
        assert list(D.shape) == rec["dense_shape"] and h(D) == rec["dense_hash"]
        Dk = np.asarray(sp.flattening(rec["split"], aln).todense())
        assert h(Dk) == rec["dense_hash"]
        R = sp.flattening(split, aln, sp.FlatFormat.reduced)
        assert list(R.shape) == rec["reduced_shape"] and h(R) == rec["reduced_hash"]
        if "reduced_key" in rec:
            np.testing.assert_array_equal(R, z[rec["reduced_key"]])
        assert_score(sp.split_score(R), rec["score_reduced"])
        if "subflat_key" in rec:
            Sm = sp.subflattening(split, aln)
            np.testing.assert_allclose(Sm, z[rec["subflat_key"]], rtol=1e-14, atol=1e-14)
            assert_score(sp.split_score(Sm), rec["score_subflat"])


def test_readme_config_golden(sp, golden_readme):
    """Config C1 (README.md:18-42 with random.seed(0)): 100 reduced flattenings bit-exact, scores in tolerance."""
    g = golden_readme
    pats = [str(p) for p in g["patterns"]]
    aln = dict(zip(pats, g["values"].tolist()))  # generate_alignment returns a PLAIN dict
    for s, split in enumerate(g["splits"]):
        a, b = str(split).split("|")
        F = sp.flattening((tuple(a), tuple(b)), aln, sp.FlatFormat.reduced)
        assert tuple(F.shape) == tuple(g["shapes"][s]) and h(F) == str(g["hashes"][s])
        if f"full_{s}" in g.files:
            np.testing.assert_array_equal(F, g[f"full_{s}"])
        assert_score(sp.split_score(F), g["scores"][s])
    alA = sp.Alignment(aln, [str(t) for t in g["taxa"]])
    for s, ref in zip(g["splits"][:6], g["sub_scores"]):
        a, b = str(s).split("|")
        assert_score(sp.split_score(sp.subflattening((tuple(a), tuple(b)), alA)), ref)


def test_sparse_score_and_sub_alignment(sp, oracle, golden_small):
    g = golden_small["sub_alignment"]
    al = sp.Alignment(dict(zip(g["patterns"], g["values"])), g["taxa"])
    for case in g["cases"]:
        sa = al.sub_alignment(tuple(case["sub"]))
        assert list(sa.taxa) == case["taxa"]
        ref = dict((p, x) for p, x in case["items"])
        assert sorted(sa.data) == sorted(ref)
        for p, x in sa.data.items():
            assert x == pytest.approx(ref[p], rel=1e-13)
    assert al.sub_alignment(("e", "b")) is al.sub_alignment(("b", "e"))  # memoised per sorted taxa tuple
    F = sp.flattening("ab|cde", al)
    Fd = np.asarray(F.todense())
    assert_score(sp.split_score(F), oracle.split_score(Fd))



def test_erickson_svd_golden(sp):
    """The batched agglomeration reproduces the reference's erickson_SVD output (phylogenetics.py:99-171) on the
    simulated alignments of tests/golden/golden_erickson.json, for both scoring methods."""
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_erickson.json")) as f:
        cases = json.load(f)
    for rec in cases:
        aln = dict(zip(rec["patterns"], rec["values"]))
        for method in (sp.Method.flattening, sp.Method.subflattening):
            got = [list(map(list, s)) for s in sp.erickson_SVD(aln, method=method)]
            ref = rec[method.name]
            assert len(got) == len(ref) == rec["n"] - 2
            # Splits with a one-taxon side score exactly 0 here, while the reference's 1 - top4/total leaves rounding
            # noise (0, 1e-8 or nan) that decides between them; and two cherries of a balanced tree can score within
            # 1e-10 of each other, so the ORDER of the picks is decided by last-bit noise on both sides.  What is
            # pinned is the set of informative splits that the agglomeration selects.
            informative = lambda seq: sorted({json.dumps(s) for s in seq if min(len(s[0]), len(s[1])) > 1})  # noqa: E731
            assert informative(got) == informative(ref), (rec["n"], method)
        # every non-trivial split the reference found is a true split of the generating tree
        nontrivial = [s for s in rec["flattening"] if min(len(s[0]), len(s[1])) > 1]
        assert all(s in rec["true_splits"] for s in nontrivial)
    assert sp.erickson_SVD(aln, method=sp.Method.distance) is not None  # scores stay at infinity, as in the reference



def test_reduced_flattening_large_side(sp, eng, oracle):
    """A side above 13 taxa cannot use the 4^side flag arrays: the rank computation goes through a sort-unique."""
    n = 16
    codes = random_codes(n, 30_000, 91, 0.0)
    tab = eng.count_patterns(eng.pack(codes))
    keys, counts, usable = oracle.get_pattern_counts_arrays(codes)
    order = np.argsort(keys)
    for ia, ib in (([0, 15], list(range(1, 15))), (list(range(2, 16)), [1, 0])):
        got = eng.flatten_reduced(tab, ia, ib).cpu().numpy()
        ref = oracle.flattening_reduced(keys[order], (counts[order] / usable), n, ia, ib)
        np.testing.assert_array_equal(got, ref)
        gotc = eng.flatten_reduced(tab, ia, ib, as_counts=True).cpu().numpy()
        np.testing.assert_array_equal(gotc, oracle.flattening_reduced(keys[order], counts[order].astype(float), n, ia, ib))


def test_generate_alignment_drop_in(sp, oracle):
    """generate_alignment keeps the reference's return type: a plain dict, keys in lexicographic A<C<G<T order, values
    count / float(sequence_length) (simulation.py:42-56); also above 31 taxa (128-bit keys)."""
    for n, N in ((6, 5000), (40, 3000)):
        tree = sp.trees.balanced_tree(n, 0.05)
        model = sp.simulation.GTR.JukesCantor(0.5)
        aln = sp.generate_alignment(tree, model, N, seed=3)
        assert type(aln) is dict and all(len(p) == n for p in aln)
        assert list(aln) == sorted(aln, key=lambda p: [oracle.STATES.index(c) for c in p])
        codes = sp.simulation.simulate_codes(tree, model, N, seed=3).cpu().numpy()
        ref, usable = oracle.get_pattern_counts_wide(codes)
        assert usable == N and aln == {p: c / float(N) for p, c in ref.items()}


# ---------------------------------------------------------------------------------------------
# kernel 3: pair tables / subflattening, batched
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,N,seed,p_bad", [(6, 3000, 31, 0.0), (8, 40_001, 32, 0.002), (20, 100_000, 33, 0.0), (33, 5000, 34, 0.01)])
def test_pair_tables(eng, oracle, n, N, seed, p_bad):
    codes = random_codes(n, N, seed, p_bad)
    aln = eng.pack(codes)
    pt = eng.pair_tables_from_alignment(aln, as_counts=True)
    valid = (codes <= 3).all(axis=0)
    c = codes[:, valid].astype(np.int64)
    ref = np.zeros((n, n, 4, 4), dtype=np.int64)
    for i in range(n):
        for j in range(n):
            np.add.at(ref[i, j], (c[i], c[j]), 1)
    np.testing.assert_array_equal(pt.N.cpu().numpy(), ref.astype(np.float64))
    H = oracle.H4.astype(np.float64)
    np.testing.assert_array_equal(pt.T.cpu().numpy(), np.einsum("cx,ijxy,dy->ijcd", H, ref.astype(np.float64), H))
    assert float(pt.total.item()) == float(valid.sum())
    if n <= 31:
        tab = eng.count_patterns(aln)
        pt2 = eng.pair_tables_from_table(tab, as_counts=True)
        np.testing.assert_array_equal(pt2.N.cpu().numpy(), ref.astype(np.float64))


def test_subflatten_scores_batched(sp, eng, oracle):
    n, N = 8, 60_000
    tree = sp.trees.balanced_tree(n, 0.08)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR((0.1, 0.2, 0.3, 0.4), (1, 2, 3, 4, 5, 6)), N, seed=5)
    aln = eng.pack(codes)
    pt = eng.pair_tables_from_alignment(aln)  # probabilities
    keys, counts, usable = oracle.get_pattern_counts_arrays(codes.cpu().numpy())
    vals = counts / usable
    splits = list(sp.all_splits(tree))
    idx = [eng.split_positions(s, tree.taxa) for s in splits]
    ma, mb = eng.masks_from_splits(idx)
    got = eng.subflatten_scores(pt, ma, mb).cpu().numpy()
    got_c = eng.subflatten_scores(pt, ma).cpu().numpy()  # complement form
    for s, (ia, ib) in enumerate(idx):
        S = oracle.subflattening(keys, vals, n, ia, ib)
        np.testing.assert_allclose(eng.subflatten(pt, ia, ib).cpu().numpy(), S, rtol=1e-12, atol=1e-14)
        ref = oracle.split_score(S)
        assert_score(got[s], ref)
        assert_score(got_c[s], ref)


# ---------------------------------------------------------------------------------------------
# kernel 4: exact-integer Gram (tensor cores vs the SIMT cross-check) and kernel 5 scores
# ---------------------------------------------------------------------------------------------
def _tiled_from_rowmajor(M):
    """row-major u8 [R, K] -> the 128x128-tile SWIZZLE_128B layout documented in csrc/gram.cu."""
    R, K = M.shape
    out = np.zeros(R * K, dtype=np.uint8)
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    rt, kt, rr, kk = r // 128, k // 128, r % 128, k % 128
    off = (rt * (K // 128) + kt) * 16384 + rr * 128 + (((kk >> 4) ^ (rr & 7)) << 4) + (kk & 15)
    out[off.ravel()] = M.ravel()
    return out


@pytest.mark.parametrize("R,K,density", [(128, 128, 1.0), (128, 1024, 0.3), (256, 256, 1.0), (256, 65536, 0.02), (512, 2048, 0.1),
                                         (1024, 16384, 0.01), (768, 640, 0.5)])
def test_gram_u8_tensor_core_exact(eng, R, K, density):
    rng = np.random.default_rng(R + K)
    M = (rng.integers(0, 256, size=(R, K)) * (rng.random((R, K)) < density)).astype(np.uint8)
    if density == 1.0:
        M[:] = 255  # worst case for the s32 accumulators
    ref = M.astype(np.int64) @ M.astype(np.int64).T
    s0 = torch.from_numpy(_tiled_from_rowmajor(M)).cuda()
    G = torch.empty((R, R), dtype=torch.float64, device="cuda")
    n_ws = int(eng.lib.spb_gram_u8_ws(R, K, 1, 1))
    ws = torch.empty(max(n_ws, 1), dtype=torch.int64, device="cuda")
    eng.call("spb_gram_u8", eng._p(s0), R, K, 1, eng._p(G), eng._p(ws), eng._st())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(G.cpu().numpy(), ref.astype(np.float64))
    G2 = torch.empty_like(G)
    eng.call("spb_gram_u8_simt", eng._p(s0), R, K, 1, eng._p(G2), eng._st())
    np.testing.assert_array_equal(G2.cpu().numpy(), ref.astype(np.float64))


@pytest.mark.parametrize("R,K", [(16, 4096), (16, 1 << 20), (64, 262144), (40, 1000 * 16), (4, 16), (32, 48)])
def test_gram_u8_small_exact(eng, R, K):
    rng = np.random.default_rng(R * 7 + K)
    M = (rng.integers(0, 256, size=(R, K)) * (rng.random((R, K)) < 0.05)).astype(np.uint8)
    M[0, :] = 255
    ref = M.astype(np.int64) @ M.astype(np.int64).T
    k4 = np.ascontiguousarray(M.reshape(R, K // 4, 4).transpose(1, 0, 2))  # [k/4][r][4]: the k4-major layout
    s0 = torch.from_numpy(k4.reshape(-1)).cuda()
    G = torch.empty((R, R), dtype=torch.float64, device="cuda")
    ws = torch.empty(int(eng.lib.spb_gram_u8_ws(R, K, 2, 1)), dtype=torch.int64, device="cuda")
    eng.call("spb_gram_u8", eng._p(s0), R, K, 2, eng._p(G), eng._p(ws), eng._st())
    np.testing.assert_array_equal(G.cpu().numpy(), ref.astype(np.float64))
    G2 = torch.empty_like(G)
    eng.call("spb_gram_u8_simt", eng._p(s0), R, K, 2, eng._p(G2), eng._st())
    np.testing.assert_array_equal(G2.cpu().numpy(), ref.astype(np.float64))


def _count_table(sp, eng, n, N, seed, bl=0.05):
    tree = sp.trees.balanced_tree(n, bl)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), N, seed=seed)
    tab = eng.count_patterns(eng.pack(codes))
    return tree, codes.cpu().numpy(), tab


def test_count_scorer_gram_exact(sp, eng, oracle):
    """F F^T from the u8 + sparse-high decomposition equals the integer Gram bit for bit (dense and reduced)."""
    n = 8
    tree, codes, tab = _count_table(sp, eng, n, 300_000, 41)
    keys, counts, _ = oracle.get_pattern_counts_arrays(codes)
    assert counts.max() >= 256  # the high part is exercised
    scorer = eng.CountScorer(tab)
    for split in [s for s in sp.all_splits(tree)][::7]:
        ia, ib = eng.split_positions(split, tree.taxa)
        for reduced in (False, True):
            G, k = scorer.gram(ia, ib, reduced)
            F = (oracle.flattening_reduced if reduced else oracle.flattening_dense)(keys, counts.astype(np.float64), n, ia, ib)
            if F.shape[0] > F.shape[1] or (not reduced and len(ia) > len(ib)):
                F = F.T
            assert k == F.shape[0]
            ref = F @ F.T  # integers < 2^53: exact in fp64
            np.testing.assert_array_equal(G[:k, :k].cpu().numpy(), ref)
    scorer.check_hi()


@pytest.mark.parametrize("n,N,seed", [(8, 200_000, 51), (10, 100_000, 52)])
def test_count_scorer_scores(sp, eng, oracle, n, N, seed):
    tree, codes, tab = _count_table(sp, eng, n, N, seed)
    keys, counts, usable = oracle.get_pattern_counts_arrays(codes)
    splits = [s for s in sp.all_splits(tree)]
    splits = splits[::max(1, len(splits) // 24)]
    idx = [eng.split_positions(s, tree.taxa) for s in splits]
    got_d = eng.score_splits_counts(tab, idx, reduced=False).cpu().numpy()
    got_r = eng.score_splits_counts(tab, idx, reduced=True).cpu().numpy()
    for s, (ia, ib) in enumerate(idx):
        ref = oracle.split_score(oracle.flattening_reduced(keys, counts / usable, n, ia, ib))
        assert_score(got_d[s], ref)
        assert_score(got_r[s], ref)


def test_split_score_large_fp64(sp, eng, oracle):
    """k > 128 goes through the block-Krylov Rayleigh-Ritz path; compare with LAPACK on the same matrix."""
    n = 10
    tree, codes, tab = _count_table(sp, eng, n, 150_000, 61)
    keys, counts, usable = oracle.get_pattern_counts_arrays(codes)
    true = set(tree.splits())
    fives = [s for s in sp.all_splits(tree) if len(s[0]) == 5]
    picks = [s for s in fives if s in true][:1] + [s for s in fives if s not in true][:2]
    for split in picks:
        ia, ib = eng.split_positions(split, tree.taxa)
        F = oracle.flattening_reduced(keys, counts / usable, n, ia, ib)
        assert min(F.shape) > 128
        assert_score(sp.split_score(F), oracle.split_score(F))
        assert_score(sp.split_score(F.T), oracle.split_score(F))



@pytest.mark.parametrize("k,L,kind", [(200, 220, "sparse"), (300, 300, "gauss"), (500, 2000, "sparse"), (130, 129, "gauss")])
def test_split_score_flat_spectrum(sp, oracle, k, L, kind):
    """Worst case for the restarted block-Krylov solver: no spectral decay after the 4th singular value."""
    rng = np.random.default_rng(k + L)
    A = rng.standard_normal((k, L)) if kind == "gauss" else rng.random((k, L)) * (rng.random((k, L)) < 0.3)
    assert_score(sp.split_score(A), oracle.split_score(A))



# ---------------------------------------------------------------------------------------------
# wide keys (up to 64 taxa, BASELINE config 4)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,N,seed,p_bad", [(64, 20_000, 71, 0.0005), (40, 50_001, 72, 0.0), (33, 777, 73, 0.01), (12, 30_000, 74, 0.0)])
def test_wide_pattern_counts(eng, oracle, n, N, seed, p_bad):
    codes = random_codes(n, N, seed, p_bad)
    if n == 64:
        codes[:, :37] = 3  # the all-T pattern: its key is all ones = the table's EMPTY marker
    wide, valid, n_, N_ = eng.pack_wide(codes)
    tab = eng.count_patterns_wide(wide, valid, n_, N_)
    ref, usable = oracle.get_pattern_counts_wide(codes)
    assert int(tab.divisor) == usable
    got = tab.to_dict()
    assert got == ref
    assert list(got) == sorted(ref, key=lambda p: [oracle.STATES.index(c) for c in p])  # lexicographic A<C<G<T
    # two site ranges merged = the whole alignment (the multi-GPU path on one device)
    t1 = eng.count_patterns_wide(wide, valid, n_, N_, 0, (N // 64) * 32)
    t2 = eng.count_patterns_wide(wide, valid, n_, N_, (N // 64) * 32, N)
    k1, c1 = t1.compact(sort=False)
    k2, c2 = t2.compact(sort=False)
    merged = eng.merge_wide_tables(n_, torch.cat([k1, k2]), torch.cat([c1, c2]), usable)
    assert merged.to_dict() == ref
    # first-occurrence order (the dict order of parsers/fasta.py:48-63)
    tf = eng.count_patterns_wide(wide, valid, n_, N_, want_first=True)
    assert list(tf.to_dict(order="first").items()) == list(ref.items())


def test_fasta_64_taxa(sp, oracle):
    """get_pattern_counts above 31 taxa goes through the 128-bit table and keeps the reference's dict order."""
    from splitp_b200.parsers import fasta
    import collections
    rng = np.random.default_rng(64)
    base = rng.integers(0, 4, size=(64, 30))
    codes = base[:, rng.integers(0, 30, size=4000)]
    codes = np.where(rng.random(codes.shape) < 0.01, rng.integers(0, 4, size=codes.shape), codes)
    chars = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    chars[rng.random(chars.shape) < 0.0005] = ord("N")
    low = rng.random(chars.shape) < 0.3
    chars = np.where(low, chars | 0x20, chars).astype(np.uint8)  # mixed case, upper-cased by the reference (:54)
    aln = collections.OrderedDict((f"t{i}", chars[i].tobytes().decode()) for i in range(64))
    counts, L = fasta.get_pattern_counts(aln)
    ref, usable = oracle.get_pattern_counts_wide(oracle.sequences_to_codes(list(aln.values())))
    assert L == usable and list(counts.items()) == list(ref.items())


@pytest.mark.parametrize("n,N,seed", [(64, 30_000, 81), (36, 100_000, 82), (10, 50_000, 83)])
def test_thin_split_scores(sp, eng, oracle, n, N, seed):
    """Reduced flattenings of 2|n-2, 3|n-3 (and 1|n-1) splits from the hashed wide table: exact Gram, scores vs LAPACK."""
    tree = sp.trees.balanced_tree(n, 0.02)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), N, seed=seed)
    wide, valid, n_, N_ = eng.pack_wide(codes)
    tab = eng.count_patterns_wide(wide, valid, n_, N_)
    ref, usable = oracle.get_pattern_counts_wide(codes.cpu().numpy())
    sides = [[0, 1], [2, 5], [n - 2, n - 1], [1, n // 2], [3], [0, 1, 2], [1, n // 2, n - 1]]
    got = eng.thin_split_scores(tab, sides).cpu().numpy()
    assert np.array_equal(got, eng.thin_split_scores(tab, sides, filtered=False).cpu().numpy())  # the filter changes nothing
    words = int(eng.lib.spb_thin_filter_words(tab.cap))
    filt = torch.empty(2 * words, dtype=torch.int32, device="cuda")
    for s, ia in enumerate(sides):
        ib = [t for t in range(n) if t not in ia]
        F = oracle.flattening_reduced_from_dict(ref, ia, ib)
        G = torch.empty((64, 64), dtype=torch.float64, device="cuda")
        eng.call("spb_thin_gram_wide", eng._p(tab.hkeys), eng._p(tab.hcounts), tab.cap, eng._p(tab.special), n, bytes(ia), len(ia),
                 eng._p(G), eng._st())
        G2 = torch.empty((64, 64), dtype=torch.float64, device="cuda")
        eng.call("spb_thin_gram_wide_filtered", eng._p(tab.hkeys), eng._p(tab.hcounts), tab.cap, eng._p(tab.special), n, bytes(ia),
                 len(ia), eng._p(filt), words, eng._p(G2), eng._st())
        R0 = 4 ** len(ia)
        assert torch.equal(G.reshape(-1)[:R0 * R0], G2.reshape(-1)[:R0 * R0])
        R = 4 ** len(ia)
        Gd = G.cpu().numpy().reshape(-1)[:R * R].reshape(R, R)
        # rows of the reduced matrix = the USED row patterns, in ascending order: compare on those
        used = [r for r in range(R) if Gd[r, r] > 0]
        assert len(used) == F.shape[0]
        np.testing.assert_array_equal(Gd[np.ix_(used, used)], F @ F.T)
        assert_score(got[s], oracle.split_score(F)) if len(ia) >= 2 else None
        if len(ia) == 1:
            assert got[s] == 0.0


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs 2 and 3): size-independent invariants
# ---------------------------------------------------------------------------------------------
def test_config2_properties(sp, eng, oracle):
    """12 taxa, 1M sites: counts sum to N; dense 6|6 Gram trace = sum of squared counts; the score of a split
    is invariant under swapping its sides and under the dense / reduced choice; true splits score lowest."""
    n, N = 12, 1_000_000
    tree, codes, tab = _count_table(sp, eng, n, N, 2)
    assert int(tab.counts.to(torch.int64).sum().item()) == N == int(tab.divisor)
    sq = int((tab.counts.to(torch.int64) ** 2).sum().item())
    scorer = eng.CountScorer(tab)
    true = set(tree.splits())
    sixes = [s for s in sp.all_splits(tree) if len(s[0]) == 6]
    t_split = [s for s in sixes if s in true][0]
    f_split = [s for s in sixes if s not in true][3]
    scores = {}
    for name, split in (("true", t_split), ("false", f_split)):
        ia, ib = eng.split_positions(split, tree.taxa)
        G, k = scorer.gram(ia, ib)
        assert k == 4096 and int(torch.diagonal(G).sum().item()) == sq
        assert torch.equal(G, G.T)
        s1 = float(scorer.score(ia, ib)[0].item())
        s2 = float(scorer.score(ib, ia)[0].item())
        s3 = float(scorer.score(ia, ib, reduced=True)[0].item())
        assert abs(s1 - s2) <= score_tol(s1) * s1 and abs(s1 - s3) <= score_tol(s1) * s1
        scores[name] = s1
    assert scores["true"] < 0.2 * scores["false"]
    # cross-check the true split against the oracle on the reduced matrix (about 1.5 s of LAPACK)
    keys, counts, usable = oracle.get_pattern_counts_arrays(codes)
    ia, ib = eng.split_positions(t_split, tree.taxa)
    ref = oracle.split_score(oracle.flattening_reduced(keys, counts / usable, n, ia, ib))
    assert_score(scores["true"], ref)



def test_int32_gram_with_correction_strip(sp, eng):
    """Large dense splits keep the Gram as int32 G0 plus the strip of the high-part correction: G0 + C equals the fp64
    Gram bit for bit, and the batched scores equal those of the fp64 route (12 taxa, 10^6 sites, 6|6 splits)."""
    n, N = 12, 1_000_000
    tree, codes, tab = _count_table(sp, eng, n, N, 2)
    scorer = eng.CountScorer(tab)
    assert scorer.n_hi > 0
    sixes = [s for s in sp.all_splits(tree) if len(s[0]) == 6][::23][:20]
    idx = [eng.split_positions(s, tree.taxa) for s in sixes]
    layout, rows_pad, pitch = scorer.geometry(4096, 4096)
    assert scorer._use_i32(layout, rows_pad, pitch)
    # (1) exact reconstruction for three splits of one batch
    s0 = scorer._buffers(layout, rows_pad, pitch, 1)[0]
    buf = scorer._buffers_i32(rows_pad, 3)
    scorer._gram_batch_i32([scorer._plan(ia, ib, False)[0] for ia, ib in idx[:3]], s0, buf, 0, layout, rows_pad, pitch)
    for b, (ia, ib) in enumerate(idx[:3]):
        m = int(buf["hm"][b].item())
        hr = buf["hr"][b, :m].to(torch.int64)
        pos = buf["pos"][b]
        assert m > 0 and torch.equal(torch.nonzero(pos >= 0).flatten(), hr)  # ascending distinct high rows
        Cfull = torch.zeros((rows_pad, rows_pad), dtype=torch.float64, device=pos.device)
        Cfull[:, hr] = buf["Cs"][b, :m].T       # columns of the high rows (symmetry)
        Cfull[hr, :] = buf["Cs"][b, :m]         # the strip rows themselves
        G, k = scorer.gram(ia, ib)
        assert k == 4096
        assert torch.equal(buf["G"][b].to(torch.float64) + Cfull, G)
    # (2) batched scores, int32 route against the fp64 route (20 splits in one batch)
    got = scorer.score_many(idx).cpu().numpy()
    scorer.int32_gram = False
    ref = scorer.score_many(idx).cpu().numpy()
    scorer.int32_gram = True
    for g, r in zip(got, ref):
        assert_score(g, r)
    one = float(scorer.score(*idx[5])[0].item())
    assert_score(got[5], one)


def test_int32_gram_without_high_counts(sp, eng):
    """A short alignment (12 taxa, 600 sites): every count stays below 256, so the correction strip is empty (m = 0)
    and the int32 route is G0 alone.  Both routes must agree."""
    tree, codes, tab = _count_table(sp, eng, 12, 600, 77)
    scorer = eng.CountScorer(tab)
    assert scorer.n_hi == 0
    idx = [(list(range(6)), list(range(6, 12))), ([0, 2, 4, 6, 8, 10], [1, 3, 5, 7, 9, 11])]
    got = scorer.score_many(idx).cpu().numpy()
    assert int(scorer._Gi[4096]["hm"][:2].sum().item()) == 0
    scorer.int32_gram = False
    ref = scorer.score_many(idx).cpu().numpy()
    for g, r in zip(got, ref):
        assert_score(g, r)


def test_config5_random_splits_32_taxa(sp, eng, oracle):
    """BASELINE configs[4] shape: random splits of a 32-taxon tree, subflattening scores (side sizes 2..16), with an
    oracle check on a sample; plus invariance of the score under swapping the two sides."""
    n, N = 32, 1_000_000  # BASELINE config 5 size
    tree = sp.trees.balanced_tree(n, 0.05)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), N, seed=5)
    aln = eng.pack(codes, want_sm=False)
    pt = eng.pair_tables_from_alignment(aln)
    rng = np.random.default_rng(5)
    splits = []
    for _ in range(3000):
        a = int(rng.integers(2, 17))
        left = sorted(rng.choice(n, size=a, replace=False).tolist())
        right = [t for t in range(n) if t not in left]
        splits.append((left, right))
    ma, mb = eng.masks_from_splits(splits)
    sc = eng.subflatten_scores(pt, ma, mb).cpu().numpy()
    sc_swapped = eng.subflatten_scores(pt, mb, ma).cpu().numpy()
    assert np.isfinite(sc).all() and (sc >= 0).all() and (sc <= 1).all()
    np.testing.assert_allclose(sc, sc_swapped, rtol=1e-9)
    c = codes.cpu().numpy().astype(np.int64)
    tables = np.zeros((n, n, 4, 4))
    for i in range(n):
        for j in range(n):
            np.add.at(tables[i, j], (c[i], c[j]), 1.0 / N)
    for s in range(0, 3000, 250):
        ref = oracle.split_score(oracle.subflattening_from_tables(tables, 1.0, splits[s][0], splits[s][1]))
        assert_score(sc[s], ref)



def test_config3_full_size_properties(sp, eng):
    """BASELINE config 3 at its full size (20 taxa, 10^7 sites): size-independent invariants of the counting side --
    linearity of the pair statistics over site ranges, every joint table summing to N, marginals consistent -- and of
    the scoring side -- all 190 two-taxon splits scored, the tree's cherries ranked first, scores independent of the
    order in which the splits are submitted."""
    n, N = 20, 10_000_000
    tree = sp.trees.balanced_tree(n, 0.05)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR((0.1, 0.2, 0.3, 0.4), (1, 2, 3, 4, 5, 6)), N, seed=3)
    aln = eng.pack(codes, want_sm=False)
    words = (N + 31) // 32
    cut = (words // 3) | 1
    whole = eng.pair_raw(aln)
    parts = eng.pair_raw(aln, 0, cut) + eng.pair_raw(aln, cut, words)
    assert torch.equal(whole, parts) and int(whole[-1].item()) == N
    pt = eng.pair_finalize(whole, n, 0.0)
    assert torch.all(pt.N.sum(dim=(2, 3)) == N)
    for i in (0, 7, 19):  # row marginals of N[i][j] do not depend on j
        m = pt.N[i].sum(dim=2)
        assert torch.equal(m, m[0].expand_as(m))
    twos = list(sp.all_splits(tree, size=2))
    idx = [eng.split_positions(s, tree.taxa) for s in twos]
    ma, mb = eng.masks_from_splits(idx)
    ptp = eng.pair_finalize(whole, n, float(N))
    sc = eng.subflatten_scores(ptp, ma, mb)
    perm = torch.randperm(len(idx), generator=torch.Generator().manual_seed(0))
    sc_perm = eng.subflatten_scores(ptp, ma[perm.numpy()], mb[perm.numpy()])
    assert torch.equal(sc[perm.cuda()], sc_perm)
    cherries = {s for s in tree.splits() if min(len(s[0]), len(s[1])) == 2}
    best = {twos[i] for i in torch.argsort(sc)[:len(cherries)].tolist()}
    assert best == cherries


def test_config3_properties(sp, eng, oracle):
    """20 taxa GTR: pair tables from the bit planes equal pair tables from the hashed pattern table; every
    table sums to the number of sites; subflattening scores of true splits are the smallest of their size."""
    n, N = 20, 2_000_000
    tree = sp.trees.balanced_tree(n, 0.05)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR((0.1, 0.2, 0.3, 0.4), (1, 2, 3, 4, 5, 6)), N, seed=3)
    aln = eng.pack(codes)
    pt = eng.pair_tables_from_alignment(aln, as_counts=True)
    assert torch.all(pt.N.sum(dim=(2, 3)) == N)
    tab = eng.count_patterns(aln)
    assert int(tab.counts.to(torch.int64).sum().item()) == N
    pt2 = eng.pair_tables_from_table(tab, as_counts=True)
    assert torch.equal(pt.N, pt2.N) and torch.equal(pt.T, pt2.T)
    true = [s for s in tree.splits() if len(s[0]) == 2 or len(s[1]) == 2]
    twos = [s for s in sp.all_splits(tree, size=2)]
    idx = [eng.split_positions(s, tree.taxa) for s in twos]
    ma, mb = eng.masks_from_splits(idx)
    sc = eng.subflatten_scores(eng.pair_tables_from_alignment(aln), ma, mb).cpu().numpy()
    best = {twos[i] for i in np.argsort(sc)[:len(true)]}
    assert best == set((tuple(a), tuple(b)) for a, b in true)
    # oracle spot check on the hashed table
    keys = tab.keys.cpu().numpy().view(np.uint64)
    vals = tab.counts.cpu().numpy().astype(np.float64) / N
    for i in (0, 17, 101):
        ref = oracle.split_score(oracle.subflattening(keys, vals, n, idx[i][0], idx[i][1]))
        assert_score(sc[i], ref)


def test_rank1_golden(sp, eng, oracle, golden_rank1):
    """Banned-pattern sparse flattenings (bit-exact), rank-1 vectors / approximation / divergence and the rank-k
    approximation against outputs of the unmodified reference (constructions.py:94-105, phylogenetics.py:331-373)."""
    for rec in golden_rank1["cases"]:
        n = rec["n"]
        aln = dict(zip(rec["patterns"], rec["values"]))
        taxa = [str(i) for i in range(n)]
        split = rec["split"]
        F = sp.flattening(split, aln, sp.FlatFormat.reduced)
        assert sp.phylogenetics.flattening_rank_1_approximation_divergence(F) == pytest.approx(rec["divergence"], rel=1e-12)
        approx, r, c = sp.phylogenetics.flattening_rank_1_approximation(F, return_vectors=True)
        np.testing.assert_allclose(r, rec["r"], rtol=1e-13)
        np.testing.assert_allclose(c, rec["c"], rtol=1e-13)
        assert list(approx.shape) == rec["approx_shape"]
        np.testing.assert_allclose(approx, oracle.rank_1_approximation(F), rtol=1e-13)
        none, r2, c2 = sp.phylogenetics.flattening_rank_1_approximation(F, return_vectors=True, dont_compute_matrix=True)
        assert none is None and r2 == r and c2 == c
        # table route (no materialised flattening), both value kinds
        a, b = ([int(ch) for ch in side] for side in split.split("|"))
        table = eng.table_from_mapping(aln)
        assert float(eng.rank1_divergence_table(table, a, b).item()) == pytest.approx(rec["divergence"], rel=1e-12)
        plain = eng.PatternTable(n, table.keys, values=table.values_f64())
        assert float(eng.rank1_divergence_table(plain, a, b).item()) == pytest.approx(rec["divergence"], rel=1e-12)
        for name, (rr, cc, vv) in rec["banned"].items():
            kw = {"ban_row_patterns": name[0]} if name[1:] == "row" else {"ban_col_patterns": name[0]}
            M = sp.constructions.sparse_flattening_with_banned_patterns(split, aln, taxa, **kw).tocoo()
            order = np.lexsort((M.col, M.row))
            assert M.row[order].tolist() == rr and M.col[order].tolist() == cc and M.data[order].tolist() == vv
        K = sp.phylogenetics.flattening_rank_k_approximation(split.split("|"), aln)
        assert list(K.shape) == rec["rank_k_shape"] and K.sum() == pytest.approx(rec["rank_k_sum"], rel=1e-12)
        keys, _ = oracle.patterns_to_keys(rec["patterns"])
        np.testing.assert_allclose(K.toarray(), oracle.rank_k_approximation(keys, np.array(rec["values"]), n, a, b),
                                   rtol=1e-12, atol=1e-18)


def test_erickson_mutual_information_golden(sp, golden_rank1):
    """erickson_SVD(method=mutual_information) picks the same splits, in the same order, as the reference
    (phylogenetics.py:136-140); the golden divergences of all splits differ by > 6e-6, far above the tolerance."""
    for rec in golden_rank1["erickson"]:
        aln = dict(zip(rec["patterns"], rec["values"]))
        got = [list(map(list, s)) for s in sp.erickson_SVD(aln, method=sp.Method.mutual_information)]
        assert got == rec["mutual_information"]
        taxa = [str(i) for i in range(rec["n"])]
        for name, ref in rec["divergences"].items():
            F = sp.flattening(name, sp.Alignment(aln, taxa), sp.FlatFormat.reduced)
            assert sp.phylogenetics.flattening_rank_1_approximation_divergence(F) == pytest.approx(ref, rel=1e-12)


def test_rank1_divergence_routes_agree(sp, eng, oracle):
    """Direct-indexed, hashed and materialised routes give the same divergence at sizes the oracle loop cannot reach:
    17 taxa, splits with a 2-, 8- and 14-taxon side (4^14 row indices -> hash table), and a split that leaves a taxon
    out (dense route)."""
    n = 17
    codes = random_codes(n, 60_000, 17, 0.0)
    table = eng.count_patterns(eng.pack(codes))
    keys = table.keys.cpu().numpy().astype(np.uint64)
    cnt = table.counts.cpu().numpy().astype(np.float64)
    total = cnt.sum()
    for a in ([0, 5], list(range(8)), list(range(3, 17)), [2, 4, 6, 8, 10, 12, 14, 16, 1, 3, 5, 7]):
        b = [t for t in range(n) if t not in a]
        rows = oracle.side_index(keys, n, a)
        cols = oracle.side_index(keys, n, b)
        ur, ri = np.unique(rows, return_inverse=True)
        uc, ci = np.unique(cols, return_inverse=True)
        rs = np.bincount(ri, weights=cnt) / total
        cs = np.bincount(ci, weights=cnt) / total
        v = cnt / total
        ref = float(np.sum(v * np.log(v / (rs[ri] * cs[ci]))))
        got = float(eng.rank1_divergence(table, a, b).item())
        assert got == pytest.approx(ref, rel=1e-12), (len(a), got, ref)
        if len(a) <= 8:
            dense = float(eng.rank1_divergence_dense(eng.flatten_reduced(table, a, b)).item())
            assert dense == pytest.approx(ref, rel=1e-12)
    part = float(eng.rank1_divergence(table, [0, 1, 2], list(range(4, 12))).item())  # taxa 3, 12..16 left out
    F = eng.flatten_reduced(table, [0, 1, 2], list(range(4, 12))).cpu().numpy()
    assert part == pytest.approx(oracle.rank_1_divergence(F), rel=1e-12)


def test_config4_full_size_properties(sp, eng):
    """BASELINE configs[3] at its full size (64 taxa, 10^8 sites, 128-bit keys): every usable site is counted once; the
    Gram of a thin split has trace = sum of squared counts (every pattern is one cell of the flattening) whatever the
    split; the column filter changes nothing; a true cherry scores below a false pair."""
    n, N = 64, 100_000_000
    tree = sp.trees.balanced_tree(n, 0.05)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), N, seed=4)
    wide, valid, n_, N_ = eng.pack_wide(codes)
    del codes
    tab = eng.count_patterns_wide(wide, valid, n_, N_)
    del wide
    _, counts = tab.compact(sort=False)
    c64 = counts.to(torch.int64)
    assert int(c64.sum().item()) == N == int(tab.divisor)
    sq = int((c64 * c64).sum().item())
    del counts, c64
    words = int(eng.lib.spb_thin_filter_words(tab.cap))
    filt = torch.empty(2 * words, dtype=torch.int32, device="cuda")
    for side in ([0, 1], [31, 32], [5]):
        R = 4 ** len(side)
        G1 = torch.empty((16, 16), dtype=torch.float64, device="cuda")
        G2 = torch.empty((16, 16), dtype=torch.float64, device="cuda")
        eng.call("spb_thin_gram_wide", eng._p(tab.hkeys), eng._p(tab.hcounts), tab.cap, eng._p(tab.special), n, bytes(side), len(side),
                 eng._p(G1), eng._st())
        eng.call("spb_thin_gram_wide_filtered", eng._p(tab.hkeys), eng._p(tab.hcounts), tab.cap, eng._p(tab.special), n, bytes(side),
                 len(side), eng._p(filt), words, eng._p(G2), eng._st())
        g1 = G1.reshape(-1)[:R * R].reshape(R, R)
        assert torch.equal(g1, G2.reshape(-1)[:R * R].reshape(R, R)) and torch.equal(g1, g1.T)
        assert int(torch.diagonal(g1).sum().item()) == sq
    cherry, false_pair = eng.thin_split_scores(tab, [[0, 1], [0, 2]]).cpu().numpy()  # positions 0, 1 = taxa t0, t1: a cherry
    assert 0.0 <= cherry < 0.5 * false_pair
