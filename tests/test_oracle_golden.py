"""Pins oracle/splitp_oracle.py against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def _table(oracle, probs):
    pats = list(probs.keys())
    keys, n = oracle.patterns_to_keys(pats)
    return keys, np.array([probs[p] for p in pats]), n


def test_four_pattern_golden(oracle, golden_small):
    g = golden_small["four_pattern"]
    keys, vals, n = _table(oracle, g["probs"])
    for k, (a, b) in enumerate(g["splits"]):
        # the reference tests pass *sets* of ints; iteration order of small-int sets is ascending
        D = oracle.flattening_dense(keys, vals, n, a, b)
        np.testing.assert_array_equal(D, np.array(g["sparse"][k]))
        R = oracle.flattening_reduced(keys, vals, n, a, b)
        np.testing.assert_array_equal(R, np.array(g["reduced"][k]))
        S = oracle.subflattening(keys, vals, n, a, b)
        np.testing.assert_allclose(S, np.array(g["subflat"][k]), rtol=1e-15, atol=1e-15)
        Sd = oracle.subflattening_direct(keys, vals, n, a, b)
        np.testing.assert_allclose(Sd, np.array(g["subflat"][k]), rtol=1e-15, atol=1e-15)
        assert oracle.split_score(R) == pytest.approx(g["scores_reduced"][k], rel=1e-12, abs=1e-12)
        assert oracle.split_score(S) == pytest.approx(g["scores_subflat"][k], rel=1e-12, abs=1e-12)


def test_subflattening_identity(oracle, golden_small):
    """tests/test_constructions.py:90-107: subflattening == (S^{x4} F S^{x4}^T)[idx, idx]."""
    g = golden_small["four_pattern"]
    keys, vals, n = _table(oracle, g["probs"])
    S = np.array([[1, -1], [1, 1]])
    S0 = np.kron(S, S)
    S4 = np.kron(S0, S0)
    idx = [3, 7, 11, 12, 13, 14, 15]
    for a, b in g["splits"]:
        F = oracle.flattening_dense(keys, vals, n, a, b)
        ref = (S4 @ F @ S4.T)[np.ix_(idx, idx)]
        np.testing.assert_allclose(oracle.subflattening(keys, vals, n, a, b), ref, rtol=1e-15, atol=1e-15)


def test_fasta_small(oracle, golden_small):
    g = golden_small["fasta_small"]
    counts, usable = oracle.get_pattern_counts(list(g["sequences"].values()))
    assert usable == g["usable"]
    assert oracle.pattern_counts_to_probs(counts, usable) == g["probs"]
    assert list(oracle.pattern_counts_to_probs(counts, usable).keys()) == list(g["probs"].keys())


def test_fasta_dirty(oracle, golden_small):
    g = golden_small["fasta_dirty"]
    counts, usable = oracle.get_pattern_counts(g["rows"])
    assert usable == g["usable"]
    assert list(counts.items()) == [tuple(x) for x in g["counts_items"]]  # same keys, counts AND dict order
    probs = oracle.pattern_counts_to_probs(counts, usable)
    assert list(probs.items()) == [tuple(x) for x in g["probs_items"]]  # bit-exact divisions


def test_all_splits(oracle, golden_small):
    g = golden_small["all_splits"]
    for taxa, d in g.items():
        if taxa == "named5":
            got = [[list(a), list(b)] for a, b in oracle.all_splits([f"t{i}" for i in range(5)])]
            assert got == d
            continue
        assert [[list(a), list(b)] for a, b in oracle.all_splits(taxa)] == d["default"]
        assert [[list(a), list(b)] for a, b in oracle.all_splits(taxa, trivial=True)] == d["trivial"]
        assert [[list(a), list(b)] for a, b in oracle.all_splits(taxa, size=2)] == d["size2"]
        assert list(oracle.all_splits(taxa, string_format=True)) == d["strings"]


def test_sub_alignment(oracle, golden_small):
    g = golden_small["sub_alignment"]
    keys, n = oracle.patterns_to_keys(g["patterns"])
    vals = np.array(g["values"])
    for case in g["cases"]:
        idx = [g["taxa"].index(c) for c in case["sub"]]
        k, v = oracle.sub_alignment(keys, vals, n, idx)
        pats = oracle.keys_to_patterns(k, len(idx))
        ref = dict((p, x) for p, x in case["items"])
        assert sorted(pats) == sorted(ref)
        for p, x in zip(pats, v):
            assert x == pytest.approx(ref[p], rel=1e-13)


def test_partial_cover_quirk(oracle, golden_small):
    q = golden_small["quirks"]
    keys, n = oracle.patterns_to_keys(["AACG", "AATG", "CATG"])
    vals = np.array([0.25, 0.5, 0.25])
    D = oracle.flattening_dense(keys, vals, n, [0], [1, 3])
    np.testing.assert_array_equal(D, np.array(q["partial_cover_sparse"]))
    R = oracle.flattening_reduced(keys, vals, n, [0], [1, 3])
    np.testing.assert_array_equal(R, np.array(q["partial_cover_reduced"]))
    assert q["unknown_format_returns_none"] is True
    assert q["subflat_string_plain_dict"] == "KeyError"
    assert q["score_3x5"] == 0.0 and q["score_4x4"] == 0.0


def test_random_tables(oracle, golden_random):
    z, meta = golden_random
    cache = {}
    for rec in meta:
        c = rec["case"]
        if c not in cache:
            pats = [str(p) for p in z[f"c{c}_patterns"]]
            keys, n = oracle.patterns_to_keys(pats)
            cache[c] = (keys, z[f"c{c}_values"], n)
        keys, vals, n = cache[c]
        a, b = rec["split"].split("|")
        ia, ib = [int(x) for x in a], [int(x) for x in b]
        D = oracle.flattening_dense(keys, vals, n, ia, ib)
        assert list(D.shape) == rec["dense_shape"] and h(D) == rec["dense_hash"]
        R = oracle.flattening_reduced(keys, vals, n, ia, ib)
        assert list(R.shape) == rec["reduced_shape"] and h(R) == rec["reduced_hash"]
        if "reduced_key" in rec:
            np.testing.assert_array_equal(R, z[rec["reduced_key"]])
        assert oracle.split_score(R) == pytest.approx(rec["score_reduced"], rel=1e-11, abs=1e-13)
        if "subflat_key" in rec:
            S = oracle.subflattening(keys, vals, n, ia, ib)
            np.testing.assert_allclose(S, z[rec["subflat_key"]], rtol=1e-12, atol=1e-14)
            assert oracle.split_score(S) == pytest.approx(rec["score_subflat"], rel=1e-9, abs=1e-12)
            # the Gram/eigen cross-check formula agrees with the reference formula
            assert oracle.split_score_gram(S) == pytest.approx(rec["score_subflat"], rel=1e-8, abs=1e-10)


def test_readme_config(oracle, golden_readme):
    """Config C1 (README.md:18-42, random.seed(0)): reduced flattenings bit-exact, scores to 1e-12."""
    g = golden_readme
    pats = [str(p) for p in g["patterns"]]
    keys, n = oracle.patterns_to_keys(pats)
    vals = g["values"]
    taxa = [str(t) for t in g["taxa"]]
    assert n == 10 and len(pats) == 1910
    # generate_alignment returns keys sorted lexicographically A<C<G<T (simulation.py:50-54)
    assert np.all(np.diff(keys.astype(np.int64)) > 0)
    got_splits = list(oracle.all_splits(taxa, string_format=True))[:100]
    assert got_splits == [str(s) for s in g["splits"]]
    for k, s in enumerate(got_splits):
        a, b = s.split("|")
        ia, ib = [taxa.index(x) for x in a], [taxa.index(x) for x in b]
        R = oracle.flattening_reduced(keys, vals, n, ia, ib)
        assert tuple(R.shape) == tuple(g["shapes"][k])
        assert h(R) == str(g["hashes"][k])
        if f"full_{k}" in g:
            np.testing.assert_array_equal(R, g[f"full_{k}"])
        assert oracle.split_score(R) == pytest.approx(float(g["scores"][k]), rel=1e-12)
    for k, s in enumerate(got_splits[:6]):
        a, b = s.split("|")
        ia, ib = [taxa.index(x) for x in a], [taxa.index(x) for x in b]
        S = oracle.subflattening(keys, vals, n, ia, ib)
        assert oracle.split_score(S) == pytest.approx(float(g["sub_scores"][k]), rel=1e-9)


def test_pair_tables_from_codes_matches_pattern_route(oracle):
    """The code-matrix route to the pair tables (used above 31 taxa) equals the pattern-table route."""
    rng = np.random.default_rng(8)
    codes = rng.integers(0, 4, size=(7, 5000)).astype(np.uint8)
    codes[rng.random(codes.shape) < 0.01] = 255
    keys, counts, usable = oracle.get_pattern_counts_arrays(codes)
    ref = oracle.pair_tables(keys, counts, 7)
    got, us = oracle.pair_tables_from_codes(codes, as_counts=True)
    assert us == usable
    np.testing.assert_array_equal(got, ref)
    S1 = oracle.subflattening(keys, counts / usable, 7, [0, 3, 5], [1, 2, 4, 6])
    S2 = oracle.subflattening_from_tables(oracle.pair_tables_from_codes(codes)[0], 1.0, [0, 3, 5], [1, 2, 4, 6])
    np.testing.assert_allclose(S1, S2, rtol=1e-12, atol=1e-14)


def test_wide_oracle_routes_agree(oracle):
    """The any-number-of-taxa routes (byte-row counting, dict-based reduced flattening) equal the uint64-key routes."""
    rng = np.random.default_rng(9)
    base = rng.integers(0, 4, size=(9, 40))
    codes = base[:, rng.integers(0, 40, size=3000)].astype(np.uint8)
    codes[rng.random(codes.shape) < 0.03] = rng.integers(0, 4)
    codes[rng.random(codes.shape) < 0.002] = 255
    d, usable = oracle.get_pattern_counts_wide(codes)
    keys, counts, us = oracle.get_pattern_counts_arrays(codes)
    assert us == usable and d == dict(zip(oracle.keys_to_patterns(keys, 9), counts.tolist()))
    assert list(d) == oracle.keys_to_patterns(keys, 9)  # first-occurrence order
    for ia in ([0, 1], [7], [3, 8]):
        ib = [t for t in range(9) if t not in ia]
        np.testing.assert_array_equal(oracle.flattening_reduced_from_dict(d, ia, ib),
                                      oracle.flattening_reduced(keys, counts.astype(float), 9, ia, ib))


def test_rank1_golden(oracle, golden_rank1):
    """Banned-pattern flattenings, rank-1 vectors / approximation / divergence and the rank-k approximation against
    the reference's outputs (constructions.py:94-101, phylogenetics.py:331-373)."""
    for rec in golden_rank1["cases"]:
        n = rec["n"]
        keys, _ = oracle.patterns_to_keys(rec["patterns"])
        vals = np.array(rec["values"])
        a, b = ([int(ch) for ch in side] for side in rec["split"].split("|"))
        F = oracle.flattening_reduced(keys, vals, n, a, b)
        r, c = oracle.rank_1_vectors(F)
        assert r.tolist() == rec["r"] and c.tolist() == rec["c"]
        approx = oracle.rank_1_approximation(F)
        assert list(approx.shape) == rec["approx_shape"] and h(approx) == rec["approx_sha"]
        assert oracle.rank_1_divergence(F) == pytest.approx(rec["divergence"], rel=1e-14)
        for name, (rr, cc, vv) in rec["banned"].items():
            code = "ACGT".index(name[0])
            kw = {"ban_row": code} if name[1:] == "row" else {"ban_col": code}
            rows, cols, v = oracle.flattening_coo_banned(keys, vals, n, a, b, **kw)
            order = np.lexsort((cols, rows))
            assert rows[order].tolist() == rr and cols[order].tolist() == cc and v[order].tolist() == vv
        K = oracle.rank_k_approximation(keys, vals, n, a, b)
        assert list(K.shape) == rec["rank_k_shape"]
        assert K.sum() == pytest.approx(rec["rank_k_sum"], rel=1e-12)
        if "rank_k" in rec:
            ref = np.zeros(K.shape)
            ref[rec["rank_k"][0], rec["rank_k"][1]] = rec["rank_k"][2]
            np.testing.assert_allclose(K, ref, rtol=1e-12, atol=1e-18)
