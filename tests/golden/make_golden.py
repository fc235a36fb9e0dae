"""Generate golden fixtures by running the UNMODIFIED reference (js51/SplitP v0.3.2).

Run in the build container only (the reference is not present on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_golden.py

Writes tests/golden/golden_small.json, golden_readme.npz, golden_random.npz, golden_erickson.json, golden_rank1.json.  The fixtures pin
oracle/splitp_oracle.py and (through it, or directly) the CUDA path.  Nothing here is imported by
the product.
"""
import hashlib
import json
import os
import random
import sys
import tempfile

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import splitp  # noqa: E402
from splitp.parsers import fasta  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


class _T:  # minimal tree stand-in for all_splits (needs .taxa / .get_taxa())
    def __init__(self, taxa):
        self.taxa = list(taxa)

    def get_taxa(self):
        return self.taxa


def small():
    """The reference's own golden vectors, re-derived by calling the reference."""
    out = {}
    probs = {"ATCG": 2 / 5, "GATC": 1 / 5, "CGAT": 1 / 5, "TCGA": 1 / 5}  # tests/test_constructions.py:5-10
    splits = [([0, 1], [2, 3]), ([0, 2], [1, 3]), ([0, 3], [1, 2])]
    out["four_pattern"] = {"probs": probs, "splits": splits, "sparse": [], "reduced": [], "subflat": [], "scores_reduced": [],
                           "scores_subflat": []}
    for a, b in splits:
        sp = (set(a), set(b))  # the reference tests pass sets of ints (test_constructions.py:12)
        out["four_pattern"]["sparse"].append(np.asarray(splitp.flattening(sp, probs).todense()).tolist())
        red = splitp.flattening(sp, probs, splitp.FlatFormat.reduced)
        out["four_pattern"]["reduced"].append(red.tolist())
        sub = splitp.subflattening(sp, probs)
        out["four_pattern"]["subflat"].append(sub.tolist())
        out["four_pattern"]["scores_reduced"].append(float(splitp.split_score(red)))
        out["four_pattern"]["scores_subflat"].append(float(splitp.split_score(sub)))

    # FASTA golden (tests/test_parsers.py:16-26): same 4x5 alignment, written by us.
    seqs = {" TAXON_1": "AAGCT", " TAXON_2": "TTAGC", " TAXON_3": "CCTAG", " TAXON_4": "GGCTA"}
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        for k, v in seqs.items():
            f.write(f">{k}\n{v}\n")
        path = f.name
    pp, L = fasta.pattern_probs_from_alignment(path, return_sequence_length=True)
    out["fasta_small"] = {"sequences": seqs, "probs": pp, "usable": L}
    os.unlink(path)

    # counting with invalid characters, lower case, and wrapped lines
    rng = random.Random(11)
    n, N = 7, 400
    rows = []
    for _ in range(n):
        s = "".join(rng.choice("ACGTacgtN-RY?" if rng.random() < 0.08 else "ACGT") for _ in range(N))
        rows.append(s)
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        for i, s in enumerate(rows):
            f.write(f">t{i}\n")
            for k in range(0, N, 61):
                f.write(s[k:k + 61] + "\n")
        path = f.name
    aln = fasta.read_alignment_from_file(path)
    counts, L = fasta.get_pattern_counts(aln)
    out["fasta_dirty"] = {"rows": rows, "names": list(aln.keys()), "counts_items": list(counts.items()), "usable": L,
                          "probs_items": list(fasta.pattern_counts_to_probs(counts, L).items())}
    os.unlink(path)

    # all_splits (splits.py:27-59)
    out["all_splits"] = {}
    for taxa in ["0123", "012345", "0123456", "01234567", "0123456789"]:
        t = _T(taxa)
        out["all_splits"][taxa] = {
            "default": [[list(a), list(b)] for a, b in splitp.all_splits(t)],
            "trivial": [[list(a), list(b)] for a, b in splitp.all_splits(t, trivial=True)],
            "size2": [[list(a), list(b)] for a, b in splitp.all_splits(t, size=2)],
            "strings": list(splitp.all_splits(t, string_format=True)),
        }
    t = _T([f"t{i}" for i in range(5)])
    out["all_splits"]["named5"] = [[list(a), list(b)] for a, b in splitp.all_splits(t)]

    # Alignment.sub_alignment (alignment.py:10-31)
    rng = np.random.default_rng(3)
    pats = sorted({"".join(rng.choice(list("ACGT"), 5)) for _ in range(60)})
    vals = rng.integers(1, 50, len(pats)).astype(float)
    vals /= vals.sum()
    al = splitp.alignment.Alignment(dict(zip(pats, vals.tolist())), "abcde")
    out["sub_alignment"] = {"patterns": pats, "values": vals.tolist(), "taxa": "abcde", "cases": []}
    for sub in ["ace", "eb", "abcd"]:
        sa = al.sub_alignment(tuple(sub))
        out["sub_alignment"]["cases"].append({"sub": sub, "taxa": list(sa.taxa), "items": sorted(sa.data.items())})

    # quirks
    q = {}
    q["unknown_format_returns_none"] = splitp.flattening("01|23", {"ACGT": 1.0}, "nope") is None
    try:
        splitp.subflattening("01|23", {"ACGT": 1.0})
        q["subflat_string_plain_dict"] = "ok"
    except Exception as e:  # KeyError(4) per the survey
        q["subflat_string_plain_dict"] = type(e).__name__
    q["score_3x5"] = float(splitp.split_score(np.arange(15.0).reshape(3, 5) + 1))
    q["score_4x4"] = float(splitp.split_score(np.arange(16.0).reshape(4, 4) ** 1.5 + 1))
    # overwrite (not accumulate) when the split does not cover all taxa (constructions.py:43,101)
    tab = {"AACG": 0.25, "AATG": 0.5, "CATG": 0.25}
    al3 = splitp.alignment.Alignment(tab, "0123")
    q["partial_cover_sparse"] = np.asarray(splitp.flattening((["0"], ["1", "3"]), al3).todense()).tolist()
    q["partial_cover_reduced"] = splitp.flattening((["0"], ["1", "3"]), al3, splitp.FlatFormat.reduced).tolist()
    out["quirks"] = q
    with open(os.path.join(HERE, "golden_small.json"), "w") as f:
        json.dump(out, f, indent=0)


def readme():
    """Config C1 = README.md:18-42 with random.seed(0): 10 taxa, bl 0.05, 10,000 bp, first 100 splits."""
    random.seed(0)
    np.random.seed(0)
    model = splitp.model.GTR.JukesCantor(1 / 2)
    tree = splitp.trees.balanced_newick_tree(10, 0.05)
    splits = splitp.all_splits(tree)
    aln = splitp.generate_alignment(tree, model, 10_000)
    pats = list(aln.keys())
    vals = np.array(list(aln.values()))
    sp_list, scores, shapes, hashes, fulls = [], [], [], [], {}
    for s in range(100):
        split = next(splits)
        F = splitp.flattening(split, aln, splitp.FlatFormat.reduced)
        scores.append(float(splitp.split_score(F)))
        shapes.append(F.shape)
        hashes.append(h(F))
        sp_list.append("".join(split[0]) + "|" + "".join(split[1]))
        if s in (0, 50, 99):
            fulls[f"full_{s}"] = F
    # sparse path scores for a few splits (ARPACK; nondeterministic at ~5e-12)
    true_splits = sorted("".join(a) + "|" + "".join(b) for a, b in tree.splits())
    sub_scores = []
    data = {}
    alA = splitp.alignment.Alignment(dict(aln), tree.taxa)
    for s in sp_list[:6]:
        a, b = s.split("|")
        sub_scores.append(float(splitp.split_score(splitp.subflattening((tuple(a), tuple(b)), alA, data))))
    np.savez_compressed(os.path.join(HERE, "golden_readme.npz"), patterns=np.array(pats), values=vals,
                        splits=np.array(sp_list), scores=np.array(scores), shapes=np.array(shapes),
                        hashes=np.array(hashes), true_splits=np.array(true_splits), taxa=np.array(tree.taxa),
                        sub_scores=np.array(sub_scores), **fulls)


def rand():
    """Random pattern tables, every split of 4..8 taxa: sparse / reduced / subflattening / scores."""
    rng = np.random.default_rng(2024)
    store = {}
    meta = []
    for case, (n, P) in enumerate([(4, 30), (5, 80), (6, 250), (7, 300), (8, 500)]):
        taxa = "0123456789"[:n]
        pats = sorted({"".join(rng.choice(list("ACGT"), n, p=[.4, .3, .2, .1])) for _ in range(P)})
        rng.shuffle(pats)
        counts = rng.integers(0 if case == 1 else 1, 200, len(pats))  # case 1 has explicit zeros
        vals = counts / counts.sum()
        aln = splitp.alignment.Alignment(dict(zip(pats, vals.tolist())), taxa)
        store[f"c{case}_patterns"] = np.array(pats)
        store[f"c{case}_counts"] = counts
        store[f"c{case}_values"] = vals
        splits = list(splitp.all_splits(_T(taxa), string_format=True))
        if n >= 7:
            splits = splits[::3]
        data = {}
        for si, s in enumerate(splits):
            a, b = s.split("|")
            sp = (tuple(a), tuple(b))
            D = np.asarray(splitp.flattening(sp, aln).todense())
            R = splitp.flattening(sp, aln, splitp.FlatFormat.reduced)
            rec = {"case": case, "split": s, "dense_hash": h(D), "dense_shape": list(D.shape), "reduced_hash": h(R),
                   "reduced_shape": list(R.shape), "score_reduced": float(splitp.split_score(R))}
            if n <= 6:
                S = splitp.subflattening(sp, aln, data)
                store[f"c{case}_s{si}_subflat"] = S
                rec["score_subflat"] = float(splitp.split_score(S))
                rec["subflat_key"] = f"c{case}_s{si}_subflat"
            if si < 3:
                store[f"c{case}_s{si}_reduced"] = R
                rec["reduced_key"] = f"c{case}_s{si}_reduced"
            meta.append(rec)
    store["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "golden_random.npz"), **store)


def erickson():
    """erickson_SVD (phylogenetics.py:99-171) on simulated 6- and 8-taxon alignments, both scoring methods."""
    out = []
    for n, N, seed in ((6, 3000, 1), (8, 4000, 2)):
        random.seed(seed)
        np.random.seed(seed)
        tree = splitp.trees.balanced_newick_tree(n, 0.1)
        aln = splitp.generate_alignment(tree, splitp.model.GTR.JukesCantor(1 / 2), N)
        rec = {"n": n, "patterns": list(aln.keys()), "values": list(aln.values()),
               "true_splits": sorted([list(map(list, s)) for s in tree.splits()])}
        for method in (splitp.Method.flattening, splitp.Method.subflattening):
            res = splitp.phylogenetics.erickson_SVD(aln, method=method)
            rec[method.name] = [list(map(list, s)) for s in res]
        out.append(rec)
    with open(os.path.join(HERE, "golden_erickson.json"), "w") as f:
        json.dump(out, f)


def erickson_trace():
    """Per-step candidate scores of the reference's erickson_SVD, taken from its own `show_work` output
    (phylogenetics.py:143-144 prints the {pair: (pair, split, score)} table of every agglomeration step): for every step
    the chosen split, its score and the runner-up's score, so that a test can pin the ORDER of the picks wherever the
    reference's margin is far above rounding noise."""
    import contextlib
    import io
    with open(os.path.join(HERE, "golden_erickson.json")) as f:
        cases = json.load(f)
    out = []
    for rec in cases:
        aln = dict(zip(rec["patterns"], rec["values"]))
        tr = {"n": rec["n"]}
        for method in (splitp.Method.flattening, splitp.Method.subflattening, splitp.Method.mutual_information):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                res = splitp.phylogenetics.erickson_SVD(aln, method=method, show_work=True)
            steps = []
            for line, chosen in zip([l for l in buf.getvalue().splitlines() if l.startswith("Scores: ")], res):
                table = eval(line[len("Scores: "):], {"np": np, "nan": float("nan"), "inf": float("inf")})
                vals = sorted(float(v[2]) for v in table.values() if not np.isnan(float(v[2])))
                steps.append({"chosen": [list(s) for s in chosen], "best": vals[0], "second": vals[1] if len(vals) > 1 else None,
                              "candidates": len(table), "nan": int(sum(np.isnan(float(v[2])) for v in table.values()))})
            assert len(steps) == len(res)
            tr[method.name] = steps
        out.append(tr)
    with open(os.path.join(HERE, "golden_erickson_trace.json"), "w") as f:
        json.dump(out, f)


def rank1():
    """Banned-pattern sparse flattenings (constructions.py:58-105), rank-1 / rank-k approximations and the rank-1
    divergence (phylogenetics.py:331-373), erickson_SVD with Method.mutual_information (phylogenetics.py:136-140)."""
    out = {"cases": [], "erickson": []}
    rng = np.random.default_rng(11)
    for n, P, splits in ((4, 40, ["01|23", "02|13", "0|123"]), (5, 120, ["01|234", "024|13", "3|0124"]),
                         (6, 300, ["012|345", "05|1234", "135|024"])):
        pats = set()
        while len(pats) < P:
            pats.add("".join("ACGT"[i] for i in rng.integers(0, 4, n)))
        pats = sorted(pats)
        rng.shuffle(pats)
        counts = rng.integers(1, 50, len(pats))
        total = int(counts.sum())
        aln = {p: int(c) / float(total) for p, c in zip(pats, counts)}
        taxa = [str(i) for i in range(n)]
        for split in splits:
            rec = {"n": n, "split": split, "patterns": pats, "values": [aln[p] for p in pats]}
            F = splitp.flattening(split, aln, splitp.FlatFormat.reduced)
            rec["divergence"] = float(splitp.phylogenetics.flattening_rank_1_approximation_divergence(F))
            approx, r, c = splitp.phylogenetics.flattening_rank_1_approximation(F, return_vectors=True)
            rec["r"], rec["c"], rec["approx_sha"] = r, c, h(approx)
            rec["approx_shape"] = list(approx.shape)
            banned = {}
            for char in ("ACGT" if n == 4 else "AG"):
                for side in ("row", "col"):
                    kw = {"ban_row_patterns": char} if side == "row" else {"ban_col_patterns": char}
                    M = splitp.constructions.sparse_flattening_with_banned_patterns(split, aln, taxa, **kw).tocoo()
                    order = np.lexsort((M.col, M.row))
                    banned[char + side] = [M.row[order].tolist(), M.col[order].tolist(), M.data[order].tolist()]
            rec["banned"] = banned
            K = splitp.phylogenetics.flattening_rank_k_approximation(split.split("|"), aln)
            K = K.tocoo()
            order = np.lexsort((K.col, K.row))
            rec["rank_k_shape"] = list(K.shape)
            rec["rank_k_sha"] = h(K.toarray())
            rec["rank_k_sum"] = float(K.toarray().sum())
            if n == 4:
                rec["rank_k"] = [K.row[order].tolist(), K.col[order].tolist(), K.data[order].tolist()]
            out["cases"].append(rec)
    for n, N, seed in ((6, 3000, 1), (8, 2000, 3)):
        random.seed(seed)
        np.random.seed(seed)
        tree = splitp.trees.balanced_newick_tree(n, 0.1)
        aln = splitp.generate_alignment(tree, splitp.model.GTR.JukesCantor(1 / 2), N)
        res = splitp.phylogenetics.erickson_SVD(aln, method=splitp.Method.mutual_information)
        # the divergence of every split, so that the test can tell near-ties from real disagreements
        div = {}
        for s in splitp.all_splits(_T([str(i) for i in range(n)]), trivial=False):
            F = splitp.flattening(s, aln, splitp.FlatFormat.reduced)
            div["|".join("".join(x) for x in s)] = float(splitp.phylogenetics.flattening_rank_1_approximation_divergence(F))
        out["erickson"].append({"n": n, "patterns": list(aln.keys()), "values": list(aln.values()),
                                "mutual_information": [list(map(list, s)) for s in res], "divergences": div,
                                "true_splits": sorted([list(map(list, s)) for s in tree.splits()])})
    with open(os.path.join(HERE, "golden_rank1.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "erickson_trace":
        erickson_trace()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "erickson":
        erickson()
    elif len(sys.argv) > 1 and sys.argv[1] == "rank1":
        rank1()
    else:
        small()
        rand()
        readme()
        erickson()
        rank1()
    print("golden fixtures written")
