"""Diagnostic: erickson_SVD(method=mutual_information) golden case, repeated, with per-split divergences against the oracle."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import splitp_b200 as sp
from oracle import splitp_oracle as O
eng = sp.engine
g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_rank1.json")))
for ri, rec in enumerate(g["erickson"]):
    aln = dict(zip(rec["patterns"], rec["values"]))
    n = len(rec["patterns"][0])
    keys, _ = O.patterns_to_keys(rec["patterns"])
    vals = np.asarray(rec["values"], dtype=np.float64)
    table = eng.table_from_mapping(aln)
    import itertools
    worst = 0.0
    for a in range(1, n // 2 + 1):
        for left in itertools.combinations(range(n), a):
            right = [t for t in range(n) if t not in left]
            for ia, ib in ((list(left), right), (right, list(left))):
                ref = O.rank_1_divergence(O.flattening_dense(keys, vals, n, ia, ib))
                got = [float(eng.rank1_divergence(table, ia, ib).item()) for _ in range(3)]
                err = max(abs(x - ref) / abs(ref) for x in got)
                worst = max(worst, err)
                if err > 1e-10:
                    print("MISMATCH record", ri, "split", ia, ib, "ref", ref, "got", got)
    print("record", ri, "n", n, "worst rel err of rank1_divergence over all splits:", worst)
    for rep in range(8):
        tr = []
        got = [list(map(list, s)) for s in sp.erickson_SVD(aln, method=sp.Method.mutual_information, trace=tr)]
        ok = got == rec["mutual_information"]
        print("  rep", rep, "matches golden:", ok)
        if not ok:
            for step, t in enumerate(tr):
                print("    step", step, "chosen", t["chosen"])
                for split, v in t["scores"].items():
                    ia, ib = eng.split_positions(split, sorted(set(split[0]) | set(split[1])))
                    ref = O.rank_1_divergence(O.flattening_dense(keys, vals, n, ia, ib))
                    flag = "" if abs(v - ref) <= 1e-10 * abs(ref) else "   <-- WRONG"
                    print("      ", split, float(v), ref, flag)

# ---- erickson_SVD(Method.flattening / subflattening): our per-step candidate scores next to the reference's best / runner-up ----
cases = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_erickson.json")))
traces = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_erickson_trace.json")))
for ci, (rec, tr) in enumerate(zip(cases, traces)):
    aln = dict(zip(rec["patterns"], rec["values"]))
    for method in (sp.Method.flattening, sp.Method.subflattening):
        mine = []
        got = sp.erickson_SVD(aln, method=method, trace=mine)
        print("case", ci, "n", rec["n"], method.name)
        for step, (r, m) in enumerate(zip(tr[method.name], mine)):
            sc = sorted((float(v), k) for k, v in m["scores"].items())
            print("  step", step, "ref chosen", r["chosen"], "best", r["best"], "second", r["second"])
            print("         mine chosen", [list(x) for x in m["chosen"]], "lowest three:", [(round(v, 9), k) for v, k in sc[:3]])
