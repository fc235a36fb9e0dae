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
    for rep in range(5):
        got = [list(map(list, s)) for s in sp.erickson_SVD(aln, method=sp.Method.mutual_information)]
        print("  rep", rep, "matches golden:", got == rec["mutual_information"], got if got != rec["mutual_information"] else "")
