"""Times the library (torch) operations that remain inside product routes, next to the stage they belong to (VERDICT r1 item
10: "replace or time").  One B200; prints one JSON object.

  * torch.sort of a hashed pattern table (engine.count_patterns, n > 12) against the counting kernels of the same call;
  * torch.unique in engine.marginalise for sub-alignments above 12 taxa (Alignment.sub_alignment), against the own-kernel
    route used up to 12 taxa;
  * torch.unique (twice) in engine._flatten_reduced_large (reduced flattening with a side above 13 taxa).
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import splitp_b200 as sp  # noqa: E402
from splitp_b200 import engine as eng  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), out


out = {}
n, N = 20, 10_000_000
tree = sp.trees.balanced_tree(n, 0.05)
codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), N, seed=3)
aln = eng.pack(codes)
ms_sorted, tab = timed(lambda: eng.count_patterns(aln, sort=True))
ms_unsorted, tab_u = timed(lambda: eng.count_patterns(aln, sort=False))
ms_sort, _ = timed(lambda: torch.sort(tab_u.keys))
out["count_patterns_20taxa_1e7"] = {"patterns": int(tab.num), "ms_with_sort": round(ms_sorted, 3), "ms_without_sort": round(ms_unsorted, 3),
                                    "ms_torch_sort_alone": round(ms_sort, 3)}
sub14 = list(range(14))
ms_m14, t14 = timed(lambda: eng.marginalise(tab, sub14))
sub12 = list(range(12))
ms_m12, t12 = timed(lambda: eng.marginalise(tab, sub12))
rows, _ = eng.flatten_coo(tab, sub12, [])
ms_u12, _ = timed(lambda: torch.unique(rows, return_inverse=True))
out["marginalise"] = {"to_14_taxa_ms (torch.unique route)": round(ms_m14, 3), "patterns_14": int(t14.num),
                      "to_12_taxa_ms (own kernels)": round(ms_m12, 3), "patterns_12": int(t12.num),
                      "torch.unique alone on the 12-taxon indices ms": round(ms_u12, 3)}
ia, ib = list(range(6)), list(range(6, 20))
ms_red, F = timed(lambda: eng._flatten_reduced_large(tab, ia, ib, True), reps=3)
out["reduced_flattening_6|14_of_20"] = {"ms": round(ms_red, 3), "shape": list(F.shape)}
print(json.dumps(out))
