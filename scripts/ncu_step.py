"""One step of the c2 / c3 workload bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
usage: python scripts/ncu_step.py [--per-size M] [--sites N] [--workload c2|c3]     (M = 0: all splits)"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splitp_b200 as sp
ap = argparse.ArgumentParser()
ap.add_argument("--per-size", type=int, default=0)
ap.add_argument("--sites", type=int, default=None)
ap.add_argument("--workload", default="c2")
a = ap.parse_args()
eng = sp.engine
if a.workload == "c2":
    n = 12
    tree = sp.trees.balanced_tree(n, 0.05)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), a.sites or 1_000_000, seed=2)
    splits = list(sp.all_splits(tree))
    if a.per_size:
        keep, seen = [], {}
        for s in splits:
            k = len(s[0])
            seen[k] = seen.get(k, 0) + 1
            if seen[k] <= a.per_size:
                keep.append(s)
        splits = keep
    idx = [eng.split_positions(s, tree.taxa) for s in splits]

    def step(scorer=None):
        aln = eng.pack(codes, want_planes=False)
        tab = eng.count_patterns(aln)
        scorer = scorer or eng.CountScorer(tab)
        scorer.table = tab
        return scorer, scorer.score_many(idx)
    scorer, out = step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    scorer, out = step(scorer)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    n = 20
    tree = sp.trees.balanced_tree(n, 0.05)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR((0.1, 0.2, 0.3, 0.4), (1, 2, 3, 4, 5, 6)), a.sites or 10_000_000, seed=3)
    splits = list(sp.all_splits(tree))
    idx = [eng.split_positions(s, tree.taxa) for s in splits]
    ma, mb = eng.masks_from_splits(idx)

    def step():
        aln = eng.pack(codes, want_sm=False)
        pt = eng.pair_tables_from_alignment(aln)
        return eng.subflatten_scores(pt, ma, mb)
    out = step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", len(idx), float(out.min()), float(out.max()), "launches", eng.lib.spb_launch_count())
