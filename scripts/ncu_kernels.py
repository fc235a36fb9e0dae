"""Small driver that launches every streaming / scoring kernel a few times on realistic inputs, for `ncu --set full`.
usage: python scripts/ncu_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splitp_b200 as sp
eng = sp.engine
# 12 taxa x 32M sites: pack, count (direct), flatten scatter, pair statistics
tree = sp.trees.balanced_tree(12, 0.05)
codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), 32_000_000, seed=2)
aln = eng.pack(codes)
tab = eng.count_patterns(aln)
raw = eng.pair_raw(aln)
# 20 taxa x 10M sites: count (hash), pair statistics, subflattening scores of 65536 splits
tree20 = sp.trees.balanced_tree(20, 0.05)
codes20 = sp.simulation.simulate_codes(tree20, sp.simulation.GTR((0.1, 0.2, 0.3, 0.4), (1, 2, 3, 4, 5, 6)), 10_000_000, seed=3)
aln20 = eng.pack(codes20)
tab20 = eng.count_patterns(aln20)
pt = eng.pair_tables_from_alignment(aln20)
import numpy as np
masks = np.arange(3, 3 + 65536, dtype=np.uint64) * 7 % (1 << 20)
masks = masks[(masks & (masks - 1)) != 0]
sc = eng.subflatten_scores(pt, masks)
# scoring stage of c2: 16 splits of size 6, 16 of size 3
splits = [s for s in sp.all_splits(tree) if len(s[0]) in (3, 6)]
idx = [eng.split_positions(s, tree.taxa) for s in splits if len(s[0]) == 3][:16] + \
      [eng.split_positions(s, tree.taxa) for s in splits if len(s[0]) == 6][:16]
scorer = eng.CountScorer(tab)
out = scorer.score_many(idx)
torch.cuda.synchronize()
print("ok", tab.num, tab20.num, float(sc.min()), float(out.max()))
