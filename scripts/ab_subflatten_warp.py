"""A/B of the two subflattening scorers on all 524,267 splits of a 20-taxon alignment: block-wide Jacobi kernel against
the warp-per-split tridiagonal / bisection kernel (SPB_SUBFLATTEN_WARP=1).  Prints the largest difference relative to
the parity tolerance and both timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import splitp_b200 as sp
eng = sp.engine
n, N = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
tree = sp.trees.balanced_tree(n, 0.05)
codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR((0.1, 0.2, 0.3, 0.4), (1, 2, 3, 4, 5, 6)), N, seed=3)
pt = eng.pair_tables_from_alignment(eng.pack(codes, want_sm=False))
idx = [eng.split_positions(s, tree.taxa) for s in sp.all_splits(tree)]
ma_np, mb_np = eng.masks_from_splits(idx)
ma = torch.from_numpy(ma_np.view(np.int64)).cuda()
mb = torch.from_numpy(mb_np.view(np.int64)).cuda()
res = {}
for mode in ("0", "1"):
    os.environ["SPB_SUBFLATTEN_WARP"] = mode
    out = eng.subflatten_scores(pt, ma, mb)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        out = eng.subflatten_scores(pt, ma, mb)
    b.record()
    torch.cuda.synchronize()
    res[mode] = (out.cpu().numpy(), a.elapsed_time(b) / 3)
ref, got = res["0"][0], res["1"][0]
eps = np.finfo(float).eps
tol = np.maximum(1e-9, 64 * eps / np.maximum(ref * ref, 1e-300))
rel = np.abs(got - ref) / np.maximum(ref, 1e-300)
print(f"splits {len(ref)}  block {res['0'][1]:.2f} ms  warp {res['1'][1]:.2f} ms  speed-up {res['0'][1] / res['1'][1]:.2f}x")
print(f"finite {np.isfinite(got).all()}  max |diff| / tolerance {np.nanmax(rel / tol):.4f}  max rel diff {np.nanmax(rel):.3e}  "
      f"score range {ref.min():.3e} .. {ref.max():.3e}")
