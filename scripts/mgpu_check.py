"""torchrun --nproc-per-node N scripts/mgpu_check.py : sharded counting / pair statistics equal the single-GPU result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import splitp_b200 as sp
from splitp_b200 import distributed as spd
eng = sp.engine
rank, local, world = spd.init_from_env()
for n, N in ((10, 300_001), (16, 500_000)):
    tree = sp.trees.balanced_tree(n, 0.05)
    codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), N, seed=n)
    aln = eng.pack(codes)
    ref = eng.count_patterns(aln, want_first=True)
    got = spd.count_patterns_sharded(aln, rank, world, want_first=True)
    assert torch.equal(ref.keys, got.keys) and torch.equal(ref.counts, got.counts) and torch.equal(ref.first, got.first), (n, "full")
    assert ref.divisor == got.divisor
    b, e = spd.shard_range(N, rank, world, 32)
    loc = eng.pack(codes[:, b:e].contiguous())
    got2 = spd.count_patterns_sharded(loc, rank, world, local=True)
    assert torch.equal(ref.keys, got2.keys) and torch.equal(ref.counts, got2.counts), (n, "local")
    pt_ref = eng.pair_tables_from_alignment(aln, as_counts=True)
    pt = spd.pair_tables_sharded(aln, rank, world, as_counts=True)
    assert torch.equal(pt_ref.N, pt.N) and torch.equal(pt_ref.T, pt.T)
    splits = list(sp.all_splits(tree))[:200]
    idx = [eng.split_positions(s, tree.taxa) for s in splits]
    pb, pe = spd.shard_range(len(idx), rank, world)
    ma, mb = eng.masks_from_splits(idx[pb:pe])
    sc = spd.gather_scores(eng.subflatten_scores(pt, ma, mb), len(idx), rank, world)
    ma, mb = eng.masks_from_splits(idx)
    assert torch.equal(sc, eng.subflatten_scores(pt_ref, ma, mb))
    if n == 10:
        full = eng.score_splits_counts(ref, idx[:40])
        mine = eng.score_splits_counts(got, idx[:40][spd.shard_range(40, rank, world)[0]:spd.shard_range(40, rank, world)[1]])
        assert torch.equal(spd.gather_scores(mine, 40, rank, world), full)
if world > 1:
    dist.barrier()
print(f"rank {rank}/{world}: multi-GPU parity ok", flush=True)
if world > 1:
    dist.destroy_process_group()
