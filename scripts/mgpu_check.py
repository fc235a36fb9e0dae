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
    if n <= eng.DIRECT_MAX_TAXA:  # compacted-list exchange of the direct table (the bench path)
        got3 = eng.count_patterns(loc, gather_fn=spd.make_gather_fn())
        assert torch.equal(ref.keys, got3.keys) and torch.equal(ref.counts, got3.counts) and ref.divisor == got3.divisor, (n, "gather")
        from splitp_b200 import batch
        idx20 = [eng.split_positions(s_, tree.taxa) for s_ in list(sp.all_splits(tree))[:60]]
        sharded = batch.SplitScorer(idx20, None, sp.Method.flattening, rank, world).device_scores(codes[:, b:e].contiguous())
        single = eng.score_splits_counts(ref, idx20)
        assert torch.equal(sharded, single), (n, "SplitScorer")
        repl = batch.SplitScorer(idx20, None, sp.Method.flattening, rank, world, sites="replicated").device_scores(codes)
        assert torch.equal(repl, single), (n, "SplitScorer, replicated sites")
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
# wide (128-bit) keys: sharded compression + merge equals the single-GPU table; thin-split scores agree
n, N = 64, 200_001
tree = sp.trees.balanced_tree(n, 0.02)
codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), N, seed=64)
codes[:, :11] = 3  # the all-T pattern (EMPTY marker of the wide table)
wide, valid, n_, N_ = eng.pack_wide(codes)
ref = eng.count_patterns_wide(wide, valid, n_, N_)
got = spd.count_patterns_wide_sharded(wide, valid, n_, N_, rank, world)
kr, cr = ref.compact()
kg, cg = got.compact()
assert torch.equal(kr, kg) and torch.equal(cr, cg) and ref.divisor == got.divisor
b, e = spd.shard_range(N, rank, world, 32)
lw, lv, _, lN = eng.pack_wide(codes[:, b:e].contiguous())
got2 = spd.count_patterns_wide_sharded(lw, lv, n_, lN, rank, world, local=True)
kg2, cg2 = got2.compact()
assert torch.equal(kr, kg2) and torch.equal(cr, cg2)
# hash-partitioned merge: the partitions are disjoint, their union is the global table, replication restores it
part, usable_all = spd.count_patterns_wide_partitioned(lw, lv, n_, lN, rank, world)
assert usable_all == int(ref.divisor)
kp, cp = part.compact()
if world > 1:
    owner = spd._owner_of(kp, world)
    assert bool((owner == rank).all()), "a partition holds a key it does not own"
    sizes = spd.all_gather_varlen(torch.tensor([kp.shape[0]], dtype=torch.int64, device=kp.device))
    assert int(sizes.sum().item()) == kr.shape[0], (sizes.tolist(), kr.shape[0])
    assert int(sizes.max().item()) < 0.75 * kr.shape[0] or world == 1  # memory per rank shrinks
full = spd.replicate_wide_table(part, n_, usable_all, world)
kf, cf = full.compact()
assert torch.equal(kr, kf) and torch.equal(cr, cf)
sides = [[0, 1], [5, 40], [62, 63]]
assert torch.equal(eng.thin_split_scores(ref, sides), eng.thin_split_scores(got2, sides))
if world > 1:
    dist.barrier()
print(f"rank {rank}/{world}: multi-GPU parity ok", flush=True)
if world > 1:
    dist.destroy_process_group()
