"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink).

The path shards two ways (BASELINE.json north_star):
  * counting  -- alignment SITES are split into contiguous ranges, one per rank; every rank counts its
                 range and the count tables are summed with an allreduce (direct-indexed table, n <= 12
                 taxa) or, for hashed tables, gathered as (key, count) lists and merged on every rank.
                 The pair-statistics table of the subflattening path is reduced the same way.
  * scoring   -- the SPLIT list is cut into contiguous ranges, one per rank; no collective during the
                 compute; the scores are all-gathered at the end.
All reductions are integer sums, so the sharded result is bit-identical to the single-GPU result.
The host-side logic (ranges, gathers) is backend-agnostic and is tested with gloo on CPU tensors.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env():
    """Process-group setup under torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    return rank, local, world


def shard_range(total, rank, world, align=1):
    """Contiguous [begin, end) share of `total` units for `rank`; boundaries are multiples of `align`
    (sites are aligned to 32 so that a shard starts on a validity / bit-plane word)."""
    blocks = (total + align - 1) // align
    per, extra = divmod(blocks, world)
    b = rank * per + min(rank, extra)
    e = b + per + (1 if rank < extra else 0)
    return min(b * align, total), min(e * align, total)


def make_reduce_fn(group=None):
    """reduce_fn(tensor, op) hook for engine.count_patterns: in-place allreduce of integer tables."""
    def reduce_fn(t, op):
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return t
        if op == "sum":
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        elif op == "min_u32":  # uint32 stored in int32: widen so that 0xFFFFFFFF stays the largest value
            w = t.to(torch.int64) & 0xFFFFFFFF
            dist.all_reduce(w, op=dist.ReduceOp.MIN, group=group)
            t.copy_(torch.where(w > 0x7FFFFFFF, w - (1 << 32), w).to(torch.int32))
        else:
            raise ValueError(op)
        return t
    return reduce_fn


def make_gather_fn(group=None):
    """gather_fn(keys, counts) hook for engine.count_patterns: all-gather of the compacted per-rank pattern lists."""
    def gather_fn(keys, counts):
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return keys, counts
        return all_gather_varlen_pair(keys, counts, group)
    return gather_fn


def all_gather_varlen_pair(a, b, group=None):
    """all_gather_varlen of two equally long 1-D tensors with ONE size exchange."""
    world = dist.get_world_size(group)
    n = torch.tensor([a.shape[0]], dtype=torch.int64, device=a.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    out = []
    for t in (a, b):
        pad = torch.zeros(m, dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        out.append(torch.cat([x[:s] for x, s in zip(bufs, sizes)]))
    return out[0], out[1]


def all_gather_varlen(t, group=None):
    """Concatenation over ranks of 1-D tensors of different lengths (rank order)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def gather_scores(local_scores, total, rank, world, group=None):
    """Scores of split range shard_range(total, rank, world) from every rank -> all `total` scores."""
    if world == 1:
        return local_scores
    out = all_gather_varlen(local_scores, group)
    assert out.shape[0] == total
    return out


def shard_strided(items, rank, world):
    """Round-robin share of a list: items[rank::world].  all_splits lists splits by ascending size and the cost
    of a split grows steeply with its size, so a strided deal balances the ranks where contiguous ranges do not."""
    return items[rank::world]


def gather_strided(local_scores, total, rank, world, group=None):
    """Inverse of shard_strided for the score vector: out[r::world] = scores of rank r."""
    if world == 1:
        return local_scores
    m = (total + world - 1) // world
    pad = torch.zeros(m, dtype=local_scores.dtype, device=local_scores.device)
    pad[:local_scores.shape[0]] = local_scores
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.empty(total, dtype=local_scores.dtype, device=local_scores.device)
    for r in range(world):
        cnt = len(range(r, total, world))
        out[r::world] = bufs[r][:cnt]
    return out


# ------------------------------------------------------------------------------------------------
# cost-balanced split partition for the count-flattening scores
# ------------------------------------------------------------------------------------------------
# A strided deal gives every rank an eighth of EVERY size class: five small batches per rank, each with its own chain of
# latency-bound solver kernels (8 GPUs, 12 taxa: 3.9 ms of eigen stage per rank against 2.5 ms = an eighth of the one-GPU stage).
# Dealing contiguous runs of the class-ordered split list instead leaves most ranks with ONE class; the runs are cut at equal
# modelled cost.  The model (B200 rates measured by bench.py / profiles/r2_launches_c2.md; only RATIOS between classes matter):
#   scatter + clear  P / 3e10 s (2 P byte stores at 1e11 /s plus launch gaps) + 1.5 us of per-split launch / host overhead
#                    (calibrated on the two-GPU run: rank 0 with 393 6|6 splits 24.3 ms, rank 1 with the rest 26.5 ms)
#   Gram             max(tensor ops of the executed upper-triangle blocks at 2990 TOP/s, S0 bytes at 6 TB/s); dp4a rows <= 32: 1.5 TB/s
#   eigen            rows >= 1024: two G0 Q products at 5.9 TB/s + strip passes + correction; below: ~0.8 us
#   fixed            one solver chain per (rank, class): 700 us for Krylov classes (rows > 128), 140 us otherwise
# (constants re-fitted to the per-rank times of the 8-GPU runs, profiles/r2_bench_n8.json `partition`; SplitScorer re-fits them
# again at run time, refit_costs below)
def flattening_cost_us(n_taxa, a, patterns=65536):
    """Modelled device time of ONE count-flattening score with a short side of `a` taxa, microseconds."""
    R, Cc = 4.0 ** a, 4.0 ** (n_taxa - a)
    scatter = patterns / 3e10 * 1e6 + 1.5  # scatter + clear phase as measured (2.0 us at 60 k patterns) + 1.5 us of launch / host overhead
    rows_pad = max(R, 128.0) if R > 32 else R
    pitch = max(Cc, 128.0)
    if R >= 256:
        T = rows_pad / 256.0
        gram = T * (T + 1) / 2.0 * (256.0 * 256.0 * pitch * 2.0) / 2990e12 * 1e6
    else:
        gram = 0.0
    gram = max(gram, rows_pad * pitch / (6e12 if R > 32 else 1.5e12) * 1e6)
    if R >= 1024:
        eigen = 2.0 * R * R * 4.0 / 5.9e12 * 1e6 + 4.0 * 130.0 * R * 8.0 / 3e12 * 1e6 + 1.0
    else:
        eigen = 0.8
    return scatter + gram + eigen


def flattening_fixed_us(a):
    return 700.0 if 4 ** a > 128 else 140.0


def partition_by_cost(classes, costs, fixed, world):
    """classes[i], costs[i]: size class and modelled cost of item i of a CLASS-ORDERED list; fixed[c]: per-rank cost of touching
    class c.  Returns `world` half-open ranges [b, e) covering the list that minimise the largest modelled rank time (its items +
    the fixed cost of every class it touches): bisection on that bottleneck, each trial a greedy left-to-right fill."""
    n = len(costs)

    def fill(limit):
        cuts, b = [], 0
        for _ in range(world):
            acc, e, cur = 0.0, b, None
            while e < n:
                add = costs[e] + (fixed[classes[e]] if classes[e] != cur else 0.0)
                if acc + add > limit and e > b:
                    break
                acc += add
                cur = classes[e]
                e += 1
            cuts.append((b, e))
            b = e
        return cuts if b == n else None

    hi = float(sum(costs)) + sum(fixed[c] for c in set(classes)) + 1.0
    lo = 0.0
    best = fill(hi) or [(0, n)] + [(n, n)] * (world - 1)
    for _ in range(50):
        mid = 0.5 * (lo + hi)
        cuts = fill(mid)
        if cuts is None:
            lo = mid
        else:
            hi, best = mid, cuts
    return best


RANK_BASE_US = 400.0  # per-rank time outside the per-split work: pack, count, compaction, gather


def refit_costs(per, fixed, rank_classes, rank_times_us, ridge=0.05):
    """Per-class costs re-fitted to MEASURED rank times.  per[a]: modelled cost per split; rank_classes[r] = {a: number of splits of
    class a on rank r}; rank_times_us[r]: measured time of rank r.  Model: t_r = RANK_BASE_US + sum_a n_ra per[a] s_a + s_f
    sum_{a on r} fixed[a]; the scale factors (one per class, one for the fixed chain costs) solve the ridge problem
    min |A s - t'|^2 + (ridge mean(t'))^2 |s - 1|^2 (few ranks leave s under-determined: it then stays at the model) and are
    clipped to [0.5, 2].  Returns (per, fixed) rescaled."""
    import numpy as np
    classes = sorted(per)
    nc = len(classes)
    A = np.zeros((len(rank_classes), nc + 1))
    t = np.zeros(len(rank_classes))
    for r, held in enumerate(rank_classes):
        t[r] = rank_times_us[r] - RANK_BASE_US
        A[r, nc] = sum(fixed[a] for a in held)
        for a, cnt in held.items():
            A[r, classes.index(a)] = cnt * per[a]
    lam = ridge * max(float(np.mean(np.abs(t))), 1.0)
    M = np.vstack([A, lam * np.eye(nc + 1)])
    rhs = np.concatenate([t, lam * np.ones(nc + 1)])
    sol = np.clip(np.linalg.lstsq(M, rhs, rcond=None)[0], 0.5, 2.0)
    return ({a: per[a] * float(sol[i]) for i, a in enumerate(classes)}, {a: fixed[a] * float(sol[nc]) for a in fixed})


def gather_by_position(local_scores, positions, total, group=None):
    """Inverse of an arbitrary partition: every rank writes its scores at `positions` (int64 device tensor) of a zero vector of
    length `total`; the SUM all-reduce then holds every score on every rank (x + 0.0 is exact)."""
    full = torch.zeros(total, dtype=local_scores.dtype, device=local_scores.device)
    if positions.numel():
        full.index_copy_(0, positions, local_scores)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full


def count_patterns_sharded(aln, rank, world, group=None, want_first=False, local=False):
    """Pattern table of the WHOLE alignment.  local=False: every rank holds the full alignment and counts the
    site range shard_range(N, r, world, 32) of it.  local=True: `aln` IS this rank's site shard (then `first`
    indices are shard-relative and not supported).  The per-rank tables are combined on every rank."""
    from . import engine
    b, e = (0, aln.N) if local else shard_range(aln.N, rank, world, 32)
    if local and want_first:
        raise NotImplementedError("first-site indices need the full alignment on every rank (local=False)")
    if world == 1:
        return engine.count_patterns(aln, want_first=want_first)
    if aln.n <= engine.DIRECT_MAX_TAXA:
        return engine.count_patterns(aln, b, e, want_first=want_first, reduce_fn=make_reduce_fn(group))
    # hashed tables: gather the per-rank (key, count) lists and merge them into one table on every rank
    local = engine.count_patterns(aln, b, e, want_first=want_first, sort=False)
    keys = all_gather_varlen(local.keys, group)
    counts = all_gather_varlen(local.counts, group)
    first = all_gather_varlen(local.first, group) if want_first else None
    usable = torch.tensor([int(local.divisor)], dtype=torch.int64, device=keys.device)
    dist.all_reduce(usable, group=group)
    merged = engine.merge_tables(aln.n, keys, counts, first)
    merged.divisor = float(int(usable.item()))
    merged.taxa = aln.taxa
    return merged


def pair_tables_sharded(aln, rank, world, group=None, as_counts=False):
    """Pair tables of the whole alignment: each rank reduces its word range, raw statistics are summed."""
    from . import engine
    words = (aln.N + 31) // 32
    b, e = shard_range(words, rank, world)
    raw = engine.pair_raw(aln, b, e)
    if world > 1:
        dist.all_reduce(raw, group=group)
    return engine.pair_finalize(raw, aln.n, 0.0 if as_counts else float(int(raw[-1].item())))


def _owner_of(keys, world):
    """Owner rank of every pattern key (hash partition): a multiplicative mix of the key words, reduced modulo world.
    keys: int64 [P] (uint64 keys) or int64 [P, 2] (128-bit keys {lo, hi})."""
    k = keys if keys.dim() == 1 else keys[:, 0] ^ (keys[:, 1] * -7046029254386353131)  # 0x9E3779B97F4A7C15 as int64
    k = (k ^ (k >> 31)) * -4658895280553007687  # 0xBF58476D1CE4E5B9 as int64; arithmetic shift keeps the mix deterministic
    k = k ^ (k >> 29)
    return (k & 0x7FFFFFFF) % world


def exchange_by_owner(keys, counts, world, group=None):
    """Hash-partitioned exchange: every (key, count) entry travels to the rank that owns its key
    (all_to_all with uneven splits); returns the entries this rank owns, duplicates from different ranks included.
    On NVSwitch the all-to-all is uniform-cost, so the partition needs no topology awareness."""
    if world == 1:
        return keys, counts
    owner = _owner_of(keys, world)
    order = torch.argsort(owner, stable=True)
    keys, counts, owner = keys[order].contiguous(), counts[order].contiguous(), owner[order]
    send = torch.bincount(owner, minlength=world).to(torch.int64)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    send_l, recv_l = send.tolist(), recv.tolist()
    total = int(sum(recv_l))
    width = 1 if keys.dim() == 1 else keys.shape[1]
    rk = torch.empty((total,) if width == 1 else (total, width), dtype=keys.dtype, device=keys.device)
    rc = torch.empty(total, dtype=counts.dtype, device=counts.device)
    dist.all_to_all_single(rk, keys, output_split_sizes=recv_l, input_split_sizes=send_l, group=group)
    dist.all_to_all_single(rc, counts, output_split_sizes=recv_l, input_split_sizes=send_l, group=group)
    return rk, rc


def count_patterns_wide_partitioned(wide, valid, n, N, rank, world, group=None, local=True):
    """128-bit-key pattern compression with a hash-PARTITIONED result: every rank compresses its site shard, the
    (key, count) lists are exchanged by owner (exchange_by_owner) and merged locally, so that rank r ends up holding
    the patterns it owns with their GLOBAL counts: memory and merge work per rank shrink with the number of ranks (the
    replicate-everything form, count_patterns_wide_sharded, re-hashes every pattern of every rank on every rank).
    Returns (partition WideTable, usable sites of the whole alignment)."""
    from . import engine
    b, e = (0, N) if local else shard_range(N, rank, world, 32)
    mine = engine.count_patterns_wide(wide, valid, n, N, b, e)
    usable = torch.tensor([int(mine.divisor)], dtype=torch.int64, device=wide.device)
    if world == 1:
        return mine, int(usable.item())
    dist.all_reduce(usable, group=group)
    keys, counts = mine.compact(sort=False)
    rk, rc = exchange_by_owner(keys, counts, world, group)
    part = engine.merge_wide_tables(n, rk, rc, int(usable.item()))
    return part, int(usable.item())


def replicate_wide_table(part, n, usable, world, group=None):
    """Full table on every rank from the hash partitions: the merged, duplicate-free partitions are all-gathered and
    inserted (what consumers that look patterns up across the whole table -- the thin-split Gram -- need)."""
    from . import engine
    if world == 1:
        return part
    keys, counts = part.compact(sort=False)
    lo = all_gather_varlen(keys[:, 0].contiguous(), group)
    hi = all_gather_varlen(keys[:, 1].contiguous(), group)
    cnt = all_gather_varlen(counts, group)
    return engine.merge_wide_tables(n, torch.stack([lo, hi], dim=1), cnt, usable)


def count_patterns_wide_sharded(wide, valid, n, N, rank, world, group=None, local=False):
    """128-bit-key pattern table of the whole alignment: every rank compresses its site range (or its local shard),
    the per-rank (key, count) lists are all-gathered and merged into one table on every rank.  At 64 taxa nearly
    every site is a distinct pattern, so this exchange moves about as many bytes as the packed input."""
    from . import engine
    b, e = (0, N) if local else shard_range(N, rank, world, 32)
    mine = engine.count_patterns_wide(wide, valid, n, N, b, e)
    if world == 1:
        return mine
    keys, counts = mine.compact(sort=False)
    lo = all_gather_varlen(keys[:, 0].contiguous(), group)
    hi = all_gather_varlen(keys[:, 1].contiguous(), group)
    cnt = all_gather_varlen(counts, group)
    usable = torch.tensor([int(mine.divisor)], dtype=torch.int64, device=lo.device)
    dist.all_reduce(usable, group=group)
    return engine.merge_wide_tables(n, torch.stack([lo, hi], dim=1), cnt, int(usable.item()))
