"""Batched entry point of the engine: scores of MANY splits of one alignment in one call.

This is the call that replaces the reference's per-split loops

    for split in splitp.all_splits(tree):                     # README.md:37-41, examples/playground.py:29-39
        splitp.split_score(splitp.flattening(split, aln, FlatFormat.reduced))      # or subflattening(split, aln)

and the inner loop of `erickson_SVD` (phylogenetics.py:126-140).  `score_splits` takes the alignment as HOST data
(what `parsers.fasta.read_alignment_from_file` returns, or a byte / code matrix, or the 2-bit planes written by
`pack_planes_host`), copies it to the device, and returns one float64 score per split as a numpy array:

    Method.flattening     pack -> pattern count (kernel 1) -> per split: u8 scatter (kernel 2) -> exact integer Gram on the
                          tensor cores (kernel 4) -> eigen-solver (kernel 5); scores equal those of the reduced / dense /
                          sparse flattening (dropping all-zero rows and columns does not change the singular values)
    Method.subflattening  pack -> pair tables -> H N H^T -> per split: gather + Gram + eigenvalues (kernel 3)

Under `torch.distributed` (one process per GPU) the splits are dealt to the ranks for scoring (flattening scores: contiguous
runs of the size-class-ordered list cut at equal modelled cost; subflattening scores: round-robin) and every rank receives
the scores of all splits.  The counting stage has two forms: `sites="shard"` -- every rank passes ITS contiguous site
shard (distributed.shard_range) and the pattern tables are merged -- or `sites="replicated"` -- every rank passes the
WHOLE alignment and counts it itself, with no exchange at all.  The second form wins while counting the whole alignment
is cheaper than the exchange (`replicate_sites`: 12 taxa x 10^6 sites count in 0.16 ms, the list exchange between 8
ranks takes 1.5 ms).

`PhaseTimer` collects per-phase device time (CUDA events on the launch stream) for bench.py's `phase_ms`.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine
from .engine import PhaseTimer, _span  # noqa: F401  (PhaseTimer is re-exported: bench.py's `phase_ms`)
from .enums import Method


# --------------------------------------------------------------------------------------------
# host-side 2-bit format (what a FASTA -> packed converter would write once)
# --------------------------------------------------------------------------------------------
def pack_planes_host(codes):
    """uint8 [n, N] codes 0..3 (anything else = unusable site, fasta.py:55-57) -> (planes uint32 [n, 2, Wp],
    valid uint32 [Wp]) in the device layout of include/splitp_b200.h: 2 bits per base instead of 8."""
    codes = np.asarray(codes, dtype=np.uint8)
    n, N = codes.shape
    Wp = int(engine.lib.spb_plane_words(N))
    pad = Wp * 32 - N
    ok = (codes <= 3).all(axis=0)
    planes = np.zeros((n, 2, Wp), dtype=np.uint32)
    for t in range(n):
        row = np.where(ok, codes[t], 0)
        for bit in range(2):
            bits = np.concatenate([(row >> bit) & 1, np.zeros(pad, np.uint8)]).astype(np.uint8)
            planes[t, bit] = np.packbits(bits, bitorder="little").view(np.uint32)
    valid = np.packbits(np.concatenate([ok, np.zeros(pad, bool)]).astype(np.uint8), bitorder="little").view(np.uint32)
    return planes, valid.copy()


class HostPlanes:
    """2-bit packed alignment on the host (pinned when `pin=True`): the e2e input format of the subflattening path."""

    def __init__(self, planes, valid, n, N, pin=False):
        self.planes = torch.from_numpy(np.ascontiguousarray(planes).view(np.int32))
        self.valid = torch.from_numpy(np.ascontiguousarray(valid).view(np.int32))
        if pin:
            self.planes, self.valid = self.planes.pin_memory(), self.valid.pin_memory()
        self.n, self.N = int(n), int(N)

    @property
    def nbytes(self):
        return self.planes.numel() * 4 + self.valid.numel() * 4

    def to_device(self):
        dev = engine.device()
        return engine.DeviceAlignment(self.n, self.N, None, self.planes.to(dev, non_blocking=True),
                                      self.valid.to(dev, non_blocking=True))


# --------------------------------------------------------------------------------------------
def _as_alignment(alignment, want_sm, want_planes):
    """Anything the caller may hold -> engine.DeviceAlignment (+ taxa if known)."""
    if isinstance(alignment, engine.DeviceAlignment):
        return alignment
    if isinstance(alignment, HostPlanes):
        return alignment.to_device()
    if isinstance(alignment, dict):  # {taxon: sequence string}: parsers.fasta.read_alignment_from_file
        from .parsers.fasta import _to_bytes_matrix
        return engine.pack(_to_bytes_matrix(list(alignment.values())), is_ascii=True, taxa=list(alignment.keys()),
                           want_sm=want_sm, want_planes=want_planes)
    if isinstance(alignment, torch.Tensor):
        chars = alignment if alignment.is_cuda else alignment.to(engine.device(), non_blocking=True)
        return engine.pack(chars, want_sm=want_sm, want_planes=want_planes)
    arr = np.asarray(alignment)
    is_ascii = arr.dtype.kind in "SU" or (arr.dtype == np.uint8 and arr.size and arr.max() > 3 and arr.max() != 255)
    return engine.pack(arr.astype(np.uint8, copy=False), is_ascii=bool(is_ascii), want_sm=want_sm, want_planes=want_planes)


def _positions(splits, taxa):
    out = []
    for left, right in splits:
        if taxa is not None and len(left) and not isinstance(left[0], (int, np.integer)):
            out.append(engine.split_positions((left, right), taxa))
        else:
            out.append((list(left), list(right)))
    return out


# Counting n taxa x N sites costs about 1.1 ms per 1.2e9 site-taxa (count_class_kernel, B200); merging the ranks' pattern
# lists costs 0.4 ms (2 ranks) .. 1.5 ms (8 ranks) whatever the size.  Below this many site-taxa every rank counts everything.
REPLICATE_MAX_SITE_TAXA = 2.0e8


def replicate_sites(n_taxa, n_sites, world):
    """True when `SplitScorer(..., sites="replicated")` is the faster form of the counting stage (flattening scores)."""
    return world > 1 and float(n_taxa) * float(n_sites) <= REPLICATE_MAX_SITE_TAXA


class SplitScorer:
    """Reusable state of `score_splits` for one split list: encoded splits, device masks, the CountScorer's buffers.
    Keeping one per (alignment shape, split list) is what a serving loop would do; `score_splits` builds a throw-away
    one."""

    def __init__(self, splits, taxa=None, method=Method.flattening, rank=0, world=1, group=None, sites="shard", refit_steps=(2,)):
        from . import distributed as spd
        if sites not in ("shard", "replicated"):
            raise ValueError('sites must be "shard" or "replicated"')
        self.method, self.rank, self.world, self.group = method, rank, world, group
        self.count_world = 1 if sites == "replicated" else world  # ranks that share the counting stage
        self.idx_all = _positions(splits, taxa)
        self.S = len(self.idx_all)
        self.positions = None
        self._steps = 0
        self._cls = None
        # calls after which the multi-GPU split partition is re-cut from the measured rank times (the first call pays the allocations;
        # the call after a re-cut reallocates batch buffers, so two listed steps must be at least two calls apart)
        self.refit_steps = tuple(refit_steps)
        if method == Method.flattening and world > 1 and self.S:
            # contiguous runs of the class-ordered list, cut at equal modelled cost (distributed.partition_by_cost): most ranks
            # then score ONE size class in one batch instead of an eighth of every class
            self._n = len(self.idx_all[0][0]) + len(self.idx_all[0][1])
            self._cls = [min(len(ia), len(ib)) for ia, ib in self.idx_all]
            self._per = {a: spd.flattening_cost_us(self._n, a) for a in set(self._cls)}
            self._fixed = {a: spd.flattening_fixed_us(a) for a in self._per}
            self._cut()
        else:
            self.idx_mine = spd.shard_strided(self.idx_all, rank, world)
        self.reduce_fn = spd.make_reduce_fn(group) if self.count_world > 1 else None
        self.gather_fn = spd.make_gather_fn(group) if self.count_world > 1 else None
        self.scorer = None
        self.masks = None
        self.timer = None
        self.refit_enabled = os.environ.get("SPB_PARTITION_REFIT", "1") != "0"
        self.rank_times_us = None
        if method == Method.subflattening:
            ma, mb = engine.masks_from_splits(self.idx_mine)
            dev = engine.device()
            self.masks = (torch.from_numpy(ma.view(np.int64)).to(dev), torch.from_numpy(mb.view(np.int64)).to(dev))
        elif method != Method.flattening:
            raise NotImplementedError("score_splits: Method.flattening or Method.subflattening")

    REFIT_DAMPING = 0.5

    def _cut(self):
        """(Re)partitions the class-ordered split list from the current per-class costs."""
        from . import distributed as spd
        per, cls = self._per, self._cls
        order = sorted(range(self.S), key=lambda i: (-per[cls[i]], i))
        ranges = spd.partition_by_cost([cls[i] for i in order], [per[cls[i]] for i in order], self._fixed, self.world)
        self._held = []
        for b, e in ranges:
            held = {}
            for i in order[b:e]:
                held[cls[i]] = held.get(cls[i], 0) + 1
            self._held.append(held)
        mine = order[ranges[self.rank][0]:ranges[self.rank][1]]
        self.idx_mine = [self.idx_all[i] for i in mine]
        self.positions = mine

    def _refit(self, my_us):
        """All ranks exchange the device time of their share, re-fit the per-class costs (distributed.refit_costs) and re-cut."""
        import torch.distributed as dist

        from . import distributed as spd
        mine = torch.tensor([my_us], dtype=torch.float64, device=engine.device())
        every = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=self.group)
        times = [float(x.item()) for x in every]
        self.rank_times_us = times
        # a call that paid a one-off cost (the call after a re-cut reallocates its batch buffers: measured 42 ms instead of 21 ms)
        # says nothing about the steady state: keep the partition when a rank's measured / modelled ratio is more than 30 % off the median ratio
        ratios = [t_us / (spd.RANK_BASE_US + sum(c * self._per[a] + self._fixed[a] for a, c in held.items()))
                  for held, t_us in zip(self._held, times)]
        middle = sorted(ratios)[len(ratios) // 2]
        if any(not 0.7 * middle <= r <= 1.3 * middle for r in ratios):
            return
        per, fixed = spd.refit_costs(self._per, self._fixed, self._held, times)
        # half a step towards the fitted costs: the measured call is not quite the steady state (2 GPUs: the rank holding 405 6|6
        # matrices measured 25.3 ms in its second call but runs 24.2 ms later; the full step overshot, 23.3 | 24.9 ms)
        d = self.REFIT_DAMPING
        self._per = {a: self._per[a] + d * (per[a] - self._per[a]) for a in per}
        self._fixed = {a: self._fixed[a] + d * (fixed[a] - self._fixed[a]) for a in fixed}
        self._cut()

    def device_scores(self, alignment, gram_hook=None):
        """Scores of ALL splits as a device tensor.  `alignment`: this rank's site shard, or the whole alignment when the
        scorer was built with sites="replicated" (any form _as_alignment takes)."""
        import torch.distributed as dist

        from . import distributed as spd
        t = self.timer
        self._steps += 1
        refit = self._cls is not None and self._steps in self.refit_steps and self.refit_enabled
        if refit:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        if self.method == Method.flattening:
            with _span(t, "h2d+pack"):
                aln = _as_alignment(alignment, want_sm=True, want_planes=False)
            with _span(t, "count"):
                if aln.n <= engine.DIRECT_MAX_TAXA:
                    table = engine.count_patterns(aln, gather_fn=self._timed_gather if self.gather_fn else None)
                elif self.count_world > 1:  # hashed tables: gather + hash merge on every rank
                    table = spd.count_patterns_sharded(aln, self.rank, self.world, self.group, local=True)
                else:
                    table = engine.count_patterns(aln)
            if self.scorer is None:
                self.scorer = engine.CountScorer(table)
            self.scorer.table = table
            self.scorer.timer = t
            out = self.scorer.score_many(self.idx_mine, big_hook=gram_hook)
        else:
            with _span(t, "h2d+pack"):
                aln = _as_alignment(alignment, want_sm=False, want_planes=True)
            with _span(t, "pairs"):
                raw = engine.pair_raw(aln)
            if self.count_world > 1:
                with _span(t, "allreduce"):
                    dist.all_reduce(raw, group=self.group)
            with _span(t, "subflatten+score"):
                pt = engine.pair_finalize(raw, aln.n, -1.0)  # divisor = usable sites, read on the device: no host sync
                out = engine.subflatten_scores(pt, self.masks[0], self.masks[1])
        if refit:
            ev1.record()
        with _span(t, "gather"):
            if self.positions is not None:
                if not isinstance(self.positions, torch.Tensor):
                    self.positions = torch.tensor(self.positions, dtype=torch.int64, device=out.device)
                res = spd.gather_by_position(out, self.positions, self.S, self.group)
            else:
                res = spd.gather_strided(out, self.S, self.rank, self.world, self.group)
        if refit:
            ev1.synchronize()
            self._refit(ev0.elapsed_time(ev1) * 1e3)
        return res

    def _timed_gather(self, keys, counts):
        with _span(self.timer, "exchange"):
            return self.gather_fn(keys, counts)

    def __call__(self, alignment):
        return self.device_scores(alignment).cpu().numpy()


def score_splits(alignment, splits, method=Method.flattening, taxa=None):
    """One float64 score per split (numpy array, order of `splits`).

    alignment  {taxon: sequence} dict, uint8 [n_taxa, n_sites] matrix of ASCII bytes or codes 0..3 (host or device),
               `HostPlanes`, or `engine.DeviceAlignment`
    splits     pairs (left, right) of taxon labels (with `taxa`, or the dict's key order) or of taxon positions
    method     Method.flattening (scores of the flattening, constructions.py:7-102 + phylogenetics.py:280-300) or
               Method.subflattening (constructions.py:108-198)
    """
    if taxa is None and isinstance(alignment, dict):
        taxa = list(alignment.keys())
    rank = world = None
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    return SplitScorer(splits, taxa, method, rank or 0, world or 1)(alignment)
