#!/usr/bin/env python
"""Benchmark of the SplitP hot path on B200 (contract: see the task statement; metric: BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3] [--impl reference]

Workload c2 (default; BASELINE.json configs[1]): balanced 12-taxon Jukes-Cantor tree, branch length 0.05,
1,000,000 sites (seeded synthetic alignment), ALL 2,035 splits scored from dense count flattenings
(4^a x 4^b, 4096 x 4096 for the 462 6|6 splits) through the exact-integer tensor-core Gram + eigen-solver.
One step = one pass of the whole path: pack -> pattern count -> (count allreduce) -> per split
flatten + Gram + high-part correction + score -> (score all-gather).
Workload c3 (configs[2], scaled by --sites): 20-taxon GTR tree, pair tables from the bit planes, subflattening
scores of all 524,267 splits.

N > 1 (torchrun): sites are sharded for counting (integer allreduce of the count table / pair statistics)
and splits are sharded for scoring; total work is fixed => "scaling": "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(n=12, sites=1_000_000, bl=0.05, model="JC", seed=2,
               desc="BASELINE configs[1]: balanced 12-taxon JC tree bl=0.05, 1M sites, all 2035 splits, dense count "
                    "flattenings (6|6 = 4096x4096), exact u8 tensor-core Gram + eigen score"),
    "c4": dict(n=64, sites=10_000_000, bl=0.05, model="JC", seed=4, pair_splits=True,
               desc="BASELINE configs[3] (scaled: 10M sites by default, --sites 100000000 for the full size): balanced 64-taxon JC "
                    "tree bl=0.05, 128-bit pattern compression + reduced flattening scores of all 2016 2|62 splits"),
    "c5": dict(n=32, sites=1_000_000, bl=0.05, model="JC", seed=5, random_splits=100_000,
               desc="BASELINE configs[4]: 10^5 random splits (side sizes 2..16, numpy default_rng(5)) of a balanced 32-taxon JC "
                    "tree bl=0.05, 1M sites, subflattening scores"),
    "c3": dict(n=20, sites=10_000_000, bl=0.05, model="GTR", seed=3,
               desc="BASELINE configs[2]: balanced 20-taxon GTR tree bl=0.05, 10M sites, subflattening scores of all "
                    "524267 splits"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--sites", type=int, default=None, help="override the number of sites (parity/debug runs)")
    ap.add_argument("--max-splits", type=int, default=None, help="score only the first M splits (debug runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_splits(splits_mod, tree, wl):
    """All splits of the tree (all_splits order), or `random_splits` random ones for the sweep workload."""
    if wl.get("pair_splits"):
        return list(splits_mod.all_splits(tree, size=2))
    if not wl.get("random_splits"):
        return list(splits_mod.all_splits(tree))
    rng = np.random.default_rng(5)
    taxa, n, out = list(tree.taxa), len(tree.taxa), []
    for _ in range(wl["random_splits"]):
        a = int(rng.integers(2, n // 2 + 1))
        left = set(rng.choice(n, size=a, replace=False).tolist())
        out.append((tuple(t for i, t in enumerate(taxa) if i in left), tuple(t for i, t in enumerate(taxa) if i not in left)))
    return out


def make_model(sim, name):
    return sim.GTR.JukesCantor(0.5) if name == "JC" else sim.GTR((0.1, 0.2, 0.3, 0.4), (1, 2, 3, 4, 5, 6))


# ------------------------------------------------------------------------------------------------
# clocks (nvidia-smi sampled DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed loops: an in-process NVML (nvidia_ml_py) thread polling
    the two cheap per-field queries every 100 ms, initialised before the warm-up.  Measured on the c2 step
    (scripts/step_jitter.py): this leaves the 86 ms steps within +3 ms, whereas a looping `nvidia-smi` child, a poll
    from the launching thread, or cudaMemGetInfo inside the step each produced sporadic 100-170 ms steps.
    Falls back to a looping nvidia-smi child when NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.handle, self.max_mhz = None, None, None
        self.sm, self.reasons, self.active = [], set(), False
        self.quit = threading.Event()

    def _poll(self):
        n = self.nvml
        masks = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self.quit.wait(0.1):
            if not self.active:
                continue
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:  # noqa: BLE001
                continue
            if self.active:
                self.sm.append(mhz)
                self.reasons.update(k for k, m in masks.items() if r & m)

    def start(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            # torch device `index` of this process -> NVML handle via the PCI bus id (CUDA_VISIBLE_DEVICES-proof)
            bus = getattr(torch.cuda.get_device_properties(self.index), "pci_bus_id", None)
            if isinstance(bus, int):
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.handle = h
                        break
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)  # the first query initialises (about 20 ms)
            self.nvml = pynvml
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] NVML sampling unavailable ({type(exc).__name__}: {exc}); falling back to nvidia-smi", file=sys.stderr)
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "250",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            if self.active:
                self.rows.append([x.strip() for x in line.split(",")])

    def reset(self):
        """Start of a timed loop: forget earlier samples and record from now on."""
        self.sm, self.reasons, self.rows = [], set(), []
        self.active = True

    def sample(self):
        """One extra sample from the calling thread (used when a timed loop is shorter than the polling period)."""
        if self.nvml is not None and not self.sm:
            try:
                self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
            except Exception:  # noqa: BLE001
                pass

    def close(self):
        self.quit.set()
        if self.proc is not None:
            self.proc.terminate()
            self.proc = None

    def stop(self):
        """End of a timed loop: summary of the samples taken since reset()."""
        self.active = False
        if self.nvml is not None:
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml thread, 100 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi -lms 250"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy / LAPACK restatement of the reference path) on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(workload, codes_np, tree, splits, budget_s=25.0):
    """Times the oracle on the host cores.  c2: pattern counting of ALL sites once, then one split per side
    size (reduced flattening + LAPACK split score, the reference's README loop); the whole-job rate is the
    size-weighted estimate 2035 / (t_count + sum_a n_a * t_a).  c3: pattern counting + pair-table
    subflattening + score of a bounded number of splits."""
    from oracle import splitp_oracle as O
    cores = os.cpu_count() or 1
    n = codes_np.shape[0]
    pos = {t: i for i, t in enumerate(tree.taxa)}
    by_size = {}
    for s in splits:
        by_size.setdefault(min(len(s[0]), len(s[1])), []).append(s)
    per_size, spent = {}, 0.0
    if workload == "c4":
        sub = codes_np[:, : min(codes_np.shape[1], 200_000)]
        t0 = time.perf_counter()
        d, usable = O.get_pattern_counts_wide(sub)
        t_count = time.perf_counter() - t0
        done, t2 = 0, time.perf_counter()
        while done < len(splits) and time.perf_counter() - t2 < budget_s:
            s = splits[(done * 7919) % len(splits)]
            ia = [pos[x] for x in s[0]]
            O.split_score(O.flattening_reduced_from_dict(d, ia, [t for t in range(n) if t not in ia]))
            done += 1
        t_split = (time.perf_counter() - t2) / max(done, 1)
        scale = codes_np.shape[1] / sub.shape[1]
        value = len(splits) / (scale * (t_count + len(splits) * t_split))
        return {"value": value, "unit": "split-scores/s", "cores": cores, "kind": "port",
                "sample": f"oracle port on the first {sub.shape[1]} of {codes_np.shape[1]} sites: byte-row pattern count ({t_count:.2f} s) + "
                          f"{done} 2|62 splits ({t_split:.3f} s each: dict-based reduced flattening + LAPACK score); both stages are "
                          f"linear in the number of distinct patterns (= sites at 64 taxa), so value = S / (scale * (t_count + S * t_split)) "
                          f"with scale = {scale:.0f}", "sites_per_s": sub.shape[1] / t_count}
    if workload == "c2" or n <= 31:
        t0 = time.perf_counter()
        keys, counts, usable = O.get_pattern_counts_arrays(codes_np)
        t_count = time.perf_counter() - t0
        vals = counts / float(usable)
    if workload == "c2":
        for a in sorted(by_size):
            s = by_size[a][len(by_size[a]) // 2]
            ia, ib = [pos[x] for x in s[0]], [pos[x] for x in s[1]]
            t1 = time.perf_counter()
            O.split_score(O.flattening_reduced(keys, vals, n, ia, ib))
            per_size[a] = time.perf_counter() - t1
            spent += per_size[a]
        total = t_count + sum(len(by_size[a]) * per_size[a] for a in per_size)
        value = len(splits) / total
        sample = (f"oracle port: pattern count of all {codes_np.shape[1]} sites ({t_count:.2f} s) + one split per side size "
                  f"{sorted(per_size)} (reduced flattening + LAPACK gesdd score: "
                  + ", ".join(f"{a}:{per_size[a]:.3f}s" for a in sorted(per_size))
                  + f"); value = {len(splits)} / (t_count + sum_a n_a t_a)")
    else:
        t1 = time.perf_counter()
        if n <= 31:
            tables = O.pair_tables(keys, vals, n)
            tot, how = vals.sum(), f"pattern count of all {codes_np.shape[1]} sites ({t_count:.2f} s) + numpy pair tables from the pattern table"
        else:  # pattern keys are uint64 in the oracle: above 31 taxa the tables come straight from the code matrix
            t_count = 0.0
            tables, _ = O.pair_tables_from_codes(codes_np)
            tot, how = 1.0, f"numpy pair tables straight from the {codes_np.shape[1]}-site code matrix"
        t_pairs = time.perf_counter() - t1
        done, t2 = 0, time.perf_counter()
        while done < len(splits) and time.perf_counter() - t2 < budget_s:
            s = splits[(done * 7919) % len(splits)]
            O.split_score(O.subflattening_from_tables(tables, tot, [pos[x] for x in s[0]], [pos[x] for x in s[1]]))
            done += 1
        t_split = (time.perf_counter() - t2) / max(done, 1)
        value = len(splits) / (t_count + t_pairs + len(splits) * t_split)
        sample = (f"oracle port: {how} ({t_pairs:.2f} s) + {done} splits ({t_split * 1e3:.3f} ms each, subflattening from "
                  f"tables + LAPACK score); value = {len(splits)} / (t_count + t_pairs + S * t_split)")
        t_count = max(t_count + t_pairs, 1e-9)
    return {"value": value, "unit": "split-scores/s", "cores": cores, "kind": "port", "sample": sample,
            "sites_per_s": codes_np.shape[1] / t_count}


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries write there too (NCCL prints "NCCL version ..." when the
    environment sets NCCL_DEBUG), so file descriptor 1 points at stderr for the whole run and emit() restores it."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(line, flush=True)


def main():
    args = parse()
    quiet_stdout()
    wl = dict(WORKLOADS[args.workload])
    if args.sites:
        wl["sites"] = args.sites
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        import torch
        from splitp_b200 import simulation, splits as splits_mod, trees
        tree = trees.balanced_tree(wl["n"], wl["bl"])
        # same generator as the GPU arm when a device is present (identical alignment), CPU generator otherwise
        gen_dev = "cuda" if torch.cuda.is_available() else "cpu"
        codes = simulation.simulate_codes(tree, make_model(simulation, wl["model"]), wl["sites"], wl["seed"], device=gen_dev).cpu().numpy()
        splits = make_splits(splits_mod, tree, wl)[: args.max_splits]
        vals = []
        for _ in range(args.warmup):
            cpu_reference_sample(args.workload, codes, tree, splits, budget_s=3.0)
        t0 = time.perf_counter()
        for _ in range(max(args.steps, 1)):
            cb = cpu_reference_sample(args.workload, codes, tree, splits, budget_s=10.0)
            vals.append(cb["value"])
        dt = (time.perf_counter() - t0) / max(args.steps, 1)
        v = float(np.mean(vals))
        cb["value"] = v
        emit(json.dumps({"impl": "reference", "metric": "split_scores_per_sec", "value": v, "unit": "split-scores/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": wl["desc"], "sites": wl["sites"], "splits": len(splits)},
                          "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "split-scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist

    import splitp_b200 as sp
    from splitp_b200 import distributed as spd
    eng = sp.engine
    rank, local, world = spd.init_from_env()
    dev = eng.device()

    tree = sp.trees.balanced_tree(wl["n"], wl["bl"])
    n, N = wl["n"], wl["sites"]
    codes_full = sp.simulation.simulate_codes(tree, make_model(sp.simulation, wl["model"]), N, wl["seed"])  # same on every rank
    sb, se = spd.shard_range(N, rank, world, 32)
    codes_dev = codes_full[:, sb:se].contiguous()          # this rank's site shard, resident in HBM
    codes_pin = codes_dev.cpu().pin_memory()               # e2e: host buffer
    splits = make_splits(sp.splits, tree, wl)[: args.max_splits]
    S = len(splits)
    idx_all = [eng.split_positions(s, tree.taxa) for s in splits]
    idx_mine = spd.shard_strided(idx_all, rank, world)  # round-robin: balances the steeply size-dependent split cost
    reduce_fn = spd.make_reduce_fn() if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    prof = {"gram": [], "pairs": [], "count": []}
    state = {}
    if args.workload not in ("c2", "c4"):
        ma_np, mb_np = eng.masks_from_splits(idx_mine)
        ma = torch.from_numpy(ma_np.view(np.int64)).to(dev)
        mb = torch.from_numpy(mb_np.view(np.int64)).to(dev)

    def timed(key, fn, units=1):
        if not state.get("profile"):
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        prof[key].append((a, b, units))
        return r

    def step(codes, probe=None):
        """One pass of the hot path over this rank's shard.  Returns the scores of ALL splits (device).
        probe() (clock sample) is called once, from this thread, at a point where the GPU has queued work."""
        calls = [0]

        def gram_hook(f, nb):
            r = timed("gram", f, nb)
            calls[0] += 1
            if probe is not None and calls[0] == 8:
                probe()
            return r

        if args.workload == "c2":
            aln = eng.pack(codes, want_planes=False)
            table = timed("count", lambda: eng.count_patterns(aln, reduce_fn=reduce_fn))
            scorer = state.get("scorer")
            if scorer is None:
                scorer = state["scorer"] = eng.CountScorer(table)
            scorer.table = table
            out = scorer.score_many(idx_mine, big_hook=gram_hook if (state.get("profile") or probe is not None) else None)
            if probe is not None and calls[0] < 8:
                probe()
            scorer.check_hi()
        elif args.workload == "c4":
            wide, valid, n_, N_ = eng.pack_wide(codes)
            table = timed("count", lambda: spd.count_patterns_wide_sharded(wide, valid, n_, N_, rank, world, local=True))
            out = eng.thin_split_scores(table, [min((ia, ib), key=len) for ia, ib in idx_mine])
            if probe is not None:
                probe()
        else:
            aln = eng.pack(codes, want_sm=False)
            raw = timed("pairs", lambda: eng.pair_raw(aln))
            if world > 1:
                dist.all_reduce(raw)
            pt = eng.pair_finalize(raw, n, float(N))
            out = eng.subflatten_scores(pt, ma, mb)
            if probe is not None:
                probe()
        return spd.gather_strided(out, S, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()  # NVML initialisation happens here, before the warm-up, so that it cannot disturb a timed step

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        scores = step(codes_dev)
    barrier()

    # ---- timed: device-resident inputs ----
    sampler.reset()
    launches0 = int(eng.lib.spb_launch_count())
    state["profile"] = True
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations (outside the per-step event pair)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        scores = step(codes_dev)
        b.record()
        evs.append((a, b))
    sampler.sample()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    state["profile"] = False
    launches = int(eng.lib.spb_launch_count()) - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = S / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the public call, H2D + D2H inside the timed region ----
    for _ in range(2):
        step(codes_pin.to(dev, non_blocking=True)).cpu()
    barrier()
    sampler.reset()
    evs2 = []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        host_scores = step(codes_pin.to(dev, non_blocking=True)).cpu()
        b.record()
        evs2.append((a, b))
    sampler.sample()
    barrier()
    clocks2 = sampler.stop()
    sampler.close()
    e2e_steps = [a.elapsed_time(b) for a, b in evs2]
    t2 = torch.tensor([sum(e2e_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item()) / args.steps
    e2e = {"value": S / (e2e_ms * 1e-3), "unit": "split-scores/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(codes_pin.numel()) * world, "d2h_bytes_per_step": S * 8,
           "ms_steps_rank0": [round(x, 3) for x in e2e_steps], "clocks": clocks2,
           "call": {"c2": "engine.pack + count_patterns + CountScorer.score_many",
                    "c4": "engine.pack_wide + count_patterns_wide + thin_split_scores"}.get(
                        args.workload, "engine.pack + pair_raw/pair_finalize + subflatten_scores")}

    # ---- roofline of the dominant kernel, measured live with CUDA events on the launch stream ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    roof = None
    if args.workload == "c2" and prof["gram"]:
        # one launch computes `units` Gram matrices (batched launch): time per matrix = sum of launch times / matrices
        ms = float(sum(a.elapsed_time(b) for a, b, _ in prof["gram"]) / sum(u for _, _, u in prof["gram"]))
        flops = 2.0 * 4096 ** 3  # 2 R^2 C per 6|6 split (GEMM convention)
        tiles_done, tiles_all = 272, 512
        bf16 = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
        which = "measured (MEASURED_PEAKS.json bf16 sustained x2: kind::i8 has K=32 per instruction vs 16 for bf16)" if peaks else \
            "fallback 1.59 PFLOP/s x2"
        roof = {"kernel": "gram_u8_umma_kernel<256> (tcgen05.mma kind::i8, 4096x4096x4096 per 6|6 split)", "bound": "tensor",
                "achieved": flops / (ms * 1e-3) / 1e12, "peak": 2 * bf16, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / (2 * bf16),
                "traffic": 80.8e6, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum = 17.1 MB + 63.7 MB per 4096^2 matrix, from the "
                "ncu --set full capture of a 16-matrix launch (profiles/r1_ncu_c2_gram_symv_u8_final.txt, launch 24); algorithmic "
                "16.8 MB S0 read + 67.1 MB int32 G0 written",
                "peak_source": which, "ms_per_matrix": ms, "launches_timed": len(prof["gram"]), "matrices_per_launch": 64,
                "executed_frac_of_algorithmic": tiles_done / tiles_all,
                "achieved_executed": flops * tiles_done / tiles_all / (ms * 1e-3) / 1e12,
                "frac_executed": flops * tiles_done / tiles_all / (ms * 1e-3) / 1e12 / (2 * bf16),
                "note": "algorithmic flops = full 2*R^2*C (SURVEY 8d) while the kernel computes only the 272 of 512 tiles touching the "
                        "upper triangle and mirrors the rest, so `frac` can exceed 1; `frac_executed` counts the executed MMAs only"}
    elif args.workload == "c4" and prof["count"]:
        ms = float(np.mean([a.elapsed_time(b) for a, b, _ in prof["count"]]))
        nbytes = (se - sb) * 16.0 + (se - sb) / 8.0  # one 128-bit key per site (= N n / 4 at 64 taxa) + validity mask
        hbm = peaks.get("hbm_gbs", 6650.0)
        roof = {"kernel": "count_wide_kernel (+ table setup, compaction and merge when sharded)", "bound": "hbm",
                "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm, "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s", "launch_ms": ms,
                "note": "random 16-byte CAS + 4-byte atomic per distinct pattern: DRAM-transaction bound, not streaming bound"}
    elif args.workload != "c2" and prof["pairs"]:
        ms = float(np.mean([a.elapsed_time(b) for a, b, _ in prof["pairs"]]))
        nbytes = (se - sb) * n / 4.0 + (se - sb) / 8.0  # N n / 4 (+ N / 8 validity mask)
        hbm = peaks.get("hbm_gbs", 6650.0)
        roof = {"kernel": "pair_kernel (bit-plane AND + POPC pair statistics)", "bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9,
                "peak": hbm, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm, "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s", "launch_ms": ms,
                "note": "pair_kernel is the only HBM-streaming kernel of this workload (POPC-issue bound, see DESIGN.md) and a small "
                        "share of the step; the step is dominated by the batched score kernel (subflatten_score_warp_kernel up to 21 taxa: "
                        "one warp per split, tridiagonalisation + bisection; subflatten_score_kernel above: shared-memory Jacobi; "
                        "instruction-issue bound, no HBM roofline): its throughput is `value`"}
    count_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in prof["count"]])) if prof["count"] else None

    if rank == 0:
        out = {"metric": "split_scores_per_sec", "value": value, "unit": "split-scores/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": {"c2": "u8", "c4": "u128"}.get(args.workload, "f64"), "data": "synthetic",
               "config": {"workload": wl["desc"], "taxa": n, "sites": N, "splits": S, "l2": "flushed between timed steps (256 MB fill)",
                          "parallelism": f"sites+splits sharded x{world}"},
               "sites_per_sec": (N / (count_ms * 1e-3)) if count_ms else N / (ms_per_step * 1e-3),
               "count_stage_ms": count_ms, "wall_s_timed_region": t_wall, "ms_steps_rank0": [round(x, 3) for x in step_ms],
               "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof}
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_reference_sample(args.workload, codes_full.cpu().numpy(), tree, splits, budget_s=15.0)
        elif world > 1:
            out["cpu_baseline"] = None
        # cheap sanity: the scores are finite and the tree's true splits rank first among their size
        sc = scores.cpu().numpy()
        out["checks"] = {"finite": bool(np.isfinite(sc).all()), "host_equals_device": bool(np.array_equal(sc, host_scores.numpy()))}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
