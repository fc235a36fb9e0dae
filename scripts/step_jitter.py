"""Diagnostics: per-step and per-phase times of the c2 step over many repetitions (host wall + CUDA events)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splitp_b200 as sp
eng = sp.engine
n, N = 12, 1_000_000
tree = sp.trees.balanced_tree(n, 0.05)
codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), N, seed=2)
idx = [eng.split_positions(s, tree.taxa) for s in sp.all_splits(tree)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=codes.device)
scorer = None
use_i32 = (sys.argv[1] != "f64") if len(sys.argv) > 1 else True
e2e = len(sys.argv) > 2 and sys.argv[2] == "e2e"
codes_pin = codes.cpu().pin_memory()
orig_score_i32 = eng.CountScorer._score_i32
phase = {}

def wrap(name, fn):
    def inner(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        phase[name] = phase.get(name, 0.0) + time.perf_counter() - t0
        return r
    return inner

eng.CountScorer._score_i32 = wrap("score_i32(host)", eng.CountScorer._score_i32)
eng.CountScorer._gram_batch_i32 = wrap("gram_batch_i32(host)", eng.CountScorer._gram_batch_i32)
eng.CountScorer._gram_batch = wrap("gram_batch(host)", eng.CountScorer._gram_batch)
eng.score_gram = wrap("score_gram(host)", eng.score_gram)
eng.count_patterns = wrap("count(host)", eng.count_patterns)
eng.CountScorer._plan = wrap("plan", eng.CountScorer._plan)
eng.CountScorer._buffers = wrap("buffers", eng.CountScorer._buffers)
eng.CountScorer._buffers_i32 = wrap("buffers_i32", eng.CountScorer._buffers_i32)
torch.cuda.mem_get_info = wrap("mem_get_info", torch.cuda.mem_get_info)
_orig_tensor = torch.tensor
torch.tensor = wrap("torch.tensor", _orig_tensor)

import gc
import threading
if len(sys.argv) > 4 and sys.argv[4].startswith("nvml"):
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)
    if sys.argv[4] == "nvml_thread":
        def _poll():
            while True:
                t0 = time.perf_counter()
                pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(_h)
                dt = time.perf_counter() - t0
                phase["nvml_max_ms"] = max(phase.get("nvml_max_ms", 0.0), dt)
                time.sleep(0.1)
        threading.Thread(target=_poll, daemon=True).start()
if len(sys.argv) > 3 and sys.argv[3] == "nogc":
    gc.disable()
gc.callbacks.append(lambda ph, info: phase.__setitem__(f"gc{info['generation']}", phase.get(f"gc{info['generation']}", 0) + (1e-3 if ph == "start" else 0)))
for it in range(16):
    flush.fill_(1)
    phase.clear()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    tA = time.perf_counter()
    src = codes_pin.to(codes.device, non_blocking=True) if e2e else codes
    phase["h2d(host)"] = time.perf_counter() - tA
    tA = time.perf_counter()
    aln = eng.pack(src, want_planes=False)
    phase["pack(host)"] = time.perf_counter() - tA
    table = eng.count_patterns(aln)
    if scorer is None:
        scorer = eng.CountScorer(table)
        scorer.int32_gram = use_i32
    scorer.table = table
    tA = time.perf_counter()
    out = scorer.score_many(idx)
    phase["score_many(host)"] = time.perf_counter() - tA
    if e2e:
        tA = time.perf_counter()
        host = out.cpu()
        phase["d2h(host)"] = time.perf_counter() - tA
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"step {it:2d}: gpu {a.elapsed_time(b):8.2f} ms  wall {wall:8.2f} ms  " + "  ".join(f"{k} {v * 1e3:7.2f}" for k, v in sorted(phase.items())),
          f" reserved {torch.cuda.memory_reserved() >> 20} MB")
