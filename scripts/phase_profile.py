"""Per-phase CUDA-event timing of the c2 workload (diagnostics; not a bench number)."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import splitp_b200 as sp
eng = sp.engine
n, N = 12, int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tree = sp.trees.balanced_tree(n, 0.05)
codes = sp.simulation.simulate_codes(tree, sp.simulation.GTR.JukesCantor(0.5), N, seed=2)


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for _ in range(reps):
        r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r


t, aln = timeit(lambda: eng.pack(codes, want_planes=False)); print(f"pack            {t:9.3f} ms")
t, tab = timeit(lambda: eng.count_patterns(aln)); print(f"count+compact   {t:9.3f} ms  P={tab.num}")
scorer = eng.CountScorer(tab)
splits = list(sp.all_splits(tree))
idx = [eng.split_positions(s, tree.taxa) for s in splits]
for a in (2, 3, 4, 5, 6):
    mem = [i for i in idx if len(i[0]) == a][:32]
    R, Cc = 4 ** a, 4 ** (n - a)
    layout, rows_pad, pitch = scorer.geometry(R, Cc)
    s0, G, ws = scorer._buffers(layout, rows_pad, pitch, len(mem)); s0 = s0[0]
    plan = scorer._plan(*mem[0], False)
    spl, rr, rc = plan[0], plan[1], plan[2]
    t1, _ = timeit(lambda: eng.call("spb_flatten_u8", eng._p(tab.keys), eng._p(tab.counts), tab.num, C.byref(spl), None, None, eng._p(s0), rows_pad, pitch, layout, 1, eng._p(scorer.hi_rc), eng._p(scorer.hi_val), eng._p(scorer.hi_num), scorer.hi_cap, eng._st()))
    t2, _ = timeit(lambda: eng.call("spb_gram_u8", eng._p(s0), rows_pad, pitch, layout, eng._p(G[0]), eng._p(ws), eng._st()))
    t3, _ = timeit(lambda: eng.call("spb_gram_hi_correction", eng._p(s0), rows_pad, pitch, layout, eng._p(scorer.hi_rc), eng._p(scorer.hi_val), eng._p(scorer.hi_num), scorer.hi_cap, eng._p(G[0]), eng._st()))
    t4, _ = timeit(lambda: eng.call("spb_flatten_u8_clear", eng._p(tab.keys), tab.num, C.byref(spl), None, None, eng._p(s0), rows_pad, pitch, layout, eng._st()))
    s0b = scorer._s0[(layout, rows_pad, pitch)]
    nbm = min(16, len(mem))
    tb, _ = timeit(lambda: scorer._gram_batch([scorer._plan(ia, ib, False)[0] for ia, ib in mem[:nbm]], s0b, G[:nbm], ws, layout, rows_pad, pitch))
    for b0 in range(0, len(mem), 16):
        sub = mem[b0:b0 + 16]
        scorer._gram_batch([scorer._plan(ia, ib, False)[0] for ia, ib in sub], s0b, G[b0:b0 + len(sub)], ws, layout, rows_pad, pitch)
    t5, (sc, info) = timeit(lambda: eng.score_gram(G[:len(mem)], R, want_info=True), reps=2)
    t6, _ = timeit(lambda: eng.score_gram(G[:1], R), reps=2)
    extra = ""
    if info is not None and info.shape[1] == 8:
        inf = info.cpu().numpy()
        extra = f" cycles max {inf[:,4].max():.0f} res max {inf[:,2].max():.1e}"
    print(f"a={a} R={R:5d} nnz_hi={int(scorer.hi_num[0].item())}: flatten {t1:7.3f}  gram {t2:7.3f}  hi {t3:7.3f}  clear {t4:7.3f}  "
          f"gram-stage batch{nbm} {tb/nbm:7.3f}/split  score(batch {len(mem)}) {t5:8.3f} = {t5/len(mem):7.3f}/split, score(batch 1) {t6:8.3f} ms{extra}")
t0 = time.perf_counter()
out = scorer.score_many(idx); torch.cuda.synchronize()
t1 = time.perf_counter()
out = scorer.score_many(idx); torch.cuda.synchronize()
print(f"score_many all {len(idx)} splits: {time.perf_counter() - t1:.3f} s (first call {t1 - t0:.3f} s)")
