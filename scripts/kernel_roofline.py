"""HBM roofline of the streaming kernels (pack, pattern count, pair statistics) on large alignments.
Algorithmic bytes: N*n/4 (2 bits per base) + N/8 (validity mask) for count / pairs; n*N input bytes for pack."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import splitp_b200 as sp
eng = sp.engine
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
HBM = peaks["hbm_gbs"]


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps, r


for n, N in ((12, 1_000_000), (12, 100_000_000), (20, 10_000_000), (20, 100_000_000), (31, 50_000_000)):
    tree = sp.trees.balanced_tree(n if n % 2 == 0 else n + 1, 0.05)
    model = sp.simulation.GTR((0.1, 0.2, 0.3, 0.4), (1, 2, 3, 4, 5, 6))
    codes = sp.simulation.simulate_codes(tree, model, N, seed=n)[:n].contiguous()
    t_pack, aln = timeit(lambda: eng.pack(codes))
    del codes
    alg = N * n / 4 + N / 8
    usable = torch.zeros(1, dtype=torch.int64, device="cuda")
    if n <= 12:
        table = torch.zeros(4 ** n, dtype=torch.int32, device="cuda")
        t_cnt, _ = timeit(lambda: eng.call("spb_count_direct", eng._p(aln.sm), eng._p(aln.valid), n, 0, N, eng._p(table), None, eng._p(usable), eng._st()))
        kind = "direct"
    else:
        cap = 1 << 27
        hk = torch.full((cap,), -1, dtype=torch.int64, device="cuda"); hc = torch.zeros(cap, dtype=torch.int32, device="cuda")
        ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
        t_cnt, _ = timeit(lambda: eng.call("spb_count_hash", eng._p(aln.sm), eng._p(aln.valid), n, 0, N, eng._p(hk), eng._p(hc), None, cap, eng._p(usable), eng._p(ovf), eng._st()), reps=3)
        kind = f"hash P={int((hc > 0).sum())} ovf={int(ovf.item())}"
        del hk, hc
    t_pair, _ = timeit(lambda: eng.pair_raw(aln))
    t_full, tab = timeit(lambda: eng.count_patterns(aln), reps=2)
    print(f"n={n} N={N}: pack {t_pack:8.3f} ms ({n * N / t_pack / 1e6:7.1f} GB/s in)  count[{kind}] {t_cnt:8.3f} ms "
          f"({alg / t_cnt / 1e6:7.1f} GB/s = {alg / t_cnt / 1e6 / HBM:.3f} of HBM)  pairs {t_pair:8.3f} ms ({alg / t_pair / 1e6:7.1f} GB/s = "
          f"{alg / t_pair / 1e6 / HBM:.3f})  count_patterns end-to-end {t_full:8.3f} ms P={tab.num}", flush=True)
    del aln, tab
    torch.cuda.empty_cache()
