"""One-screen summary of a bench.py JSON line.  usage: bench_brief.py file.json"""
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as exc:  # noqa: BLE001
    print("no bench line:", exc)
    sys.exit(0)
r = d.get("roofline") or {}
print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.1f}  launches {d.get('gpu_launches')}  clocks {d['clocks'].get('sm_mhz')}")
print("phase_ms", d.get("phase_ms"))
print("roofline", {k: r.get(k) for k in ("achieved", "peak", "frac", "ms_per_matrix", "share_of_step", "frac_of_sustained")})
print("checks", d.get("checks"))
for k, w in (d.get("workloads") or {}).items():
    print(" ", k, {x: w.get(x) for x in ("value", "ms_per_step", "error")}, "e2e", (w.get("e2e") or {}).get("value"))
if d.get("partition"):
    print("partition", d["partition"].get("class_cost_us"), d["partition"].get("splits_per_rank"), d["partition"].get("refit_rank_times_us"))
