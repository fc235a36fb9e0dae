"""ncu launch list (csv of gpu__time_duration.sum) -> per-kernel markdown table.  usage: launch_summary.py in.csv 'title'"""
import collections, csv, re, sys
lines = open(sys.argv[1]).readlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines[start:]):
    if row["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else 'ncu launch list'}\n")
print(f"Total kernel time {T / 1e3:.1f} ms over {sum(cnt.values())} launches (cold-cache, serialised: compare SHARES).\n")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"| {k[:80]} | {cnt[k]} | {v / 1e3:.2f} | {100 * v / T:.1f}% | {v / cnt[k]:.2f} |")
