"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    u = row["Metric Unit"]
    ns = v * 1e3 if u.startswith("us") else v * 1e6 if u.startswith("ms") else v * 1e9 if u in ("s", "second") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    agg[name][0] += 1
    agg[name][1] += ns
    tot += ns
print(f"total kernel time {tot / 1e6:.3f} ms over {sum(n for n, _ in agg.values())} launches")
print(f"{'ms':>10} {'share':>6} {'launches':>8} {'avg us':>9}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{t / 1e6:10.3f} {100 * t / tot:5.1f}% {n:8d} {t / n / 1e3:9.1f}  {k[:100]}")
