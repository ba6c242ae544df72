"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    key = re.sub(r"^.*aid::", "", name.split("(")[0])[:60] if "aid::" in name else name.split("(")[0][-60:]
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
print(f"launches {sum(v[0] for v in agg.values())}  total {tot:.1f} us")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k:62s} n={c:5d} total={t:10.1f}us avg={t / c:8.1f}us share={t / tot * 100:5.1f}%")
