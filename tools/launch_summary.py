"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares, top launches)."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0; big = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
    short = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void athtd::", "").replace("athtd::", "")
    agg[short][0] += 1; agg[short][1] += v; tot += v
    big.append((v, short, row.get("Grid Size", "")))
print(f"total {tot:.1f} us over {len(big)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} n={n:4d} {t:10.1f} us {100 * t / tot:5.1f}%")
print("top launches:")
for v, s, g in sorted(big, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 15]:
    print(f"  {v:9.1f} us {s} {g}")
