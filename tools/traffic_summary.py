"""ncu csv (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum) -> per-kernel DRAM traffic json.
Usage: python tools/traffic_summary.py gpurun_out/traffic.csv > profiles/<round>_traffic.json"""
import collections, csv, json, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: {"launches": set(), "dram_bytes_total": 0.0, "time_us": 0.0})
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void athtd::", "").replace("athtd::", "").replace("void ", "")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]; m = row["Metric Name"]
    a = agg[name]; a["launches"].add(row["ID"])
    if m.startswith("dram__bytes"): a["dram_bytes_total"] += v * scale.get(u, 1.0)
    elif m.startswith("gpu__time"): a["time_us"] += v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
out = {k: {"launches": len(v["launches"]), "dram_bytes_total": v["dram_bytes_total"], "time_us": v["time_us"],
           "dram_gbs": round(v["dram_bytes_total"] / max(v["time_us"], 1e-9) / 1e3, 1)} for k, v in agg.items()}
out["_note"] = "one batch-32 bf16 forward (6 s segments), ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, summed per kernel"
print(json.dumps(out, indent=1))
