"""Summarise `ncu --page source --csv` output: runs of SASS instructions with the same executed count (basic blocks), their share of the
kernel's warp instructions and stall samples, and an opcode histogram.   usage: ncu_source_blocks.py file.csv kernel-substring [min_share]"""
import csv, sys, collections
path, pat = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
rows, on, hdr = [], False, None
for r in csv.reader(open(path)):
    if r and r[0] == "Kernel Name":
        if on: break
        on = pat in r[1]; hdr = None; continue
    if not on: continue
    if hdr is None: hdr = r; continue
    rows.append(r)
ix = {h: i for i, h in enumerate(hdr)}
ie, isamp, isrc = ix["Instructions Executed"], ix["# Samples"], ix["Source"]
tot = sum(int(r[ie]) for r in rows); tots = sum(int(r[isamp]) for r in rows)
print(f"{len(rows)} SASS lines, {tot} warp instructions, {tots} samples")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
blocks, cur = [], None
for n, r in enumerate(rows):
    e = int(r[ie])
    if cur is None or e != cur["e"]:
        cur = {"e": e, "start": n, "rows": []}; blocks.append(cur)
    cur["rows"].append(r)
for b in blocks:
    w = b["e"] * len(b["rows"]); s = sum(int(r[isamp]) for r in b["rows"])
    if w / tot < min_share and s / max(tots, 1) < min_share: continue
    ops = collections.Counter(r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1] for r in b["rows"])
    st = collections.Counter()
    for r in b["rows"]:
        for c in stall_cols: st[c[6:]] += int(r[ix[c]])
    top = " ".join(f"{k}:{v}" for k, v in st.most_common(4))
    print(f"line {b['start']:5d} n={len(b['rows']):4d} exec={b['e']:9d} inst {100*w/tot:5.1f}% samples {100*s/max(tots,1):5.1f}% | {top}")
    print("      " + " ".join(f"{k}:{v}" for k, v in ops.most_common(14)))
