"""SASS evidence per kernel of libathtd.so: counts of the tensor-core / TMA / TMEM mnemonics (cuobjdump -sass).
Usage: python tools/sass_summary.py > profiles/<round>_sass_summary.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "audio-to-sheet-music_b200", "libathtd.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = {"UTCHMMA (tcgen05.mma)": r"\bUTCHMMA", "UTMALDG (TMA load)": r"\bUTMALDG", "LDTM (tcgen05.ld)": r"\bLDTM",
        "UTCBAR (tcgen05.commit)": r"\bUTCBAR", "SYNCS (mbarrier)": r"\bSYNCS", "HMMA (mma.sync bf16)": r"\bHMMA",
        "LDSM (ldmatrix)": r"\bLDSM", "MUFU": r"\bMUFU", "FFMA": r"\bFFMA", "LDG": r"\bLDG", "STG": r"\bSTG", "RED/ATOM": r"\b(RED|ATOM)"}
cur, counts, sizes = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); sizes[cur] = 0; continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        sizes[cur] += 1
        for k, p in pats.items():
            if re.search(p, line): counts[cur][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(so, root)}  (sm_100a); instruction counts per kernel")
for name, d in zip(counts, dem):
    short = re.sub(r"\(.*", "", d).replace("athtd::", "").replace("void ", "")
    c = counts[name]
    if sizes[name] < 50: continue
    tags = ", ".join(f"{k.split(' ')[0]}={v}" for k, v in c.items() if v)
    print(f"{short[:70]:70s} {sizes[name]:6d} instr | {tags}")
