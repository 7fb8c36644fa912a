"""Summarise `ncu --page source --csv` output: top source lines / SASS by warp-stall samples per kernel."""
import csv, sys, collections, re
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = open(path, errors="replace").read()
# the CSV holds several kernels back to back, each starting with a header row containing "Source"
blocks = re.split(r'(?m)^(?="Kernel Name")', txt)
rows = list(csv.reader(txt.splitlines()))
hdr = None; kernel = None
acc = collections.defaultdict(lambda: collections.defaultdict(float))
src_of = {}
for r in rows:
    if not r: continue
    if "Source" in r and ("# Samples" in " ".join(r) or "Warp Stall Sampling (All Samples)" in r or "Sampling Data (All)" in r):
        hdr = {h: i for i, h in enumerate(r)}; continue
    if len(r) == 1 or (hdr is None):
        if r and r[0].startswith("Kernel"): kernel = r[0]
        continue
    try:
        key = None
        for cand in ("Warp Stall Sampling (All Samples)", "# Samples", "Sampling Data (All)"):
            if cand in hdr: key = cand; break
        if key is None: continue
        v = float(r[hdr[key]] or 0)
    except Exception:
        continue
    src = r[hdr["Source"]][:150]
    acc[kernel][src] += v
for k, d in acc.items():
    tot = sum(d.values()) or 1
    print("=====", k, "total samples", tot)
    for src, v in sorted(d.items(), key=lambda kv: -kv[1])[:topn]:
        print(f"{100*v/tot:6.2f}%  {src}")
