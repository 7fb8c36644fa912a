"""Top SASS instructions by stall samples from `ncu --page source --csv --print-source sass` output."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, errors="replace")))
# find header rows (contain 'Source' and 'Address')
out = []
hdr = None; kernel = "?"
for r in rows:
    if not r: continue
    if "Address" in r and "Source" in r:
        hdr = {h: i for i, h in enumerate(r)}
        out.append(("HDR", list(r)))
        continue
    if hdr is None:
        kernel = " ".join(r)[:120]
        continue
    out.append(("ROW", r))
print("header:", [h for k, h in out if k == "HDR"][:1])
cols = None
data = []
for k, r in out:
    if k == "HDR":
        cols = {h: i for i, h in enumerate(r)}; continue
    data.append(r)
samp = None
for cand in ("# Samples", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)", "Samples"):
    if cand in cols: samp = cand; break
print("sample column:", samp, "n rows", len(data))
def f(x):
    try: return float(x)
    except: return 0.0
tot = sum(f(r[cols[samp]]) for r in data) or 1
data.sort(key=lambda r: -f(r[cols[samp]]))
stall_cols = [c for c in cols if c.startswith("stall_") or "Stall" in c][:0]
for r in data[:topn]:
    print(f"{100*f(r[cols[samp]])/tot:6.2f}%  {r[cols['Source']][:110]}")
