"""Executed warp instructions of an .ncu-rep grouped by the SASS function / inlined source file (dev tool)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr = "", None
agg = collections.Counter(); smp = collections.Counter()
seen = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    key = (fname, int(r[0]))
    if key in seen: continue
    seen.add(key)
    try: ie, ss = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
    except ValueError: continue
    agg[key] += ie; smp[key] += ss
ti = sum(agg.values()); ts = sum(smp.values())
# bucket by file and 25-line ranges
b = collections.Counter(); bs = collections.Counter()
for (f, l), v in agg.items():
    b[(f, l // 20 * 20)] += v; bs[(f, l // 20 * 20)] += smp[(f, l)]
print("total warp instructions", ti)
for k, v in sorted(b.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{100*v/ti:6.2f}%i {100*bs[k]/ts:6.2f}%s  {k[0]}:{k[1]}-{k[1]+19}")
