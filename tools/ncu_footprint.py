"""Static size (KB of SASS actually executed) and dynamic share per source region of an .ncu-rep (dev tool).
Answers: which code makes up the instruction-cache footprint of the hot path?"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = None, []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ie = hdr.index("Instructions Executed")
ex = [int(r[ie] or 0) for r in data]
# map SASS address -> source line via the cuda,sass view (source rows followed by their sass rows are not
# available in csv), so use nvdisasm-style line info from the second view: per source line, instruction counts
out2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows2 = list(csv.reader(io.StringIO(out2)))
fname, h2 = "", None
stat = collections.Counter(); dyn = collections.Counter(); seen=set()
for r in rows2:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h2 = r; continue
    if h2 is None or not r[0].isdigit(): continue
    key=(fname,int(r[0]))
    if key in seen: continue
    seen.add(key)
    try:
        d = int(r[h2.index("Instructions Executed")])
    except ValueError: continue
    # static instruction count of a line is not in the csv; approximate by executed/avg? -> use 'Address' column count
    dyn[(fname, int(r[0]) // bucket * bucket)] += d
tot = sum(dyn.values())
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
print(f"static executed: {sum(1 for e in ex if e>0)*16/1024:.1f} KB of {len(ex)*16/1024:.1f} KB")
for k, v in sorted(dyn.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{100*v/tot:6.2f}%  {k[0]}:{k[1]}-{k[1]+bucket-1}")
