"""Summarise an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line (dev tool)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
fname = ""; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples")
    try: out.append((int(r[ie]), int(r[ss]), fname, int(r[0]), r[1][:100]))
    except ValueError: pass
ti = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print("total inst", ti, "samples", ts)
key = 1 if (len(sys.argv) > 3 and sys.argv[3] == "samples") else 0
for o in sorted(out, key=lambda o: -o[key])[:top]:
    print(f"{100*o[0]/ti:6.2f}%i {100*o[1]/ts:6.2f}%s  {o[2]}:{o[3]:<5} {o[4]}")
