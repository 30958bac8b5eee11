"""Executed warp instructions and stall samples per source line of one source file, from an .ncu-rep captured with
--import-source on (dev tool):  python tools/ncu_lines.py REP FILE FIRST_LINE LAST_LINE"""
import csv, io, subprocess, collections, sys
rep=sys.argv[1]; fsel=sys.argv[2]; lo=int(sys.argv[3]); hi=int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr = "", None
agg = collections.Counter(); smp = collections.Counter(); seen=set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    key=(fname,int(r[0]))
    if key in seen: continue
    seen.add(key)
    try: ie, ss = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
    except ValueError: continue
    agg[key]+=ie; smp[key]+=ss
ti=sum(agg.values()); ts=sum(smp.values())
import os
src=open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'hans_b200', 'csrc', fsel)).read().split('\n')
tot=0
for (f,l),v in sorted(agg.items()):
    if f==fsel and lo<=l<=hi:
        tot+=v
        if v/ti>0.0008 or smp[(f,l)]/ts>0.002:
            print(f"{l:5d} {100*v/ti:5.2f}%i {100*smp[(f,l)]/ts:5.2f}%s  {src[l-1].strip()[:140]}")
print("sum", 100*tot/ti)
