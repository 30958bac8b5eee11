"""Instruction-cache footprint of a profiled kernel by source region (dev tool): static SASS bytes that were
executed, from the .ncu-rep's per-instruction counts joined with nvdisasm line info of the built library.
    python tools/ncu_static.py <rep> <mangled kernel name substring> [bucket lines]"""
import re, collections, csv, subprocess, io, sys, os, tempfile
rep, kname = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = None, []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ie = hdr.index('Instructions Executed')
base = int(data[0][0], 16)
off_ex = {int(r[0], 16) - base: int(r[ie] or 0) for r in data}
d = tempfile.mkdtemp()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.system(f"cd {d} && cuobjdump -xelf all {root}/hans_b200/libhans_b200.so >/dev/null 2>&1 && nvdisasm --print-line-info *.cubin > dis.txt 2>/dev/null")
fn = None; cur = None
per = collections.defaultdict(lambda: [0, 0, 0])
for line in open(f"{d}/dis.txt"):
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m: fn = m.group(1); cur = None; continue
    if fn is None or kname not in fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s', line)
    if m and cur:
        e = off_ex.get(int(m.group(1), 16), 0)
        key = (cur[0], cur[1] // bucket * bucket)
        per[key][0] += 1; per[key][1] += (e > 0); per[key][2] += e
tot = sum(v[2] for v in per.values()) or 1
print(f"{'region':32s} staticKB  execKB   dyn%")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k[0]:24s}{k[1]:6d}  {v[0]*16/1024:7.1f} {v[1]*16/1024:7.1f} {100*v[2]/tot:6.2f}")
print("total executed KB:", sum(v[1] for v in per.values()) * 16 / 1024)
