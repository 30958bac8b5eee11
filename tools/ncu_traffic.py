"""profiles/r02_traffic.json entry from an `ncu --set full` capture of the walk kernel:
    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep C3:1024 profiles/r02_ncu_c3_lean.txt
dram__bytes_read.sum + dram__bytes_write.sum of the first captured launch (bench.py reads the file for roofline.traffic)."""
import csv, io, json, os, subprocess, sys

rep, key, src = sys.argv[1], sys.argv[2], sys.argv[3]
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
H, U, r = raw[0], raw[1], raw[2]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = 0.0
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    i = H.index(name)
    tot += float(r[i]) * scale[U[i]]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[key] = {"bytes_per_launch": tot, "kernel": r[H.index("Kernel Name")], "source": src,
          "duration_ms_under_ncu": float(r[H.index("gpu__time_duration.sum")]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[U[H.index("gpu__time_duration.sum")]]}
json.dump(d, open(path, "w"), indent=1)
print(key, d[key])
