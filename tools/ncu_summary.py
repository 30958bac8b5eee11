"""Summarise an .ncu-rep (dev tool, run where ncu is installed; no GPU needed):
key launch metrics, stall reasons per issued instruction, hot code footprint, hottest source lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [nlines] > profiles/rNN_xxx.txt
"""
import csv, io, subprocess, sys
import numpy as np

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 40


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
H, U = raw[0], raw[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
for r in raw[2:]:
    print("== launch", r[H.index("ID")], r[H.index("Kernel Name")])
    for k in KEYS:
        if k in H:
            i = H.index(k)
            print(f"  {k} [{U[i]}] = {r[i]}")
    st = [(float(r[i] or 0), h.split("issue_stalled_")[1].split("_per_issue")[0]) for i, h in enumerate(H)
          if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    print("  stall cycles per issued instruction: " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True) if v >= 0.02))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
hdr, ex = None, []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        ex.append(int(r[hdr.index("Instructions Executed")] or 0))
ex = np.array(ex)
if ex.size:
    tot = ex.sum()
    s = np.sort(ex)[::-1]
    cs = np.cumsum(s) / tot
    print(f"== code: {ex.size * 16 / 1024:.1f} KB static SASS, {(ex > 0).sum() * 16 / 1024:.1f} KB executed; "
          + ", ".join(f"{int(f * 100)}% of dynamic instructions in {(int(np.searchsorted(cs, f)) + 1) * 16 / 1024:.1f} KB" for f in (0.5, 0.9, 0.99)))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
fname, hdr, out = "", None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit():
        continue
    try:
        key = (fname, int(r[0]), r[1][:96])
        ie, ss = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    if key not in out:  # several launches: keep the first
        out[key] = (ie, ss)
ti = sum(v[0] for v in out.values()) or 1
ts = sum(v[1] for v in out.values()) or 1
print(f"== hottest source lines (% of instructions executed, % of stall samples)")
for key, v in sorted(out.items(), key=lambda kv: -kv[1][1])[:nlines]:
    print(f"{100 * v[0] / ti:6.2f}%i {100 * v[1] / ts:6.2f}%s  {key[0]}:{key[1]:<5} {key[2]}")
