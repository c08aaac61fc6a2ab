#!/usr/bin/env python
"""Per-instruction stall summary of one kernel in an .ncu-rep (needs `ncu` on PATH; run on the authoring box).
   python tools/ncu_stalls.py gpurun_out/attn_v2.ncu-rep regex:gf_attn [top_n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
for key in ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct"):
    if key in h:
        print(f"{key:90s} {rows[2][h.index(key)]} {rows[1][h.index(key)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
isrc, ins, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[2:] if len(r) > max(ins, iex) and r[ins].isdigit()]
tot = sum(int(r[ins]) for r in data)
print("total samples", tot, "sass lines", len(data))
agg = {}
for r in data:
    for i, c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[i] or 0)
print("  ".join(f"{c[6:]}={100 * v / tot:.1f}%" for c, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
ops = {}
for r in data:
    t = r[isrc].split()
    o = t[1] if t[0].startswith("@") else t[0]
    e = ops.setdefault(o.split(".")[0], [0, 0])
    e[0] += int(r[ins]); e[1] += int(r[iex])
print("opcode: samples% executed")
for o, (s, e) in sorted(ops.items(), key=lambda x: -x[1][0])[:22]:
    print(f"  {o:14s} {100 * s / tot:5.1f}% {e}")
top = sorted(range(len(data)), key=lambda i: -int(data[i][ins]))[:top_n]
for i in sorted(top):
    r = data[i]
    st = sorted([(int(r[j] or 0), c[6:]) for j, c in stall_cols], reverse=True)[:2]
    print(i, r[isrc][:72].ljust(72), r[ins].rjust(7), r[iex].rjust(10), st)
