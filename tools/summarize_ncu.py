#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/r01_launches.csv profiles/r01_launches_summary.md [--last-step N]
  python tools/summarize_ncu.py full gpurun_out/r01_block.ncu-rep profiles/r01_block_ncu.csv

`launches`: per-kernel totals of the `--metrics gpu__time_duration.sum` pass (cold-cache, serialised: compare SHARES).
`full`: key metrics of every kernel in an `ncu --set full` report (needs ncu on PATH to read the .ncu-rep).
"""
from __future__ import annotations

import csv
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "")
    return name[:90]


def launches(src: str, dst: str, ours_only_from: str | None = None) -> None:
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((int(r["ID"]), short(r["Kernel Name"]), v * scale, r["Grid Size"], r["Block Size"]))
    ours = [r for r in rows if r[1].startswith(("gf::", "gf_"))]
    agg: "OrderedDict[str, list]" = OrderedDict()
    for _, name, ms, _, _ in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    tot_ours = sum(r[2] for r in ours)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`: every launch serialised with cold caches; "
                "SHARES are comparable with bench.py's CUDA-event breakdown, absolute times are not.\n\n")
        f.write(f"launches: {len(rows)} total, {len(ours)} from libgoalforce_b200.so; device time {tot:.1f} ms total, "
                f"{tot_ours:.1f} ms in our kernels\n\n")
        f.write("| kernel | launches | total ms | share of all | share of ours |\n|---|---:|---:|---:|---:|\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            so = f"{100 * ms / tot_ours:.2f}%" if name.startswith(("gf::", "gf_")) else "-"
            f.write(f"| `{name}` | {n} | {ms:.3f} | {100 * ms / tot:.2f}% | {so} |\n")
    print(f"wrote {dst}: {len(rows)} launches, ours {len(ours)}")


def full(src: str, dst: str) -> None:
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in idx]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel"] + [f"{k} [{units[idx[k]]}]" for k in keys])
        for r in rows[2:]:
            w.writerow([r[idx["ID"]], short(r[idx["Kernel Name"]])] + [r[idx[k]] for k in keys])
    print(f"wrote {dst}: {len(rows) - 2} kernels, {len(keys)} metrics")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "launches":
        launches(src, dst)
    elif mode == "full":
        full(src, dst)
    else:
        raise SystemExit(__doc__)
