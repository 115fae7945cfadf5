#!/usr/bin/env python
"""Markdown table from `ncu --set full` reports: one row per distinct kernel (first launch of each).
usage: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/rNN_ncu_summary.md"""
import csv, io, re, subprocess, sys
COLS = [("time us", "gpu__time_duration.sum"), ("dram rd MB", "dram__bytes_read.sum"), ("dram wr MB", "dram__bytes_write.sum"),
        ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("fp64 %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        ("lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), ("warps/SM", "sm__warps_active.avg.per_cycle_active"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("smem KB/CTA", "launch__shared_mem_per_block_dynamic"), ("inst M", "smsp__inst_executed.sum")]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
seen, lines = set(), []
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"\((int|bool|unsigned int)\)", "", r[ix["Kernel Name"]].split("(const")[0].split("(float")[0]).replace("void ", "").replace("ctd::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        name = re.sub(r"\(.*$", "", name)
        if name in seen:
            continue
        seen.add(name)
        vals = []
        for label, key in COLS:
            if key not in ix or r[ix[key]] in ("", "n/a"):
                vals.append("")
                continue
            v = float(r[ix[key]].replace(",", ""))
            u = units[ix[key]]
            if label.startswith("dram rd") or label.startswith("dram wr") or label.startswith("time"):
                v *= SCALE.get(u, 1.0)
            if label == "smem KB/CTA":
                v *= {"byte": 1e-3, "Kbyte": 1.0}.get(u, 1e-3) if "block" in u or "byte" in u else 1e-3
            if label == "inst M":
                v *= 1e-6
            vals.append(("%.4g" % v))
        lines.append("| " + name[:64] + " | " + " | ".join(vals) + " |")
print("| kernel | " + " | ".join(l for l, _ in COLS) + " |")
print("|" + "---|" * (len(COLS) + 1))
print("\n".join(lines))
