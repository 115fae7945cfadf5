#!/usr/bin/env python
"""One-table summary of an .ncu-rep (`ncu --set full` capture) for profiles/: duration, DRAM bytes, pipe
utilisation, occupancy per kernel.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu(mufu) %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__waves_per_multiprocessor", "waves/SM")]
print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    vals = []
    for key, _ in cols:
        if key in ix:
            v, u = r[ix[key]], units[ix[key]]
            try:
                v = "%.4g" % float(v)
            except ValueError:
                pass
            vals.append(v + (" " + u if u and u not in ("%",) else ""))
        else:
            vals.append("n/a")
    print("| " + name + " | " + " | ".join(vals) + " |")
