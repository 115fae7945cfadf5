#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report: DRAM bytes (read + write) per launch for the kernels of
the bench chain, keyed by bench.py's op names.  usage: python tools/ncu_traffic.py rep.ncu-rep > profiles/traffic.json"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
import re

def op_name(kernel):
    """bench.py's op name for a kernel of the chain (None = not part of it)."""
    k = re.sub(r"\((int|bool|unsigned int)\)", "", kernel).replace(" ", "")
    if "lcn_tma_kernel" in k:
        return "lcn_fwd"
    if "photo_fwd_bwd_box9_tma" in k:
        return "sad_fwd_bwd"
    if "photo_fwd_box9_tma" in k:
        return "sad_fwd"
    if "photo_bwd_box9_tma" in k:
        return "sad_bwd"
    if "photo_fwd_census9" in k:
        return "census_sad_fwd"
    m = re.search(r"photo_bwd_census9<3,(\w+),", k)
    if m:
        return "census_sad_fwd_bwd" if m.group(1) in ("1", "true") else "census_sad_bwd"
    if "masked_sums_kernel" in k:
        return "masked_sums"
    return None

out = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = op_name(name)
    if key is None or key in out:
        continue
    rd = float(r[ix["dram__bytes_read.sum"]]) * SCALE[units[ix["dram__bytes_read.sum"]]]
    wr = float(r[ix["dram__bytes_write.sum"]]) * SCALE[units[ix["dram__bytes_write.sum"]]]
    out[key] = {"kernel": name.split("(")[0].replace("void ", ""), "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                "duration_us": float(r[ix["gpu__time_duration.sum"]]) * (1e-3 if units[ix["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1)}
print(json.dumps({"source": "ncu --set full, one launch per kernel at batch 8 x 480x640 (%s); dram writes that are still in L2 "
                            "when the kernel ends are not counted by ncu" % rep.split("/")[-1], "kernels": out}, indent=1))
