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
NAMES = {"lcn_tma_kernel": "lcn_fwd", "photo_fwd_box9_tma": "sad_fwd", "photo_bwd_box9_tma": "sad_bwd",
         "photo_fwd_census9": "census_sad_fwd", "photo_bwd_census9<3, false>": "census_sad_bwd",
         "photo_bwd_census9<3, true>": "census_sad_fwd_bwd", "masked_sums_kernel": "masked_sums"}
out = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = next((v for k, v in NAMES.items() if k.split("<")[0] in name and (("<" not in k) or k.split("<")[1].rstrip(">") in name.replace("(int)", "").replace("(bool)", "").replace("1", "true").replace("0", "false"))), None)
    if key is None or key in out:
        continue
    rd = float(r[ix["dram__bytes_read.sum"]]) * SCALE[units[ix["dram__bytes_read.sum"]]]
    wr = float(r[ix["dram__bytes_write.sum"]]) * SCALE[units[ix["dram__bytes_write.sum"]]]
    out[key] = {"kernel": name.split("(")[0].replace("void ", ""), "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                "duration_us": float(r[ix["gpu__time_duration.sum"]]) * (1e-3 if units[ix["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1)}
print(json.dumps({"source": "ncu --set full, one launch per kernel at batch 8 x 480x640 (%s); dram writes that are still in L2 "
                            "when the kernel ends are not counted by ncu" % rep.split("/")[-1], "kernels": out}, indent=1))
