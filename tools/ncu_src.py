#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: per-opcode executed counts / stall samples, top lines,
stall reasons.  usage: ncu -i rep --page source --csv --kernel-name regex:X | python tools/ncu_src.py"""
import collections
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for sec in sections[:1]:
    hdr, data = sec["hdr"], sec["data"]
    ix = {n: i for i, n in enumerate(hdr)}
    I = lambda r, n: int(float(r[ix[n]] or 0))
    print("==", sec["name"][:100], len(data), "sass lines")
    tot_s = sum(I(r, "# Samples") for r in data)
    tot_i = sum(I(r, "Instructions Executed") for r in data)
    print("samples", tot_s, "warp-instructions", tot_i)
    c, s = collections.Counter(), collections.Counter()
    for r in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip()).split(" ")[0].split(".")[0]
        c[op] += I(r, "Instructions Executed")
        s[op] += I(r, "# Samples")
    for op, n in c.most_common(24):
        print("  %-8s inst %10d (%.1f%%)  samples %7d (%.1f%%)" % (op, n, 100.0 * n / max(tot_i, 1), s[op], 100.0 * s[op] / max(tot_s, 1)))
    stalls = [n for n in hdr if n.startswith("stall_") and not n.endswith("_not_issued")]
    agg = {n: sum(I(r, n) for r in data) for n in stalls}
    print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot_s, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:top_n]:
        print("  %s smp %6d inst %9d  %s" % (r[ix["Address"]][-5:], I(r, "# Samples"), I(r, "Instructions Executed"), r[ix["Source"]][:100]))
    conf = sum(I(r, "L1 Wavefronts Shared Excessive") for r in data) if "L1 Wavefronts Shared Excessive" in ix else 0
    print("shared excessive wavefronts:", conf, "of", sum(I(r, "L1 Wavefronts Shared") for r in data))
