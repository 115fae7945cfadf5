#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libctd_b200.so (cuobjdump -sass): evidence that the TMA / mbarrier / MUFU / atomic
instructions the design claims are really in the binary.  Writes profiles/<round>_sass_summary.txt.

    python tools/sass_summary.py [r02]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "connecting_the_dots_b200", "libctd_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
WATCH = ["UTMALDG", "UBLKCP", "SYNCS", "MUFU.RSQ", "MUFU.RCP", "MUFU", "ATOMG", "ATOMS", "ATOM", "RED", "SHFL", "LDS", "STS", "LDG", "STG", "BAR.SYNC",
         "DADD", "DFMA", "FFMA", "HMMA", "UTCHMMA", "UTCMMA"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                cur[w] += 1
rows = []
for name, c in kernels.items():
    d = demangle(name)
    d = d.replace("(anonymous namespace)::", "").replace("ctd::", "").replace("void ", "")
    d = re.sub(r"\((?:const |float|double|long|int|unsigned|CUtensorMap|SymGeom|Geom|uchar|char|bool|void).*$", "", d)[:70]
    rows.append((d, c))
cols = ["_total"] + [w for w in WATCH if any(c[w] for _, c in rows)]
out = ["# cuobjdump -sass libctd_b200.so (sm_100a): instruction counts per kernel (static, whole function body)",
       "# columns: " + " ".join(cols), ""]
w = max(len(d) for d, _ in rows)
out.append("%-*s " % (w, "kernel") + " ".join("%9s" % c.strip("_") for c in cols))
for d, c in sorted(rows):
    out.append("%-*s " % (w, d) + " ".join("%9d" % c[k] for k in cols))
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
out.append("")
out.append("%-*s " % (w, "ALL KERNELS (%d)" % len(rows)) + " ".join("%9d" % tot[k] for k in cols))
path = os.path.join(ROOT, "profiles", tag + "_sass_summary.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path, len(rows), "kernels;", {k: tot[k] for k in ("UTMALDG", "UBLKCP", "SYNCS", "MUFU.RSQ", "ATOMG", "RED", "HMMA", "UTCHMMA")})
