"""Build an A/B variant of libctd_b200.so with extra nvcc defines for ONE source file (experiments only).
python tools/experiments/build_variant.py census_sym.cu -DCTD_CS_NW=5 -o /root/repo/gpurun_out/libctd_nw5.so
Use with CTD_B200_LIB=<path>."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from connecting_the_dots_b200 import _build
src, out = sys.argv[1], sys.argv[sys.argv.index("-o") + 1]
defs = [a for a in sys.argv[2:] if a.startswith("-D")]
_build.build()
obj = out + ".o"
r = subprocess.run([_build.NVCC] + _build.FLAGS + defs + ["-c", os.path.join(_build.CSRC, src), "-o", obj], capture_output=True, text=True)
assert r.returncode == 0, r.stderr[-3000:]
objs = [os.path.join(_build.OBJ, s.replace(".cu", ".o")) for s in _build.SOURCES if s != src] + [obj]
r = subprocess.run([_build.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-Xcompiler", "-fPIC"], capture_output=True, text=True)
assert r.returncode == 0, r.stderr[-3000:]
os.remove(obj)
print("built", out)
