"""Generate tests/golden/lcn_cython.npz from the UNMODIFIED reference Cython module data/lcn/lcn.pyx.

The .pyx is compiled where it lies (cythonize into a temporary directory, nothing copied into this repo) and its
`normalize` is called on seeded inputs; inputs and outputs are stored.  Run in the build container only.

    python tests/golden/make_golden_lcn_cython.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data/lcn/lcn.pyx"


def build():
    d = tempfile.mkdtemp(prefix="ctd_lcn_pyx_")
    with open(os.path.join(d, "setup.py"), "w") as f:
        f.write("from setuptools import setup, Extension\nfrom Cython.Build import cythonize\n"
                "setup(ext_modules=cythonize([Extension('lcn', [%r])], language_level=3, build_dir=%r))\n" % (REF, d))
    subprocess.run([sys.executable, "setup.py", "build_ext", "--build-lib", d, "--build-temp", d], cwd=d, check=True, capture_output=True)
    sys.path.insert(0, d)
    import lcn
    return lcn


def main():
    lcn = build()
    rng = np.random.RandomState(21)
    out = {}
    for name, (M, N, ks, eps) in {"a": (24, 37, 4, 0.01), "b": (40, 33, 5, 0.1), "c": (9, 9, 4, 0.05), "d": (12, 20, 0, 0.01), "e": (7, 30, 4, 0.01)}.items():
        x = rng.rand(M, N).astype(np.float32)
        if name == "b":
            x[10:25, 5:20] = 0.5      # flat window: std = 0, division by eps only
        l, s = lcn.normalize(x, ks, eps)
        out[name + "_x"], out[name + "_lcn"], out[name + "_std"] = x, np.asarray(l), np.asarray(s)
        out[name + "_args"] = np.array([ks, eps], np.float64)
    np.savez_compressed(os.path.join(HERE, "lcn_cython.npz"), **out)
    print("wrote lcn_cython.npz", {k: v.shape for k, v in out.items() if k.endswith("_lcn")})


if __name__ == "__main__":
    main()
