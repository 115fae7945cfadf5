"""Build the UNMODIFIED reference CPU extension into oracle/_ref/ (test infrastructure only).

Compiles /root/reference/torchext/ext/ext_cpu.cpp (+ ext.h, common.h, co_types.h) where the
sources lie -- nothing is copied into this repo -- with torch.utils.cpp_extension, flags
`-O3` only (no -march, no -ffast-math: SURVEY.md section 8c, the -O0 and -O3 builds are
bit-identical and contain no FMA instructions).  Output: oracle/_ref/ctd_ref_ext_cpu.so,
git-ignored but shipped to the GPU box by gpurun.  /root/reference does not exist on the
GPU box, so there this script only reports whether the prebuilt .so is present.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the result.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/torchext/ext"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "ctd_ref_ext_cpu"


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    """Returns the path of the built module, or None when neither sources nor a prebuilt .so exist."""
    src = os.path.join(REF_SRC, "ext_cpu.cpp")
    if not os.path.exists(src):
        return so_path() if os.path.exists(so_path()) else None
    deps = [src] + [os.path.join(REF_SRC, f) for f in ("ext.h", "common.h", "co_types.h")]
    if os.path.exists(so_path()) and all(os.path.getmtime(so_path()) >= os.path.getmtime(d) for d in deps):
        return so_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=[src], extra_include_paths=[REF_SRC], extra_cflags=["-O3"],
         build_directory=OUT_DIR, verbose=verbose, is_python_module=False)
    return so_path()


def load_ref():
    """Import the prebuilt reference module (pybind11; needs torch imported first)."""
    import importlib.util
    import torch  # noqa: F401  (the module links against libtorch)
    p = so_path()
    if not os.path.exists(p):
        return None
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", p)
    sys.exit(0 if p else 1)
