"""Build the UNMODIFIED reference extensions into oracle/_ref/ (test infrastructure only).

Compiles /root/reference/torchext/ext/ext_cpu.cpp (+ ext.h, common.h, co_types.h) where the
sources lie -- nothing is copied into this repo -- with torch.utils.cpp_extension, flags
`-O3` only (no -march, no -ffast-math: SURVEY.md section 8c, the -O0 and -O3 builds are
bit-identical and contain no FMA instructions).  Output: oracle/_ref/ctd_ref_ext_cpu.so,
git-ignored but shipped to the GPU box by gpurun.  /root/reference does not exist on the
GPU box, so there this script only reports whether the prebuilt .so is present.

The reference's own CUDA extension (ext_cuda.cpp + ext_kernel.cu, the generic grid-stride kernel of
common_cuda.h:159-170) is built the same way for sm_100 (torchext/setup.py:10-17 passes no arch flags, so
TORCH_CUDA_ARCH_LIST=10.0 picks -gencode arch=compute_100,code=sm_100) into oracle/_ref/ctd_ref_ext_cuda.so:
the "existing GPU kernel" speed bar, timed by tools/bench_ops.py and bench.py's
`gpu_reference` block on the same box and the same buffers.  A comparison leg only: nothing in
connecting_the_dots_b200/ loads it.

Only tests/, __graft_entry__.smoke(), tools/bench_ops.py and bench.py's baseline legs may load the results.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/torchext/ext"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "ctd_ref_ext_cpu"
NAME_CUDA = "ctd_ref_ext_cuda"


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


def so_path_cuda():
    return os.path.join(OUT_DIR, NAME_CUDA + ".so")


def build_cuda(verbose=False):
    """The reference's CUDA extension for sm_100 (nvcc cross-compiles without a GPU, ~40 s)."""
    srcs = [os.path.join(REF_SRC, f) for f in ("ext_cuda.cpp", "ext_kernel.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return so_path_cuda() if os.path.exists(so_path_cuda()) else None
    deps = srcs + [os.path.join(REF_SRC, f) for f in ("ext.h", "common.h", "common_cuda.h", "co_types.h")]
    if os.path.exists(so_path_cuda()) and all(os.path.getmtime(so_path_cuda()) >= os.path.getmtime(d) for d in deps):
        return so_path_cuda()
    out = os.path.join(OUT_DIR, "cuda")  # its own ninja directory: the two builds must not share build.ninja
    os.makedirs(out, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load
    load(name=NAME_CUDA, sources=srcs, extra_include_paths=[REF_SRC], extra_cflags=["-O3"],
         build_directory=out, verbose=verbose, is_python_module=False, with_cuda=True)
    import shutil
    shutil.copy2(os.path.join(out, NAME_CUDA + ".so"), so_path_cuda())
    return so_path_cuda()


def load_ref(cuda=False):
    """Import the prebuilt reference module (pybind11; needs torch imported first)."""
    import importlib.util
    import torch  # noqa: F401  (the module links against libtorch)
    p = so_path_cuda() if cuda else so_path()
    if not os.path.exists(p):
        return None
    spec = importlib.util.spec_from_file_location(NAME_CUDA if cuda else NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", p)
    pc = None
    try:
        pc = build_cuda(verbose="-v" in sys.argv)
    except Exception as e:  # the GPU bar is optional; the CPU extension is the oracle's anchor
        print("oracle/_ref (cuda) not built:", e)
    print("oracle/_ref (cuda):", pc)
    sys.exit(0 if p else 1)
