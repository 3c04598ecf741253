"""Build libdxtex_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.

nvcc cross-compiles without a GPU.  Numeric contract of the build:
  -fmad=false            no multiply-add contraction: the BC1-5 / convert / mip kernels must be
                         bit-exact against the reference CPU build (which has no FMA either)
  (default) -prec-div=true -prec-sqrt=true, no --use_fast_math
  -Xcompiler -ffp-contract=off   same for the host code that builds the triangle-filter tables
"""
import os, subprocess, sys, shutil

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OUT = os.path.join(OUT_DIR, "libdxtex_b200.so")
SOURCES = [os.path.join(CSRC, "dxb_api.cu")]
HOST_SOURCES = [os.path.join(HERE, "host", "DirectXTexB200.cpp")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    deps = list(SOURCES) + [os.path.join(HERE, "..", "include", "dxtex_b200.h")]
    for root in (CSRC, os.path.join(HERE, "host")):
        if os.path.isdir(root):
            deps += [os.path.join(root, f) for f in os.listdir(root)]
    return deps


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in _deps() if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = SOURCES + [s for s in HOST_SOURCES if os.path.exists(s)]
    cmd = [_nvcc(), "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-fmad=false",
           "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
           "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
           "-cudart", "static", "-shared",
           "-I", os.path.join(HERE, "..", "include"), "-I", CSRC,
           "-DDXB_BUILDING_LIB",
           "-o", OUT] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libdxtex_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
