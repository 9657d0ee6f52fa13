"""Builds libuvs_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python uncalibrated-visual-servoing_b200/build.py [--force] [--verbose]

One object per .cu (compiled in parallel), linked with the static CUDA runtime so the
library has no dependency on torch or on a system libcudart.  -lineinfo keeps ncu's
source page usable.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
# tuning builds: UVS_DENSE_MINB=<n> python build.py --out libuvs_b200_minb<n>.so ; select with UVS_B200_LIB=<path>
LIB = os.path.join(HERE, "libuvs_b200.so")
# longest compiles first (8 worker threads): the per-M units of the three kernel families dominate the build
SOURCES = ["blockdiag_m8.cu", "blockdiag_m6.cu", "twist_m8.cu", "dense_m8.cu", "twist_m6.cu", "dense_m6.cu", "blockdiag_m2.cu",
           "dense_m2.cu", "twist_m2.cu", "dense_quad_m8.cu", "dense_quad_m6.cu", "api.cu", "aux_kernels.cu", "blockdiag_kernels.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
         "-Xptxas", "-v"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(os.path.dirname(HERE), "include", "uvs_b200.h"), os.path.abspath(__file__)]


def _extra_defs():
    names = ["UVS_DENSE_MINB", "UVS_DENSE_WARPS", "UVS_DENSE_L_SMEM", "UVS_DENSE_MAXNREG", "UVS_BD_MINB", "UVS_DENSE_FAST",
             "UVS_PHASE_CLOCKS", "UVS_BD_PAD", "UVS_QUAD_MAXNREG", "UVS_DENSE_QUAD", "UVS_QUAD_GAIN_RB", "UVS_QUAD_RANK_UNROLL", "UVS_BLK_STRIDE"]
    names += sorted(k for k in os.environ if k.startswith("UVS_OPT_"))
    return [f"-D{k}={os.environ[k]}" for k in names if os.environ.get(k)]


def _compile(src, verbose, tag=""):
    obj = os.path.join(OBJ, src.replace(".cu", tag + ".o"))
    cmd = [NVCC] + FLAGS + _extra_defs() + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, src.replace(".cu", tag + ".ptxas.log"))
    with open(log, "w") as fh:
        fh.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return obj


def build(force=False, verbose=False, out=None):
    os.makedirs(OBJ, exist_ok=True)
    lib = os.path.join(HERE, out) if out else LIB
    tag = "." + os.path.splitext(out)[0] if out else ""
    if not force and os.path.exists(lib):
        newest = max(os.path.getmtime(p) for p in _deps())
        if os.path.getmtime(lib) >= newest:
            return lib
    with cf.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 4, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose, tag), SOURCES))
    cmd = [NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=out))
