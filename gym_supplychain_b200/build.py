"""Build libscg.so (the CUDA extension behind the C ABI in include/scg.h) for sm_100a, in-tree.

    python -m gym_supplychain_b200.build [--force]

Plain `nvcc` invocations, no torch extension machinery: the product is a C-ABI shared library
loaded with ctypes.  The fp64 translation unit is compiled with -fmad=false (bit-exact parity with
numpy needs unfused multiply-adds); the fp32 one keeps FMA contraction.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libscg.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
UNITS = {
    "scg_api.cu": [],
    "scg_bg.cu": [],
    "scg_jit.cu": [],
    "scg_sc_f32.cu": [],
    "scg_sc_f64.cu": ["-fmad=false"],
    "scg_sc_mixed.cu": [],
}


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libscg.so cannot be built")
    return exe


def _deps_mtime():
    newest = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            newest = max(newest, os.path.getmtime(os.path.join(root, f)))
    return max(newest, os.path.getmtime(os.path.abspath(__file__)))


def build(force=False, verbose=False, defines=(), tag=None):
    """Compile every translation unit (in parallel) and link libscg.so.  Returns the library path.
    `defines` + `tag` build an experiment variant (extra -D flags) as libscg_<tag>.so next to the product library;
    load it with SCG_LIB=<path> (kernel A/B measurements, never the default)."""
    global BUILD, LIB
    if tag:
        BUILD, LIB = os.path.join(HERE, "_build_" + tag), os.path.join(HERE, "libscg_%s.so" % tag)
    os.makedirs(BUILD, exist_ok=True)
    newest = _deps_mtime()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    nvcc = _nvcc()

    def compile_one(item):
        src, extra = item
        obj = os.path.join(BUILD, src[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            return obj
        cmd = [nvcc] + ARCH + COMMON + list(defines) + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, res.stderr[-4000:]))
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as pool:
        objs = list(pool.map(compile_one, UNITS.items()))
    cmd = [nvcc] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s" % res.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    tag = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--tag=")), None)
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=[a for a in sys.argv if a.startswith("-D") or a.startswith("--ptxas-options")], tag=tag)
    print(path)
