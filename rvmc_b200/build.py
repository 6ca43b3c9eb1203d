"""Builds rvmc_b200/librvmc_b200.so (sm_100a only) with nvcc:  python -m rvmc_b200.build [--force] [-v]

The library is built IN-TREE so that it travels with a snapshot of the repo; it is git-ignored.
nvcc cross-compiles without a GPU.  There is exactly one target architecture and no fallback.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OUT = os.path.join(HERE, "librvmc_b200.so")
OBJ_DIR = os.path.join(HERE, "csrc", "build")
SOURCES = ["abi_core.cu", "population.cu", "sigma_fused.cu", "biascorr.cu", "pipe_peaks.cu", "host_bits.cpp"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-I", INCLUDE]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    files += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    files.append(os.path.abspath(__file__))
    return files


def is_stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force=False, verbose=False, bounds_check=False):
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path.
    `bounds_check=True` builds librvmc_b200_bounds.so with device-side assertions on the shared-memory
    indexing (-DRVMC_BOUNDS_CHECK); run it with RVMC_B200_LIB=<path> (compute-sanitizer stand-in)."""
    global OUT, OBJ_DIR
    if bounds_check:
        out, obj_dir = OUT.replace(".so", "_bounds.so"), OBJ_DIR + "_bounds"
        saved = (OUT, OBJ_DIR)
        OUT, OBJ_DIR = out, obj_dir
        try:
            return _build(True, verbose, ["-DRVMC_BOUNDS_CHECK"])
        finally:
            OUT, OBJ_DIR = saved
    if not force and not is_stale():
        return OUT
    return _build(force, verbose, [])


def _build(force, verbose, defines):
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = (["-Xptxas", "-v"] if verbose else []) + list(defines)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout))
        return obj, r.stdout

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    tmp = OUT + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp, "-gencode", "arch=compute_100a,code=sm_100a"] + [o for o, _ in results] + ["-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, bounds_check="--bounds" in sys.argv)
    print(path)
