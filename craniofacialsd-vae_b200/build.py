"""Builds ``libsdvae_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libsdvae_b200.so")
SOURCES = ["sdvae_abi.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh"))       # every header: a stale check that misses one runs old code
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _newest_source_mtime():
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    files.append(os.path.join(os.path.dirname(PKG_DIR), "include", "sdvae_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def is_stale():
    return (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < _newest_source_mtime()


def build_library(force=False, verbose=False):
    """Compile if missing or older than its sources.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build " + LIB_PATH)
    tmp = LIB_PATH + ".tmp%d" % os.getpid()
    extra = ["-DSDVAE_TUNING"] if os.environ.get("SDVAE_TUNING") == "1" else []   # per-role cycle counters / ablations
    extra += os.environ.get("SDVAE_EXTRA_FLAGS", "").split()                      # e.g. compile-time ablations (-DSDVAE_ABL=..)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
