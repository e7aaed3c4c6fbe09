"""In-tree build of the sm_100a library and CLI (nvcc cross-compiles without a GPU).

    python ngs-barcode-count_b200/build.py [--force]

Outputs (git-ignored, shipped to the GPU box by gpurun):
    ngs-barcode-count_b200/lib/libbc_b200.so     kernels + C ABI (include/bc_b200.h) + host side (include/bc_host.h)
    ngs-barcode-count_b200/lib/libbc_synth.so    synthetic-read generator used by bench.py and the tests
    ngs-barcode-count_b200/bin/barcode-count     the reference's command line over the GPU path
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
BIN = os.path.join(PKG, "bin")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wextra,-pthread", "-I", os.path.join(ROOT, "include")]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force=False, verbose_ptxas=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(BIN, exist_ok=True)
    headers = [os.path.join(ROOT, "include", h) for h in ("bc_b200.h", "bc_host.h")] + [
        os.path.join(CSRC, h) for h in ("bc_device.cuh", "bc_kernels.h")]
    lib = os.path.join(LIB, "libbc_b200.so")
    lib_src = [os.path.join(CSRC, "bc_kernels.cu"), os.path.join(CSRC, "bc_partition.cu"), os.path.join(CSRC, "bc_api.cu"),
               os.path.join(CSRC, "host", "bc_host.cpp")]
    extra = ["-Xptxas", "-v"] if verbose_ptxas else []
    if force or _newer(lib, lib_src + headers):
        _run([NVCC] + ARCH + COMMON + extra + ["-shared", "-o", lib] + lib_src + ["-lz"])
    synth = os.path.join(LIB, "libbc_synth.so")
    synth_src = [os.path.join(CSRC, "bc_synth.cu")]
    if os.path.exists(synth_src[0]) and (force or _newer(synth, synth_src + headers + [os.path.join(CSRC, "bc_synth.h")])):
        _run([NVCC] + ARCH + COMMON + extra + ["-shared", "-o", synth] + synth_src)
    cli = os.path.join(BIN, "barcode-count")
    cli_src = [os.path.join(CSRC, "host", "main.cpp")]
    if force or _newer(cli, cli_src + [lib] + headers):
        _run([NVCC] + ARCH + COMMON + ["-o", cli] + cli_src + ["-L", LIB, "-lbc_b200", "-lz", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../lib"])
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="--ptxas" in sys.argv)
