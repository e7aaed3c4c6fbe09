#!/usr/bin/env python3
"""Builds alternative libbc_b200 libraries (compile-time constants of the flush kernels) into
ngs-barcode-count_b200/lib/variants/ for A/B timing with tools/ab_flush.py.  Development aid; the product is the default build."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ngs-barcode-count_b200"))
import build as b  # noqa: E402

VARIANTS = {
    "checked": ["-DBC_DECODE_CHECKED=true"],
    "hash32": ["-DBC_PART_HASH32=1"],
    "nobail": ["-DBC_CHAIN_LIMIT=100000000"],
    "slots8k": ["-DBC_TABLE_SLOTS=8192"],
    "ipt8": ["-DBC_SCATTER_IPT=8"],
    "blocks5": ["-DBC_REDUCE_BLOCKS=5"],
}


def main():
    out = os.path.join(b.LIB, "variants")
    os.makedirs(out, exist_ok=True)
    b._embed_sources(os.path.join(b.CSRC, "bc_jit_sources.inc"))
    src = [os.path.join(b.CSRC, f) for f in ("bc_kernels.cu", "bc_partition.cu", "bc_api.cu", "bc_jit.cu")] + [os.path.join(b.CSRC, "host", "bc_host.cpp")]
    procs = []
    for name, defs in VARIANTS.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        lib = os.path.join(out, f"libbc_b200_{name}.so")
        cmd = [b.NVCC] + b.ARCH + b.COMMON + defs + ["-shared", "-o", lib] + src + ["-lz", "-ldl"]
        print("+", " ".join(cmd), flush=True)
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("variant build failed")


if __name__ == "__main__":
    main()
