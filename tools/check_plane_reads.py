#!/usr/bin/env python3
"""Verification aid (compute-sanitizer is closed on this pool): k_decode reads its bit planes in shared memory without a
bounds check and may touch one word past a plane (bc_decode.cuh: plane_bits).  This runs the golden cases and 2 M reads of
every BASELINE workload through the shipped library and through a build whose plane reads are bounds-checked
(tools/build_variants.py checked) — generic kernel in both — and compares every per-read output.
    python tools/build_variants.py checked && python tools/check_plane_reads.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(variant):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hashlib
    import numpy as np
    import ngs_barcode_count_b200 as bc
    if variant != "default":
        bc.LIB_PATH = os.path.join(bc.PKG, "lib", "variants", f"libbc_b200_{variant}.so")
    from helpers import golden_cases, load_golden, read_fastq
    from ngs_barcode_count_b200 import synth
    out = {}

    def digest(d):
        h = hashlib.sha256()
        for k in ("status", "offset", "repaired", "slot_index", "key_lo", "key_hi"):
            h.update(np.ascontiguousarray(d[k]).tobytes())
        return h.hexdigest()[:16]

    for case in golden_cases():
        exp, p = load_golden(case)
        fl = exp["flags"]
        run = bc.Run(p["fmt"], p["samples"], p["counted"], min_quality=fl["min_quality"], max_barcode=fl["max_barcode"],
                     max_sample=fl["max_sample"], max_constant=fl["max_constant"])
        ctr = bc.Counter(run, flags=bc.BC_CFG_NO_SPECIALIZE)
        reads = read_fastq(p["fastq"])
        out["golden/" + case] = digest(ctr.decode_only(run.pack([r[0] for r in reads], [r[1] for r in reads])))
    for name in ("del3", "crispr", "lineage", "example"):
        wl = synth.Workload(name, f"/tmp/chk_{name}", reads=2_000_000)
        run = wl.run(bc)
        ctr = bc.Counter(run, flags=bc.BC_CFG_NO_SPECIALIZE)
        b = wl.generate_device(run, 0, 2_000_000)
        import torch
        torch.cuda.synchronize()
        out["workload/" + name] = digest(ctr.decode_only(b))
        ctr.submit(b)
        out["counters/" + name] = ctr.counters()
    print(json.dumps({"variant": variant, "results": out}), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        return child(sys.argv[2])
    res = {}
    for v in ("default", "checked"):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", v], capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stderr[-2000:])
            raise SystemExit(1)
        res[v] = json.loads(r.stdout.strip().splitlines()[-1])["results"]
    same = res["default"] == res["checked"]
    print(json.dumps({"identical": same, "cases": len(res["default"]), "default": res["default"],
                      "differences": {k: (res["default"][k], res["checked"].get(k)) for k in res["default"] if res["default"][k] != res["checked"].get(k)}}, indent=1))
    raise SystemExit(0 if same else 1)


if __name__ == "__main__":
    main()
