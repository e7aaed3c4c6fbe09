#!/usr/bin/env python3
"""A/B timing of the flush (de-duplicate + count) alone: python tools/ab_flush.py [variant ...] [--reads N] [--config del3]
Each variant is a library built by tools/build_variants.py ("default" = the shipped build); run in separate processes."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(variant, config, reads):
    sys.path.insert(0, ROOT)
    import ngs_barcode_count_b200 as bc
    if variant != "default":
        bc.LIB_PATH = os.path.join(bc.PKG, "lib", "variants", f"libbc_b200_{variant}.so")
    import torch
    from ngs_barcode_count_b200 import synth
    wl = synth.Workload(config, f"/tmp/ab_{config}", reads=reads)
    run = wl.run(bc)
    ctr = bc.Counter(run, expected_reads=reads)
    step = 1 << 23
    for a in range(0, reads, step):
        b = wl.generate_device(run, a, min(step, reads - a))
        torch.cuda.synchronize()
        ctr.submit(b)
        ctr.sync()
        del b
    times = []
    for it in range(6):
        ctr.set_option("flush_two_stage", 0)  # invalidates the rows: the next counters() flushes again
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctr.counters()
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    prof = ctr.profile()
    print(json.dumps({"variant": variant, "config": config, "reads": reads, "flush_ms": sorted(times[1:])[len(times[1:]) // 2],
                      "all_ms": [round(t, 2) for t in times], "stages": prof["flush_stages"], "global": prof["flushed_global"]}), flush=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        return child(args[1], args[2], int(args[3]))
    reads, config, variants = 400_000_000, "del3", []
    i = 0
    while i < len(args):
        if args[i] == "--reads":
            reads = int(args[i + 1]); i += 2
        elif args[i] == "--config":
            config = args[i + 1]; i += 2
        else:
            variants.append(args[i]); i += 1
    for v in variants or ["default"]:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", v, config, str(reads)], check=False)


if __name__ == "__main__":
    main()
