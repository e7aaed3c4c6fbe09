#!/usr/bin/env python3
"""Development aid: the owner scatter of the multi-GPU exchange with the ranks as contexts on ONE GPU (peer stores land in
local HBM, so what is measured is the SM-side cost of the kernel).  python tools/exchange_profile.py [ranks] [reads per rank] [bulk]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import ngs_barcode_count_b200 as bc  # noqa: E402
from ngs_barcode_count_b200 import synth  # noqa: E402
from ngs_barcode_count_b200.multi import exchange_plan  # noqa: E402


def main():
    n_ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
    bulk = len(sys.argv) > 3 and sys.argv[3] == "bulk"
    wl = synth.Workload("del3", "/tmp/xprof", reads=per * n_ranks)
    run = wl.run(bc)
    ranks = [bc.Counter(run, expected_reads=per) for _ in range(n_ranks)]
    for r, c in enumerate(ranks):
        c.exchange_open(n_ranks, r, int(per * 1.3))
    for c in ranks:
        c.exchange_connect_local(ranks)
        c.set_option("exchange_bulk", int(bulk))
        c.set_profiling(True)
    step = 1 << 23
    batches = [[wl.generate_device(run, r * per + a, min(step, per - a)) for a in range(0, per, step)] for r in range(n_ranks)]
    torch.cuda.synchronize()
    for it in range(3):
        for c in ranks:
            c.reset()
            c.reset_profile()
        for r, c in enumerate(ranks):
            for b in batches[r]:
                c.submit(b)
        matrix = [c.exchange_count(n_ranks) for c in ranks]
        for r, c in enumerate(ranks):
            c.exchange_scatter(exchange_plan(matrix, r)[0])
        for c in ranks:
            c.sync()
        for r, c in enumerate(ranks):
            c.exchange_finish(exchange_plan(matrix, r)[1])
        p = ranks[0].profile()
        print(json.dumps({"ranks": n_ranks, "per_rank": per, "bulk": bulk, "ms": p["ms"], "launches": p["launches"]}), flush=True)


if __name__ == "__main__":
    main()
