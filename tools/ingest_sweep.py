#!/usr/bin/env python3
"""Development aid: bch_count_fastq on an 8 M-read DEL file over host thread counts / batch sizes / ingest options,
with the ingest thread's phase timers.  python tools/ingest_sweep.py [reads]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ngs_barcode_count_b200 as bc  # noqa: E402
from ngs_barcode_count_b200 import synth  # noqa: E402


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
    wl = synth.Workload("del3", "/tmp/ingest_sweep", reads=reads)
    path = "/tmp/ingest_sweep/r.fastq"
    wl.write_fastq(path, 0, reads, threads=os.cpu_count())
    run = wl.run(bc)
    ctr = bc.Counter(run, expected_reads=reads)
    ctr.count_fastq(path, threads=os.cpu_count(), batch_reads=1 << 20)
    ctr.counters()
    for fused, wire, pop in ((1, 1, 0), (1, 1, 1), (1, 1, 2), (0, 1, 0)):
        run.set_option("fused_ingest", fused)
        run.set_option("wire_batches", wire)
        run.set_option("mmap_populate", pop)
        for threads in (8, 16):
            for batch in (1 << 20,):
                best = None
                for _ in range(3):
                    ctr.reset()
                    t0 = time.perf_counter()
                    n = ctr.count_fastq(path, threads=threads, batch_reads=batch)
                    ctr.counters()
                    dt = time.perf_counter() - t0
                    if best is None or dt < best[0]:
                        best = (dt, ctr.ingest_stats())
                print(json.dumps({"fused": fused, "wire": wire, "populate": pop, "threads": threads, "batch": batch, "Mreads_s": round(n / best[0] / 1e6, 1),
                                  "wall_ms": round(best[0] * 1e3, 1), **{k: round(v * 1e3, 1) if k.endswith("_s") else v for k, v in best[1].items()}}), flush=True)


if __name__ == "__main__":
    main()
