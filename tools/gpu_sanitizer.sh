#!/bin/bash
# one GPU call: compute-sanitizer memcheck over the smallest golden cases (the decode kernel reads one word past a plane
# inside shared memory by design; the flush publishes shared-memory table entries with fence + CAS)
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import ngs_barcode_count_b200 as bc
from helpers import load_golden, read_fastq
for case in ("del3_umi", "lineage_raw", "crispr", "example_q20"):
    exp, p = load_golden(case)
    fl = exp["flags"]
    reads = read_fastq(p["fastq"])
    run = bc.Run(p["fmt"], p["samples"], p["counted"], min_quality=fl["min_quality"])
    ctr = bc.Counter(run, device=0)
    ctr.submit(run.pack([r[0] for r in reads], [r[1] for r in reads]))
    got = ctr.counters(); got.pop("unsupported")
    assert got == exp["counters"], (case, got, exp["counters"])
    ctr.enrich()
    ranks = [bc.Counter(run) for _ in range(2)]
    if ranks[0].profile()["deferred_count"]:
        from ngs_barcode_count_b200.multi import exchange_plan
        b = run.pack([r[0] for r in reads], [r[1] for r in reads])
        for r, c in enumerate(ranks):
            c.exchange_open(2, r, b.n + 8)
        for c in ranks:
            c.exchange_connect_local(ranks)
        ranks[0].submit(b.slice(0, b.n // 2)); ranks[1].submit(b.slice(b.n // 2, b.n))
        m = [c.exchange_count(2) for c in ranks]
        for r, c in enumerate(ranks):
            c.exchange_scatter(exchange_plan(m, r)[0])
        for c in ranks:
            c.sync()
        for r, c in enumerate(ranks):
            c.exchange_finish(exchange_plan(m, r)[1])
        tot = sum(c.counters()["matched"] for c in ranks)
        assert tot == exp["counters"]["matched"], (case, tot)
    print("sanitizer case", case, "ok", flush=True)
PY
python /tmp/san_case.py > gpurun_out/san_plain.log 2>&1 &&
compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/r2_sanitizer_memcheck.log python /tmp/san_case.py > gpurun_out/san_memcheck_stdout.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_memcheck.log
