#!/bin/bash
# one GPU call: the tests that touch the ingest / the exchange, the ingest sweep, then the bench's FASTQ leg
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider -k "wire or ingest or cli or any_length or exchange" > gpurun_out/pytest_ingest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_ingest.log
python tools/ingest_sweep.py > gpurun_out/ingest_sweep4.jsonl 2> gpurun_out/ingest_sweep4.err; tail -3 gpurun_out/ingest_sweep4.err; cat gpurun_out/ingest_sweep4.jsonl
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-others > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_fused.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_fused.json").read().strip().splitlines()[-1])
print("value", d["value"], "fastq", d["e2e_fastq"])
P
