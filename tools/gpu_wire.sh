#!/bin/bash
# one GPU call: the tests that touch the transfer form / the ingest, then the default bench with both host-batch forms
mkdir -p gpurun_out
lscpu | head -25 > gpurun_out/lscpu.txt 2>&1; nproc >> gpurun_out/lscpu.txt; numactl -H >> gpurun_out/lscpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider -k "wire or ingest or cli or any_length or golden_per_read" > gpurun_out/pytest_wire.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_wire.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_wire.json 2> gpurun_out/bench_wire.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_wire.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_wire.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "fastq", d["e2e_fastq"])
P
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-form plain --no-cpu --no-others > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err
echo "bench plain rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_plain.json").read().strip().splitlines()[-1])
print("plain e2e", d["e2e"])
P
