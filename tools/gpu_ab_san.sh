#!/bin/bash
# one GPU call: A/B timing of the flush variants, the default bench, then compute-sanitizer memcheck (one tool per call)
mkdir -p gpurun_out
timeout 1200 python tools/ab_flush.py default nobail slots8k ipt8 blocks5 > gpurun_out/ab_flush.jsonl 2> gpurun_out/ab_flush.err
cat gpurun_out/ab_flush.jsonl
timeout 300 python tools/ab_flush.py default --config lineage --reads 100000000 >> gpurun_out/ab_flush.jsonl 2>> gpurun_out/ab_flush.err
tail -1 gpurun_out/ab_flush.jsonl
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_del3.json 2> gpurun_out/bench_del3.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench_del3.err
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider -k "lean or cli or exchange or specialized" > gpurun_out/pytest_gpu_sel.log 2>&1
echo "pytest sel rc=$?"; tail -3 gpurun_out/pytest_gpu_sel.log
bash tools/gpu_sanitizer.sh
