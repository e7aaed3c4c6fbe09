#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider -k "exchange or several_contexts" > gpurun_out/pytest_x.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_x.log
for m in bulk streamed px; do python tools/exchange_profile.py 8 16777216 $m > gpurun_out/xprof_$m.jsonl 2>> gpurun_out/xprof.err; tail -1 gpurun_out/xprof_$m.jsonl; done
tail -3 gpurun_out/xprof.err
