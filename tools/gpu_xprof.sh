#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider -k "exchange or several_contexts" > gpurun_out/pytest_x.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_x.log
python tools/exchange_profile.py 8 16777216 bulk > gpurun_out/xprof_bulk.jsonl 2> gpurun_out/xprof.err; tail -1 gpurun_out/xprof_bulk.jsonl
python tools/exchange_profile.py 8 16777216 > gpurun_out/xprof_stream.jsonl 2>> gpurun_out/xprof.err; tail -1 gpurun_out/xprof_stream.jsonl
tail -3 gpurun_out/xprof.err
ncu --set full --clock-control none --import-source on -k "regex:k_owner_scatter" -s 2 -c 1 -f -o gpurun_out/r2_owner_scatter python tools/exchange_profile.py 8 16777216 bulk > gpurun_out/ncu_xprof.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_xprof.log
