#!/bin/bash
# one GPU call: GPU test suite, the default bench, then ncu (launch list + full captures with source) of a reduced bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_del3.json 2> gpurun_out/bench_del3.err
echo "bench rc=$?"; tail -c 800 gpurun_out/bench_del3.err
CMD="python bench.py --reads 33554432 --steps 1 --warmup 1 --no-cpu --no-e2e --no-others"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 200 \
    --csv --log-file gpurun_out/r2_launches_del3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_decode -s 4 -c 1 -f -o gpurun_out/r2_decode_del3 $CMD > gpurun_out/ncu_decode.log 2>&1
echo "decode rc=$?"
# the step's flush: first big histogram/scatter/reduce launches of the timed step (the warm-up step has the same 12 matching launches before)
ncu --set full --clock-control none --import-source on -k "regex:k_reduce|k_split_scatter|k_split_hist" -s 12 -c 5 -f -o gpurun_out/r2_flush_del3 $CMD > gpurun_out/ncu_flush.log 2>&1
echo "flush rc=$?"
