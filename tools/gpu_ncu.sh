#!/bin/bash
# one GPU call: plain run, then the ncu launch list and full captures (with source) of the decode and flush kernels
mkdir -p gpurun_out
CMD="python bench.py --reads 33554432 --steps 1 --warmup 1 --no-cpu --no-e2e --no-others"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 200 \
    --csv --log-file gpurun_out/r2_launches_del3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_decode -s 4 -c 1 -f -o gpurun_out/r2_decode_del3 $CMD > gpurun_out/ncu_decode.log 2>&1
echo "decode rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:k_reduce|k_split_scatter|k_split_hist|k_marginals" -s 5 -c 6 -f -o gpurun_out/r2_flush_del3 $CMD > gpurun_out/ncu_flush.log 2>&1
echo "flush rc=$?"
ls -la gpurun_out/*.ncu-rep
