#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --config crispr --reads 33554432 --steps 1 --warmup 1 --no-cpu --no-e2e --no-others"
$CMD > gpurun_out/crispr_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode -s 4 -c 1 -f -o gpurun_out/r2_decode_crispr $CMD > gpurun_out/ncu_crispr.log 2>&1
echo "crispr rc=$?"
CMD="python bench.py --config lineage --reads 33554432 --steps 1 --warmup 1 --no-cpu --no-e2e --no-others"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_lineage.csv $CMD > gpurun_out/ncu_lineage.log 2>&1
echo "lineage rc=$?"
