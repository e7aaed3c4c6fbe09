#!/bin/bash
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -Xcompiler -pthread -I include -o /tmp/host_pack_bench tools/host_pack_bench.cpp -L ngs-barcode-count_b200/lib -lbc_b200 -lz -Xlinker -rpath -Xlinker $PWD/ngs-barcode-count_b200/lib 2>&1 | grep -E "error" 
python - <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
import ngs_barcode_count_b200 as bc
from ngs_barcode_count_b200 import synth
wl = synth.Workload("del3", "/tmp/hb_del3", reads=8_000_000)
wl.write_fastq("/tmp/hb_del3/r.fastq", 0, 8_000_000, threads=16)
P
for t in 1 8 16; do for m in 1 2 3 0; do /tmp/host_pack_bench /tmp/hb_del3/r.fastq $t $m | tail -1; done; done | tee gpurun_out/host_pack_bench.txt
