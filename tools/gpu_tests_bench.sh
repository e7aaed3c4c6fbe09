#!/bin/bash
# one GPU call: the GPU test suite, then a short bench (no ncu in this call)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_del3.json 2> gpurun_out/bench_del3.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_del3.err
head -c 600 gpurun_out/bench_del3.json
