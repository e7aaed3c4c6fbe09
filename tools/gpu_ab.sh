#!/bin/bash
# one GPU call: A/B timing of flush variants (tools/build_variants.py) against the shipped build
mkdir -p gpurun_out
timeout 900 python tools/ab_flush.py default hash32 > gpurun_out/ab_hash.jsonl 2> gpurun_out/ab_hash.err
timeout 600 python tools/ab_flush.py default hash32 --config lineage --reads 100000000 >> gpurun_out/ab_hash.jsonl 2>> gpurun_out/ab_hash.err
cat gpurun_out/ab_hash.jsonl; tail -3 gpurun_out/ab_hash.err
