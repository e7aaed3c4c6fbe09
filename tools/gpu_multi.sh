#!/bin/bash
# N-GPU call: N-rank == 1-GPU parity (tests/test_multi_gpu.py), then the scaling bench at N (argument: number of GPUs)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_multi_${N}.log 2>&1
echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_${N}.log
for n in $(seq 1 1); do :; done
PORT=29731
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-others > gpurun_out/scale_${N}.json 2> gpurun_out/scale_${N}.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 5 --warmup 3 \
      > gpurun_out/scale_${N}.json 2> gpurun_out/scale_${N}.err
fi
echo "bench N=$N rc=$?"; tail -c 600 gpurun_out/scale_${N}.err; grep -o '"value": [0-9.e+]*' gpurun_out/scale_${N}.json | head -2
grep -o '"parity_n_ranks": {[^}]*}' gpurun_out/scale_${N}.json
