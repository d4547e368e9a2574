#!/bin/bash
# scaling check on one 8-GPU box: bench.py with its defaults (halo copy, parity_check, strong-scaling C5) at the given N
set -u
mkdir -p gpurun_out
for N in "$@"; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
  ( time timeout 400 $TR bench.py --gpus $N --no-cpu-baseline --steps 5 --warmup 3 ) > gpurun_out/r2s_bench_${N}gpu.json 2> gpurun_out/r2s_bench_${N}gpu.err
done
ls -la gpurun_out | tail -4
