#!/bin/bash
# halo copy (overlapped schedule + copy engines): parity on N GPUs, then default vs copy vs overlap(NCCL)
set -u
N=${1:-4}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "halo_copy" 2>&1 | tail -5 > gpurun_out/r2c_pytest_$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 280 $TR bench.py --gpus $N --no-cpu-baseline --steps 8 > gpurun_out/r2c_bench_${N}gpu.json 2> gpurun_out/r2c_bench_${N}gpu.err
timeout 280 $TR bench.py --gpus $N --no-cpu-baseline --steps 8 --halo-copy 1 > gpurun_out/r2c_bench_${N}gpu_copy.json 2> gpurun_out/r2c_bench_${N}gpu_copy.err
timeout 280 $TR bench.py --gpus $N --no-cpu-baseline --steps 8 --overlap 1 > gpurun_out/r2c_bench_${N}gpu_overlap.json 2> gpurun_out/r2c_bench_${N}gpu_overlap.err
ls -la gpurun_out | tail -5
