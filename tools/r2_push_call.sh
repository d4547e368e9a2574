#!/bin/bash
# halo push after the loop-structure fix: 1-GPU regression of the default kernel, then N-GPU default vs push
set -u
N=${1:-4}
mkdir -p gpurun_out
python tools/kbench.py --kernels 1 --dtypes f64,f32 --tfuse 4 --tiles 0,5216 --variants 2 > gpurun_out/r2p_kbench.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 280 $TR bench.py --gpus $N --no-cpu-baseline --steps 8 > gpurun_out/r2p_bench_${N}gpu.json 2> gpurun_out/r2p_bench_${N}gpu.err
timeout 280 $TR bench.py --gpus $N --no-cpu-baseline --steps 8 --halo-push 1 > gpurun_out/r2p_bench_${N}gpu_push.json 2> gpurun_out/r2p_bench_${N}gpu_push.err
ls -la gpurun_out | tail -5
