#!/bin/bash
# Multi-GPU call of the next round (gpurun --gpus N --timeout 600 -- 'bash tools/round2_multi_gpu.sh N').
# Validates the paths that have only run on the CPU emulator so far (every one wrapped in `timeout`: a wrong flag
# protocol in the halo push would spin until its own device-side limit of ~10 s per wait), then times them.
set -u
N=${1:-2}
mkdir -p gpurun_out
GIRIH_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "halo_push or uneven or xy_topologies or z_slabs" 2>&1 | tail -15 > gpurun_out/r2_multi_pytest_$N.log
for dp in build build_dp; do
  timeout 60 ./$dp/mwd_kernel --nx 256 --ny 256 --nz 256 --nt 50 --target-ts 2 --target-kernel 1 --t-dim 3 --verify 1 \
    --npz $N --gpu-push 1 --verbose 0 >> gpurun_out/r2_multi_cli_push_$N.log 2>&1
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 280 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
timeout 280 $TR bench.py --gpus $N --no-cpu-baseline --halo-push 1 > gpurun_out/r2_bench_${N}gpu_push.json 2> gpurun_out/r2_bench_${N}gpu_push.err
if [ "${OVERLAP:-0}" = 1 ]; then
timeout 280 $TR bench.py --gpus $N --no-cpu-baseline --overlap 1 > gpurun_out/r2_bench_${N}gpu_overlap.json 2> gpurun_out/r2_bench_${N}gpu_overlap.err
fi
ls -la gpurun_out
