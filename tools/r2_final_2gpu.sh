#!/bin/bash
# final check of the multi-GPU default path after the thin-slab fix: parity tests with >= 2 GPUs + one short 2-GPU bench line
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "halo_copy or thin_slabs or push or multi or slab" > gpurun_out/r2_final_tests.log 2>&1
tail -3 gpurun_out/r2_final_tests.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-configs > gpurun_out/r2_final_bench2.json 2> gpurun_out/r2_final_bench2.err
tail -c 1500 gpurun_out/r2_final_bench2.json
