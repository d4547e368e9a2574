#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/r2_k0_sustained.py 768 0,208,216,116 > gpurun_out/r2k_k0_sustained.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
# profiler passes: never a bench value
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2k_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-parity-check > gpurun_out/r2k_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_r1 -s 3 -c 1 -o gpurun_out/r2k_fused \
  python tools/kbench.py --kernels 1 --dtypes f64 --tfuse 4 --tiles 0 --variants 2 --reps 3 > gpurun_out/r2k_ncu.log 2>&1
ls -la gpurun_out | tail -6
