#!/bin/bash
# First gpurun call of a round: parity, bench line, kernel sweep incl. the opt-in sweep schedules of round 1b,
# launch list and one full ncu capture of the fused sweep.  Everything lands in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
# Copy what should be judged from gpurun_out/ to profiles/ afterwards (tools/ncu_summary.py summarises a .ncu-rep).
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu.log
GIRIH_RUN_UNVALIDATED=1 python -m pytest tests -q -m gpu -k pipelined 2>&1 | tail -5 > gpurun_out/r2_pytest_unvalidated.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
  --format=csv -lms 200 > gpurun_out/r2_clocks.csv &
SMI=$!
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --arith contract > gpurun_out/r2_bench_contract.json 2>> gpurun_out/r2_bench.err
python bench.py --e2e-mode pipelined --no-cpu-baseline > gpurun_out/r2_bench_e2e_pipelined.json 2>> gpurun_out/r2_bench.err
kill $SMI
# default tiles vs split barrier (5xxx) vs decoupled levels (7xxx), then every tile through the sweep tool
python tools/split_bench.py > gpurun_out/r2_sweep_schedules.log 2>&1
python tools/kbench.py --kernels 1 --dtypes f64,f32 --tfuse 2,3,4 --tiles 0,216,408,312,5408,5216,9408,9216 --variants 2 \
  > gpurun_out/r2_kbench_k1.log 2>&1
python tools/kbench.py --kernels 0,4,7 --dtypes f64,f32 --tfuse 1 > gpurun_out/r2_kbench_others.log 2>&1
python tools/kbench.py --kernels 2,3,5 --dtypes f64,f32 --tfuse 1,2,3 --tiles 0,9216 >> gpurun_out/r2_kbench_others.log 2>&1
# profiler passes: never a bench value
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_r1 -s 3 -c 2 -o gpurun_out/r2_fused \
  python tools/prof_one.py > gpurun_out/r2_ncu.log 2>&1
ls -la gpurun_out
