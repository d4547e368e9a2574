#!/bin/bash
# first hardware run of the solar slot: parity tests, schedule variants, ncu of the two phase kernels
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "solar or (cli_verify and 6)" > gpurun_out/r2_solar_tests.log 2>&1
tail -3 gpurun_out/r2_solar_tests.log
timeout 300 python tools/solar_bench.py 192 > gpurun_out/r2_solar_bench.log 2>&1
tail -22 gpurun_out/r2_solar_bench.log
cat > /tmp/solar_one.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import girih_b200 as G
pb = G.make_problem(6, (192, 192, 192), np.float64)
s = G.GpuStepper.for_problem(pb)
s.run_single(2)
s.close()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_solar -s 2 -c 2 -o gpurun_out/r2_solar -f python /tmp/solar_one.py > gpurun_out/r2_solar_ncu.log 2>&1
tail -2 gpurun_out/r2_solar_ncu.log
ls -la gpurun_out/r2_solar.ncu-rep
