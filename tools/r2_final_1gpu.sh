#!/bin/bash
# last hardware call of round 2: the driver's own bench command, ncu + memcheck of the solar kernels at their default
mkdir -p gpurun_out
timeout 420 python bench.py > gpurun_out/r2_final_bench1.json 2> gpurun_out/r2_final_bench1.err
tail -c 600 gpurun_out/r2_final_bench1.json; echo
cat > /tmp/solar_one.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import girih_b200 as G
n = int(sys.argv[1]); dt = np.float64 if sys.argv[2] == "f64" else np.float32
pb = G.make_problem(6, (n, n - 3, n // 2 + 1), dt)
s = G.GpuStepper.for_problem(pb)
s.run_single(2)
s.step_box(1, (2, 3, 1, n - 1, n - 5, n // 2))
s.close()
print("ok")
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_solar -s 2 -c 2 -o gpurun_out/r2_solar_default -f python /tmp/solar_one.py 192 f64 > gpurun_out/r2_solar_default_ncu.log 2>&1
tail -n 1 gpurun_out/r2_solar_default_ncu.log
timeout 200 compute-sanitizer --tool memcheck python /tmp/solar_one.py 40 f64 > gpurun_out/r2_solar_memcheck.log 2>&1
timeout 200 compute-sanitizer --tool memcheck python /tmp/solar_one.py 37 f32 >> gpurun_out/r2_solar_memcheck.log 2>&1
grep -E "ERROR SUMMARY|^ok" gpurun_out/r2_solar_memcheck.log
