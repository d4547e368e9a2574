#!/bin/bash
# exact-tile sweep: parity on the GPU, per-pass timing at 512^3 against the overlapped-tile kernel, ncu capture
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exact_tiles" 2>&1 | tail -5 > gpurun_out/r2x_pytest.log
timeout 300 python tools/kbench.py --kernels 1 --dtypes f64 --tfuse 3,4 --tiles 0,10408,10216 --variants 2 > gpurun_out/r2x_kbench.log 2>&1
timeout 300 python tools/kbench.py --kernels 1 --dtypes f64 --tfuse 4 --tiles 0,10408,10216 --variants 2 --contract 1 >> gpurun_out/r2x_kbench.log 2>&1
if [ "${NCU:-1}" = 1 ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_r1x -s 2 -c 1 -o gpurun_out/r2x_fused \
  python tools/kbench.py --kernels 1 --dtypes f64 --tfuse 4 --tiles ${NCU_TILE:-10408} --variants 2 --reps 3 > gpurun_out/r2x_ncu.log 2>&1
fi
ls -la gpurun_out | tail -8
