#!/bin/bash
# lean box kernel (tiles 208 / 216) against the round-2 default: parity on hardware, then sustained GLUP/s at 512^3
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "box_kernel_tile_seams" > gpurun_out/r2_box_lean_tests.log 2>&1
tail -n 2 gpurun_out/r2_box_lean_tests.log
timeout 300 python tools/r2_sustained_tiles.py 512 7 > gpurun_out/r2_box_lean.log 2>&1
cat gpurun_out/r2_box_lean.log
