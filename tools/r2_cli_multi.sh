#!/bin/bash
# mwd_kernel --npz N on N GPUs: --verify 1 with the default (halo copy) and with --gpu-copy 0, then performance runs
set -u
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r2m_cli_${N}gpu.log
: > $L
for flags in "" "--gpu-copy 0" "--gpu-push 1"; do
  echo "== verify ts 2 k1 fp64 $flags" >> $L
  timeout 120 ./build_dp/mwd_kernel --nx 256 --ny 256 --nz 256 --nt 50 --target-ts 2 --target-kernel 1 --t-dim 3 --verify 1 --npz $N $flags --verbose 0 >> $L 2>&1
done
echo "== verify ts 1 k0 fp32" >> $L
timeout 120 ./build/mwd_kernel --nx 256 --ny 256 --nz 256 --nt 20 --target-ts 1 --target-kernel 0 --verify 1 --npz $N --verbose 0 >> $L 2>&1
for flags in "" "--gpu-copy 0"; do
  echo "== perf ts 2 k1 fp64 512x512x$((512*N)) $flags" >> $L
  timeout 200 ./build_dp/mwd_kernel --nx 512 --ny 512 --nz $((512*N)) --nt 500 --target-ts 2 --target-kernel 1 --t-dim 7 --npz $N $flags --n-tests 3 2>&1 | grep -E "GPU true GLUP|Total RANK0 MStencil/s MAX|RANK0 Computation|RANK0 Communication|RANK0 Waiting|RANK0 Total" >> $L
done
echo "== perf ts 1 k0 fp32 1024^3" >> $L
timeout 200 ./build/mwd_kernel --nx 1024 --ny 1024 --nz 1024 --nt 200 --target-ts 1 --target-kernel 0 --npz $N --n-tests 3 2>&1 | grep -E "GPU true GLUP|RANK0 GStencil/s    MAX|RANK0 Computation|RANK0 Communication|RANK0 Waiting|RANK0 Total" >> $L
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "halo_copy or halo_push or z_slabs or uneven" 2>&1 | tail -3 >> $L
cat $L
