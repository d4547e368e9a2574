#!/bin/bash
# one GPU: the whole GPU suite (incl. the BASELINE-size goldens), smoke, the bench line with its configs block
set -u
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2f_pytest.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1
( time python bench.py ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
ls -la gpurun_out | tail -4
