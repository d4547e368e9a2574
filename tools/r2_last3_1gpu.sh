#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2_last3_bench1.json 2> gpurun_out/r2_last3_bench1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_last3_bench1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["parity_check"]["bit_exact"])
for c in d["configs"]:
    print(" ", c.get("config"), round(c.get("glups", 0), 1), round(c.get("glups_settled", 0), 1), c.get("settle_runs"), c.get("settled_sm_mhz"), c.get("error"))
PY
