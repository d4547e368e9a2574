#!/bin/bash
# the driver's round-end sequence on one GPU: pytest -m gpu, smoke(), bench.py with its flags
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_last_pytest.log 2>&1
tail -n 3 gpurun_out/r2_last_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_last_smoke.log 2>&1
tail -n 2 gpurun_out/r2_last_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_last_bench1.json 2> gpurun_out/r2_last_bench1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_last_bench1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["parity_check"]["bit_exact"], d["clocks"]["reasons"])
for c in d["configs"]:
    print(" ", c.get("config"), round(c.get("glups", 0), 1), round(c.get("pass_hbm_frac", 0), 3), c.get("error"))
PY
