#!/usr/bin/env python
"""Slot 0 (25-point, 2nd order in time) at 768^3: burst (10 steps after idle) and sustained (200 steps, repeated)
GLUP/s per tile option, with the SM clock sampled underneath.  Measurement tool."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import girih_b200 as G  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
tiles = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,208,216,116").split(",")]
for dn, dt in (("f64", np.float64), ("f32", np.float32)):
    pb = G.make_problem(0, (n, n, n), dt)
    s = G.GpuStepper.for_problem(pb)
    del pb
    for tile in tiles:
        s.set_option("tile", tile)
        time.sleep(3)
        s.run_single(10)
        burst = n ** 3 * 10 / s.elapsed_ms()["total"] / 1e6
        ms, reps = 0.0, 3
        for _ in range(reps):
            s.run_single(200)
            ms += s.elapsed_ms()["total"]
        clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True).stdout.strip()
        print(f"k0 {dn} n={n} tile={tile:3d}: burst {burst:7.1f} GLUP/s   sustained(3x200 steps) {n ** 3 * 200 * reps / ms / 1e6:7.1f} GLUP/s   "
              f"clock/power after: {clk}", flush=True)
    s.close()
