#!/usr/bin/env python
"""Slot 0 fp64 at 768^3 under the power cap: 8-row (tile 108) against 16-row (tile 116, the default) cp.async kernels,
alternating, 5 x 200 steps each without pauses, clocks and power sampled DURING the runs.  Measurement tool."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import girih_b200 as G  # noqa: E402
from bench import ClockSampler  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
for dn, dt in (("f64", np.float64), ("f32", np.float32)):
    pb = G.make_problem(0, (n, n, n), dt)
    s = G.GpuStepper.for_problem(pb)
    del pb
    for tile in (108, 116, 0, 108, 116):
        s.set_option("tile", tile)
        sampler = ClockSampler(0)
        sampler.start()
        per = []
        for _ in range(5):
            s.run_single(200)
            per.append(n ** 3 * 200 / s.elapsed_ms()["total"] / 1e6)
        clk = sampler.stop()
        print(f"k0 {dn} n={n} tile={tile:3d}: " + " ".join(f"{p:6.1f}" for p in per) + f" GLUP/s   sm {clk.get('sm_mhz')} MHz, max {clk.get('power_w_max')} W, {clk.get('reasons')}", flush=True)
    s.close()
