#!/usr/bin/env python
"""Box operator (slot 7) at 512^3: cp.async ring (tile 108) against its lean form (tile 208), alternating and repeated so that
neither sits in the power controller's start-up transient.  Measurement tool."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import girih_b200 as G  # noqa: E402
from bench import ClockSampler  # noqa: E402

n = 512
for dn, dt in (("f64", np.float64), ("f32", np.float32)):
    pb = G.make_problem(7, (n, n, n), dt)
    s = G.GpuStepper.for_problem(pb)
    del pb
    for tile in (108, 208, 108, 208, 0):
        s.set_option("tile", tile)
        sampler = ClockSampler(0)
        sampler.start()
        per = []
        for _ in range(4):
            s.run_single(200)
            per.append(n ** 3 * 200 / s.elapsed_ms()["total"] / 1e6)
        clk = sampler.stop()
        print(f"k7 {dn} n={n} tile={tile:3d}: " + " ".join(f"{p:6.1f}" for p in per) + f" GLUP/s   sm {clk.get('sm_mhz')} MHz, max {clk.get('power_w_max')} W, {clk.get('reasons')}", flush=True)
    s.close()
