#!/usr/bin/env python
"""Sustained (3 x 100 single steps) GLUP/s per tile option for the single-step operators.  Measurement tool."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import girih_b200 as G  # noqa: E402

cases = {4: [0, 8, 16], 7: [0, 116, 208, 216], 1: [0, 108, 208, 404, 408], 5: [0], 0: [0, 108]}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 7, 1]
for k in ks:
    for dn, dt in (("f64", np.float64), ("f32", np.float32)):
        pb = G.make_problem(k, (n, n, n), dt)
        s = G.GpuStepper.for_problem(pb)
        del pb
        for tile in cases[k]:
            s.set_option("tile", tile)
            s.run_single(4)
            ms, reps = 0.0, 3
            for _ in range(reps):
                s.run_single(100)
                ms += s.elapsed_ms()["total"]
            print(f"k{k} {dn} n={n} tile={tile:3d}: sustained {n ** 3 * 100 * reps / ms / 1e6:7.1f} GLUP/s", flush=True)
        s.close()
