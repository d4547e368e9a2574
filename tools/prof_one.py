#!/usr/bin/env python
"""One or a few launches of a chosen operator / depth for ncu captures."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import girih_b200 as G
ap = argparse.ArgumentParser()
ap.add_argument("--kernel", type=int, default=1)
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--tfuse", type=int, default=4)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
pb = G.make_problem(a.kernel, (a.n,) * 3, np.float64 if a.dtype == "f64" else np.float32)
s = G.GpuStepper.for_problem(pb)
s.set_option("variant", a.variant)
s.set_option("tile", a.tile)
print(s.time_pass(a.tfuse, a.reps))
s.close()
