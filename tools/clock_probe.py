#!/usr/bin/env python
"""Sample SM clock / power / throttle reasons while a kernel runs for a few seconds."""
import argparse, os, sys, statistics
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
import girih_b200 as G
from bench import ClockSampler
ap = argparse.ArgumentParser()
ap.add_argument("--kernel", type=int, default=0); ap.add_argument("--n", type=int, default=768)
ap.add_argument("--dtype", default="f64"); ap.add_argument("--tfuse", type=int, default=1)
ap.add_argument("--chunks", type=int, default=8); ap.add_argument("--reps", type=int, default=100)
a = ap.parse_args()
pb = G.make_problem(a.kernel, (a.n,) * 3, np.float64 if a.dtype == "f64" else np.float32)
s = G.GpuStepper.for_problem(pb)
smp = ClockSampler(0); smp.start()
for c in range(a.chunks):
    ms = s.time_pass(a.tfuse, a.reps)
    print(f"chunk {c}: {ms:.3f} ms/pass", flush=True)
r = smp.stop()
print(r)
import collections
sm = [float(l.split(',')[1]) for l in smp.lines if len(l.split(',')) > 8]
pw = [float(l.split(',')[3]) for l in smp.lines if len(l.split(',')) > 8]
print("clock samples:", collections.Counter(int(x) for x in sm).most_common(6))
print("power max/median:", max(pw), statistics.median(pw))
