#!/usr/bin/env python
"""Debug helper: run a case on the GPU and on the oracle, report where they differ."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import girih_b200 as G
from oracle import girih_oracle as O
ap = argparse.ArgumentParser()
ap.add_argument("--kernel", type=int, default=1)
ap.add_argument("--st", default="64,48,40")
ap.add_argument("--dtype", default="f64")
ap.add_argument("--nsteps", type=int, default=1)
ap.add_argument("--tfuse", type=int, default=1)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--zchunk", type=int, default=0)
ap.add_argument("--contract", type=int, default=0)
a = ap.parse_args()
st = tuple(int(x) for x in a.st.split(","))
dt = np.float64 if a.dtype == "f64" else np.float32
pb = G.make_problem(a.kernel, st, dt)
s = G.GpuStepper.for_problem(pb)
s.set_option("contract", a.contract); s.set_option("variant", a.variant); s.set_option("tile", a.tile); s.set_option("zchunk", a.zchunk)
s.run_fused(a.nsteps, a.tfuse)
s.download(pb.U1, pb.U2)
ob = O.make_problem(a.kernel, st, dt)
O.run_steps(ob, a.nsteps, contract=bool(a.contract))
for name, g, o in (("U1", pb.U1, ob.U1), ("U2", pb.U2, ob.U2)):
    bad = np.argwhere(g != o)
    print(name, "mismatches:", len(bad), "of", g.size)
    if len(bad):
        print("  z range", bad[:, 0].min(), bad[:, 0].max(), " y range", bad[:, 1].min(), bad[:, 1].max(),
              " x range", bad[:, 2].min(), bad[:, 2].max())
        print("  distinct z:", np.unique(bad[:, 0])[:20], " distinct y:", np.unique(bad[:, 1])[:40])
        print("  distinct x:", np.unique(bad[:, 2])[:80])
        for b in bad[:6]:
            print("  ", tuple(b), g[tuple(b)], o[tuple(b)])
