import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import girih_b200 as G
from oracle import girih_oracle as O
def run(kernel, st, dt, nsteps, tfuse, opts=()):
    pb = G.make_problem(kernel, st, dt)
    s = G.GpuStepper.for_problem(pb)
    for k, v in opts: s.set_option(k, v)
    s.run_fused(nsteps, tfuse)
    s.download(pb.U1, pb.U2)
    s.close()
    ob = O.make_problem(kernel, st, dt)
    O.run_steps(ob, nsteps)
    bad = np.argwhere(pb.U1 != ob.U1)
    print(kernel, st, dt.__name__, nsteps, tfuse, opts, "mismatch", len(bad), end=" ")
    if len(bad):
        print("z", bad[:,0].min(), bad[:,0].max(), "y", bad[:,1].min(), bad[:,1].max(), "x", bad[:,2].min(), bad[:,2].max(), "first", tuple(bad[0]))
    else: print()
for rep in range(2):
    for tf in (1, 2, 3, 4):
        for st, n in (((150, 71, 23), 9), ((61, 34, 9), 12)):
            run(1, st, np.float64, n, tf)
    for tf in (1,3,4):
        run(1, (70,75,29), np.float64, 8, tf, (("tile",408),("zchunk",3)))
