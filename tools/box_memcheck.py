import sys
sys.path.insert(0, ".")
import numpy as np
import girih_b200 as G
for dt in (np.float32, np.float64):
    for st, tile, zc in (((150, 37, 23), 208, 0), ((67, 9, 12), 216, 5), ((33, 17, 9), 0, 0)):
        pb = G.make_problem(7, st, dt)
        s = G.GpuStepper.for_problem(pb)
        s.set_option("tile", tile); s.set_option("zchunk", zc)
        s.run_single(3)
        s.close()
print("ok")
