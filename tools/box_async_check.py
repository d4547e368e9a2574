import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import girih_b200 as G
from oracle import girih_oracle as O
ok_all = True
for tile in (108, 116):
    for dt in (np.float32, np.float64):
        for st, zc in (((150, 37, 23), 0), ((64, 8, 5), 0), ((130, 20, 40), 7), ((300, 200, 70), 0)):
            pb = G.make_problem(7, st, dt); s = G.GpuStepper.for_problem(pb); s.set_option("tile", tile); s.set_option("zchunk", zc)
            s.run_single(5); s.download(pb.U1, pb.U2); s.close()
            ob = O.make_problem(7, st, dt); O.run_steps(ob, 5)
            ok = pb.U1.tobytes() == ob.U1.tobytes() and pb.U2.tobytes() == ob.U2.tobytes()
            ok_all &= ok
            if not ok: print("MISMATCH", tile, dt.__name__, st, zc)
print("box async parity:", "ALL OK" if ok_all else "FAILED")
