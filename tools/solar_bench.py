#!/usr/bin/env python
"""Solar slot (table index 6): every schedule variant (option tile 1..5) and a few z-chunk lengths, device-timed per time
step (one H + one E launch), as algorithmic HBM GB/s (104 reals per cell and step) and as the traffic of the two-phase
in-place schedule (128 reals); one parity check per variant against the CPU oracle on a small domain."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import girih_b200 as G
from bench import measured_peaks


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    peak, _ = measured_peaks()
    from oracle import girih_oracle as O
    out = []
    for dt in (np.float64, np.float32):
        es = np.dtype(dt).itemsize
        # parity of every variant first (small domain, odd sizes)
        for tile in (1, 2, 3, 4, 5):
            pb = G.make_problem(6, (70, 19, 23), dt)
            ob = O.make_problem(6, (70, 19, 23), dt)
            s = G.GpuStepper.for_problem(pb)
            s.set_option("tile", tile)
            s.run_single(3)
            s.download(pb.U1, None)
            s.close()
            O.run_steps(ob, 3)
            assert pb.U1.tobytes() == ob.U1.tobytes(), (dt, tile)
        pb = G.make_problem(6, (n, n, n), dt)
        s = G.GpuStepper.for_problem(pb)
        del pb
        cells = float(n) ** 3
        for tile, zc in [(t, 0) for t in (1, 2, 3, 4, 5)] + [(2, 8), (2, 16), (2, 64), (3, 64), (2, 1000)]:
            s.set_option("tile", tile)
            s.set_option("zchunk", zc)
            s.time_pass(1, reps=2)
            ms = s.time_pass(1, reps=20)
            rec = {"dtype": np.dtype(dt).name, "n": n, "tile": tile, "zchunk": zc, "ms_per_step": ms,
                   "glups": cells / ms / 1e6, "alg_gbs": 104 * es * cells / ms / 1e6, "alg_frac": 104 * es * cells / ms / 1e6 / peak,
                   "two_phase_gbs": 128 * es * cells / ms / 1e6, "two_phase_frac": 128 * es * cells / ms / 1e6 / peak}
            out.append(rec)
            print("solar %s n=%d tile=%d zc=%4d: %8.3f ms/step %7.3f GLUP/s  alg %7.1f GB/s (%.3f)  two-phase %7.1f GB/s (%.3f)" % (
                rec["dtype"], n, tile, zc, ms, rec["glups"], rec["alg_gbs"], rec["alg_frac"], rec["two_phase_gbs"], rec["two_phase_frac"]), flush=True)
        s.close()
    json.dump(out, open("gpurun_out/r2_solar_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
