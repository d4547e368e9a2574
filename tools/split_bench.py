#!/usr/bin/env python
"""Round-1b experiments on the fused sweep: split (arrive/wait) CTA barrier (tiles 5xxx), decoupled
levels (tiles 7xxx) and the trapezoid skip (tiles 9xxx) vs the default kernel.
Checks bit-equality of the variants on a ragged grid, then times one pass at 512^3 (fp64 and fp32,
strict and contracted arithmetic).  Measurement tool, not part of the product path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import girih_b200 as G  # noqa: E402


def parity():
    for dt, tiles in ((np.float64, (0, 5408, 5216, 7408, 7216, 9408, 9216)), (np.float32, (0, 5216, 5408, 7216, 7408, 9216, 9408))):
        ref = None
        for tile in tiles:
            for contract in (0, 1):
                pb = G.make_problem(1, (150, 71, 23), dt)
                s = G.GpuStepper.for_problem(pb)
                s.set_option("variant", 2)
                s.set_option("tile", tile)
                s.set_option("contract", contract)
                s.run_fused(13, 4)
                s.download(pb.U1, pb.U2)
                s.close()
                key = (contract,)
                ref = ref or {}
                if key not in ref:
                    ref[key] = (pb.U1.tobytes(), pb.U2.tobytes())
                ok = ref[key] == (pb.U1.tobytes(), pb.U2.tobytes())
                print(f"parity {np.dtype(dt).name} tile={tile} contract={contract}: {'OK' if ok else 'MISMATCH'}", flush=True)


def bench(n=512):
    for dt, tiles in ((np.float64, (0, 5408, 7408, 9408, 9216)), (np.float32, (0, 5216, 7216, 9216, 9408))):
        t0 = time.time()
        pb = G.make_problem(1, (n, n, n), dt)
        s = G.GpuStepper.for_problem(pb)
        print(f"setup {np.dtype(dt).name}: {time.time() - t0:.1f}s", flush=True)
        for contract in (0, 1):
            s.set_option("contract", contract)
            for T in (4, 3, 2):
                for tile in tiles:
                    s.set_option("tile", tile)
                    ms = min(s.time_pass(T, 10) for _ in range(2))
                    print(f"k1 {np.dtype(dt).name} n={n} T={T} tile={tile:4d} contract={contract}: {ms:7.3f} ms/pass "
                          f"{n ** 3 * T / ms / 1e6:7.1f} GLUP/s", flush=True)
        s.close()


if __name__ == "__main__":
    parity()
    bench()
