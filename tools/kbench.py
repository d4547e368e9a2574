#!/usr/bin/env python
"""Kernel sweep on one GPU: per-pass device time of every operator / precision / fusion depth /
tile shape at a given grid size.  Prints one line per configuration; used to pick defaults and to
fill the tables in DESIGN.md.  (Measurement tool, not part of the product path.)"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import girih_b200 as G  # noqa: E402

PEAK = 6538.9
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--kernels", default="1")
    ap.add_argument("--dtypes", default="f64")
    ap.add_argument("--tfuse", default="1,2,3,4")
    ap.add_argument("--tiles", default="0")
    ap.add_argument("--zchunks", default="0")
    ap.add_argument("--variants", default="0")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--contract", type=int, default=0)
    a = ap.parse_args()
    n = a.n
    for k in [int(x) for x in a.kernels.split(",")]:
        kd = G.kernel_info(k)
        for dn in a.dtypes.split(","):
            dt = np.float64 if dn == "f64" else np.float32
            pb = G.make_problem(k, (n, n, n), dt)
            s = G.GpuStepper.for_problem(pb)
            s.set_option("contract", a.contract)
            for variant in [int(x) for x in a.variants.split(",")]:
                s.set_option("variant", variant)
                for tile in [int(x) for x in a.tiles.split(",")]:
                    s.set_option("tile", tile)
                    for zc in [int(x) for x in a.zchunks.split(",")]:
                        s.set_option("zchunk", zc)
                        for T in [int(x) for x in a.tfuse.split(",")]:
                            if T > kd.max_tfuse or (variant == 1 and T > 1):
                                continue
                            try:
                                ms = s.time_pass(T, a.reps)
                            except G.GirihError as e:
                                print(f"k{k} {dn} T={T} tile={tile} zc={zc} v={variant}: {e}")
                                continue
                            lups = n ** 3 * T
                            bytes_alg = kd.words_per_lup * np.dtype(dt).itemsize * n ** 3
                            gbs = bytes_alg / ms / 1e6
                            print(f"k{k} {dn} n={n} T={T} tile={tile:3d} zc={zc:4d} v={variant}: {ms:8.3f} ms/pass  "
                                  f"{lups / ms / 1e6:8.1f} GLUP/s  {gbs:7.1f} GB/s alg  frac {gbs / PEAK:5.3f}", flush=True)
            s.close()


if __name__ == "__main__":
    main()
