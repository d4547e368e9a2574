#!/usr/bin/env python
"""z-wavefront temporal blocking through L2 (girih_cuda.cu, run_zwave) for the operators without a fused-sweep kernel:
GLUP/s of run_fused over `steps` time steps for every (steps in flight W, planes per block B).  Measurement tool."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import girih_b200 as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernels", default="0")
    ap.add_argument("--dtypes", default="f32,f64")
    ap.add_argument("--n", type=int, default=768)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--waves", default="1,2,3,4")
    ap.add_argument("--blocks", default="4,8,16,32")
    a = ap.parse_args()
    n = a.n
    for k in [int(x) for x in a.kernels.split(",")]:
        for dn in a.dtypes.split(","):
            dt = np.float64 if dn == "f64" else np.float32
            pb = G.make_problem(k, (n, n, n), dt)
            s = G.GpuStepper.for_problem(pb)
            del pb
            for W in [int(x) for x in a.waves.split(",")]:
                for B in ([0] if W == 1 else [int(x) for x in a.blocks.split(",")]):
                    s.set_option("zwave", W)
                    s.set_option("zwave_block", B)
                    s.run_fused(a.steps, 0)
                    s.run_fused(a.steps, 0)
                    ms = s.elapsed_ms()["total"]
                    info = s.launch_info()
                    print(f"k{k} {dn} n={n} W={W} B={B:3d}: {ms / a.steps:8.4f} ms/step  {n ** 3 * a.steps / ms / 1e6:8.1f} GLUP/s  "
                          f"launches {info['kernels']}", flush=True)
            s.close()


if __name__ == "__main__":
    main()
