#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference -> oracle/_ref via
`make -C oracle ref`):   python tests/golden/make_golden.py

For every case the reference's own time stepper (ts 0 "Spatial Blocking" = src/kernels/
nb_naive_ts.c, or ts 2 "Diamond" = src/kernels/diamond_ts.c) is executed by
oracle/_ref/ref_dump_{sp,dp} on GIRIH's deterministic initial data (src/utils.c:605-697) and the
interior of p.U1 is stored:
  * small cases  -> tests/golden/small.npz   (full interior arrays, bit patterns)
  * larger cases -> tests/golden/checksums.json (sha256 over the interior bytes, [z][y][x] order)
The reference ships no golden vectors of its own; these are outputs of the
reference itself, which is what its --verify mode compares against bit for bit.

`python tests/golden/make_golden.py baseline` writes checksums_baseline.json: sha256 of the reference's output at
the sizes BASELINE.json's configs name (C1 256^3 x 100 with both steppers, C2 512^3 x 500 Diamond = the bench workload
itself, C3 768^3 x 200 25-point, C4 512^3 x 200 per-point coefficients), fp64, so that the GPU parity suite compares
the full-size runs with the reference and not with itself.  Takes several minutes of CPU time.

`python tests/golden/make_golden.py solar` writes solar.json / solar_small.npz: outputs of the reference's solar kernel
(table slot 6) through its ts 0 / ts 1 steppers.
`python tests/golden/make_golden.py fma` writes small_fma.npz / checksums_fma.json instead: the same
cases run by the reference built with FMA contraction (oracle/_ref/ref_dump_*_fast, gcc -O3 -mfma
-ffp-contract=fast) -- the fixtures for the library's "contract" option.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import girih_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (kernel, (nx,ny,nz), nt, ts, t_dim)
SMALL = [(k, (20, 12, 14), 6, 0, 0) for k in (0, 1, 2, 3, 4, 5, 7)] + [
    (1, (18, 16, 12), 10, 2, 1),     # diamond width 4, nt 10 -> 10
    (1, (16, 16, 14), 12, 2, 3),     # diamond width 8, nt 12 -> 18
    (5, (12, 8, 10), 8, 2, 1),
    (0, (16, 16, 12), 6, 2, 1),      # r=4: diamond width 16
    (3, (33, 9, 7), 5, 0, 0),        # ragged sizes, odd nt (nt+1 steps are executed)
]
LARGE = [(k, (64, 48, 40), 32, 0, 0) for k in (0, 1, 2, 3, 4, 5)] + [
    (1, (96, 64, 64), 50, 2, 3),
    (1, (128, 128, 128), 100, 0, 0),
    (0, (96, 96, 96), 40, 0, 0),
    (4, (72, 72, 72), 20, 0, 0),
]


# BASELINE.json configs at full size (fp64): (kernel, stencil, nt, ts, t_dim)
BASELINE = [
    (1, (256, 256, 256), 100, 0, 0),      # C1, spatial blocking
    (1, (256, 256, 256), 100, 2, 7),      # C1, Diamond (README flags: t_dim 7 -> nt 114)
    (1, (512, 512, 512), 500, 2, 7),      # C2 = bench.py's workload (nt 514, 513 steps)
    (0, (768, 768, 768), 200, 2, 1),      # C3, 25-point constant coefficients (Diamond t_dim 1 -> nt 202)
    (2, (512, 512, 512), 200, 2, 3),      # C4, 7-point variable coefficients
    (5, (512, 512, 512), 200, 2, 3),      # C4, 7-point variable coefficients without symmetry
]


def main_baseline():
    import time
    assert O.have_ref(), "build the reference first: make -C oracle ref"
    path = os.path.join(HERE, "checksums_baseline.json")
    sums = json.load(open(path)) if os.path.exists(path) else {}
    nthr = len(os.sched_getaffinity(0))
    for (k, st, nt, ts, td) in BASELINE:
        kk = key(k, st, nt, ts, td, "dp")
        if kk in sums:
            continue
        t0 = time.time()
        U1, r, nte = O.ref_dump(k, st, nt, np.float64, ts, ts_extra(ts, td), threads=nthr, timeout=7200)
        nx, ny, nz = st
        it = np.ascontiguousarray(U1[r:r + nz, r:r + ny, r:r + nx])
        sums[kk] = {"sha256": hashlib.sha256(it.tobytes()).hexdigest(), "nt_effective": nte,
                    "max_abs": float(np.abs(it[np.isfinite(it)]).max()), "n_nonfinite": int((~np.isfinite(it)).sum())}
        print(kk, sums[kk], f"{time.time() - t0:.0f} s", flush=True)
        with open(path, "w") as f:
            json.dump(sums, f, indent=1, sort_keys=True)


# the solar slot (table index 6): (stencil, nt, ts); the reference needs an explicit --thread-group-size for ts 0 / 1
# (its default -1 reaches `omp parallel num_threads(-1)`, src/kernels/solar_spt_blk.ic:65-66)
SOLAR = [((24, 20, 18), 6, 0), ((17, 9, 11), 3, 0), ((40, 33, 29), 8, 1), ((64, 48, 40), 12, 0), ((130, 5, 40), 2, 0)]
SOLAR_SMALL = ((9, 7, 6), 2, 0)


def main_solar():
    """solar.json: sha256 over the reference's whole array (12 fields x [z][y][x] x (re, im), frame included) after the run;
    solar_small.npz: one tiny case in full."""
    assert O.have_ref(), "build the reference first: make -C oracle ref"
    sums, small = {}, {}
    for dt, name in ((np.float32, "sp"), (np.float64, "dp")):
        for st, nt, ts in SOLAR:
            U1, r, nte = O.ref_dump(6, st, nt, dt, ts, threads=2)
            sums[key(6, st, nt, ts, 0, name)] = {"sha256": hashlib.sha256(U1.tobytes()).hexdigest(),
                                                 "max_abs": float(np.abs(U1).max())}
        st, nt, ts = SOLAR_SMALL
        U1, r, nte = O.ref_dump(6, st, nt, dt, ts, threads=2)
        small[key(6, st, nt, ts, 0, name)] = U1
    with open(os.path.join(HERE, "solar.json"), "w") as f:
        json.dump(sums, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "solar_small.npz"), **small)
    print(len(sums), "solar checksums,", len(small), "small cases")


def ts_extra(ts, t_dim):
    return ("--t-dim", t_dim, "--thread-group-size", 1, "--num-wavefronts", 1) if ts == 2 else ()


def key(k, st, nt, ts, t_dim, dt):
    return f"k{k}_{st[0]}x{st[1]}x{st[2]}_nt{nt}_ts{ts}_td{t_dim}_{dt}"


def main(fast=False):
    assert O.have_ref(), "build the reference first: make -C oracle ref"
    small, sums = {}, {}
    tag = "_fma" if fast else ""
    for dt, name in ((np.float32, "sp"), (np.float64, "dp")):
        for (k, st, nt, ts, td) in SMALL:
            U1, r, nte = O.ref_dump(k, st, nt, dt, ts, ts_extra(ts, td), threads=1 if ts == 2 else 2, fast=fast)
            nx, ny, nz = st
            small[key(k, st, nt, ts, td, name)] = U1[r:r + nz, r:r + ny, r:r + nx].copy()
            small[key(k, st, nt, ts, td, name) + "_nteff"] = np.int32(nte)
        for (k, st, nt, ts, td) in LARGE:
            if dt == np.float32 and k == 0 and nt > 60:
                continue
            U1, r, nte = O.ref_dump(k, st, nt, dt, ts, ts_extra(ts, td), threads=4 if ts == 2 else 8, fast=fast)
            nx, ny, nz = st
            it = np.ascontiguousarray(U1[r:r + nz, r:r + ny, r:r + nx])
            sums[key(k, st, nt, ts, td, name)] = {
                "sha256": hashlib.sha256(it.tobytes()).hexdigest(), "nt_effective": nte,
                "max_abs": float(np.abs(it).max())}
    np.savez_compressed(os.path.join(HERE, "small%s.npz" % tag), **small)
    with open(os.path.join(HERE, "checksums%s.json" % tag), "w") as f:
        json.dump(sums, f, indent=1, sort_keys=True)
    print(len(small) // 2, "small cases,", len(sums), "checksums")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "solar":
        main_solar()
    elif len(sys.argv) > 1 and sys.argv[1] == "baseline":
        main_baseline()
    else:
        main(fast=(len(sys.argv) > 1 and sys.argv[1] == "fma"))
