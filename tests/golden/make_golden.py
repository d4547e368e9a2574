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
    main(fast=(len(sys.argv) > 1 and sys.argv[1] == "fma"))
