"""The CPU oracle (oracle/girih_oracle.c) against the committed reference outputs.

tests/golden/small.npz and checksums.json were produced by the UNMODIFIED reference steppers
(tests/golden/make_golden.py); equality is bit-for-bit, the reference's own criterion
(src/verification.c:842).
"""
import hashlib
import json
import os
import re

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = np.load(os.path.join(HERE, "golden", "small.npz"))
SUMS = json.load(open(os.path.join(HERE, "golden", "checksums.json")))
# the same cases from the reference built with FMA contraction (make_golden.py fma)
SMALL_FMA = np.load(os.path.join(HERE, "golden", "small_fma.npz"))
SUMS_FMA = json.load(open(os.path.join(HERE, "golden", "checksums_fma.json")))
KEY = re.compile(r"k(\d)_(\d+)x(\d+)x(\d+)_nt(\d+)_ts(\d)_td(\d)_(sp|dp)$")


def parse(key):
    m = KEY.match(key)
    k, nx, ny, nz, nt, ts, td = (int(g) for g in m.groups()[:7])
    return k, (nx, ny, nz), nt, ts, td, (np.float32 if m.group(8) == "sp" else np.float64)


def run_oracle(O, k, st, nt, ts, td, dt, contract=False):
    pb = O.make_problem(k, st, dt)
    if ts == 2:                      # diamond: nt rounded up, nt-1 steps executed
        nt = O.diamond_round_nt(nt, td)
        O.run_steps(pb, nt - 1, contract=contract)
    else:
        O.run_naive(pb, nt, contract=contract)
    return pb, nt


@pytest.mark.parametrize("key", sorted(k for k in SMALL.files if not k.endswith("_nteff")))
def test_small_golden(oracle, key):
    k, st, nt, ts, td, dt = parse(key)
    pb, nte = run_oracle(oracle, k, st, nt, ts, td, dt)
    assert nte == int(SMALL[key + "_nteff"]) or ts != 2
    gold = SMALL[key]
    assert gold.dtype == dt
    got = pb.interior()
    assert got.tobytes() == gold.tobytes()


@pytest.mark.parametrize("key", sorted(k for k in SMALL_FMA.files if not k.endswith("_nteff")))
def test_small_golden_contracted(oracle, key):
    """The oracle compiled with gcc's FMA contraction == the reference compiled the same way."""
    k, st, nt, ts, td, dt = parse(key)
    pb, _ = run_oracle(oracle, k, st, nt, ts, td, dt, contract=True)
    assert pb.interior().tobytes() == SMALL_FMA[key].tobytes()
    assert SMALL_FMA[key].tobytes() != SMALL[key].tobytes()      # and it is a different rounding


@pytest.mark.parametrize("key", sorted(SUMS_FMA))
def test_checksum_golden_contracted(oracle, key):
    k, st, nt, ts, td, dt = parse(key)
    pb, nte = run_oracle(oracle, k, st, nt, ts, td, dt, contract=True)
    assert nte == SUMS_FMA[key]["nt_effective"]
    got = np.ascontiguousarray(pb.interior())
    assert hashlib.sha256(got.tobytes()).hexdigest() == SUMS_FMA[key]["sha256"]


@pytest.mark.parametrize("key", sorted(SUMS))
def test_checksum_golden(oracle, key):
    k, st, nt, ts, td, dt = parse(key)
    pb, nte = run_oracle(oracle, k, st, nt, ts, td, dt)
    assert nte == SUMS[key]["nt_effective"]
    got = np.ascontiguousarray(pb.interior())
    assert hashlib.sha256(got.tobytes()).hexdigest() == SUMS[key]["sha256"]
    assert float(np.abs(got).max()) == SUMS[key]["max_abs"]


def test_frame_untouched_and_padding_zero(oracle):
    """Dirichlet frame and x padding are never written (xb=r..xe=nx+r, nb_naive_ts.c:189)."""
    O = oracle
    pb = O.make_problem(1, (13, 9, 7), np.float64)
    before = pb.U1.copy()
    O.run_naive(pb, 4)
    mask = np.ones(pb.U1.shape, bool)
    mask[1:-1, 1:-1, 1:14] = False
    assert np.array_equal(pb.U1[mask], before[mask])
    assert pb.shape[0] == 16 and np.all(pb.U1[:, :, 15] == 0)
    assert np.all(pb.U1[:, :, 0] == 100.1) and np.all(pb.U1[:, :, 14] == 100.1)


def test_decomposed_fill_matches_global(oracle):
    """A z-slab's fill with gb offset equals the same planes of the global fill
    (src/utils.c:630-646 uses global indices), which is what makes the serial global run the
    oracle for any decomposition (src/verification.c:52-312)."""
    O = oracle
    g = O.make_problem(0, (12, 10, 16), np.float32)
    r = g.r
    for nparts in (2, 3, 4):
        for c in range(nparts):
            lnz, gbz = O.decompose(16, nparts, c)
            s = O.make_problem(0, (12, 10, lnz), np.float32, gstencil=(12, 10, 16), gb=(0, 0, gbz),
                               first=(1, 1, int(c == 0)), last=(1, 1, int(c == nparts - 1)))
            assert np.array_equal(s.U1, g.U1[gbz:gbz + lnz + 2 * r])
            assert np.array_equal(s.U3, g.U3[gbz:gbz + lnz + 2 * r])


# ---- the solar slot (table index 6): tests/golden/solar.json / solar_small.npz, outputs of the reference's own solar
# kernel through its ts 0 / ts 1 steppers (tests/golden/make_golden.py solar) --------------------------------------------
SOLAR_SUMS = json.load(open(os.path.join(HERE, "golden", "solar.json")))
SOLAR_SMALL = np.load(os.path.join(HERE, "golden", "solar_small.npz"))


def _solar_key(name):
    # k6_<nx>x<ny>x<nz>_nt<nt>_ts<ts>_td0_<sp|dp>
    parts = name.split("_")
    st = tuple(int(v) for v in parts[1].split("x"))
    return st, int(parts[2][2:]), int(parts[3][2:]), (np.float32 if parts[5] == "sp" else np.float64)


@pytest.mark.parametrize("name", sorted(SOLAR_SUMS))
def test_solar_oracle_matches_reference_checksums(oracle, name):
    import hashlib
    st, nt, ts, dt = _solar_key(name)
    pb = oracle.make_problem(6, st, dt)
    oracle.run_naive(pb, nt)
    assert hashlib.sha256(pb.U1.tobytes()).hexdigest() == SOLAR_SUMS[name]["sha256"]
    assert float(np.abs(pb.U1).max()) == SOLAR_SUMS[name]["max_abs"]


@pytest.mark.parametrize("name", sorted(SOLAR_SMALL.files))
def test_solar_oracle_matches_reference_arrays(oracle, name):
    st, nt, ts, dt = _solar_key(name)
    pb = oracle.make_problem(6, st, dt)
    assert pb.U1.shape == SOLAR_SMALL[name].shape == (12, st[2] + 2, st[1] + 2, st[0] + 2, 2)
    before = pb.U1.copy()
    oracle.run_naive(pb, nt)
    assert pb.U1.tobytes() == SOLAR_SMALL[name].tobytes()
    # the frame is never written, every interior value of every field moved
    assert np.array_equal(pb.U1[:, 0], before[:, 0]) and np.array_equal(pb.U1[:, :, :, -1], before[:, :, :, -1])
    assert not np.any(pb.U1[:, 1:-1, 1:-1, 1:-1] == before[:, 1:-1, 1:-1, 1:-1])
