"""The oracle against the real reference run live (only where oracle/_ref was built, i.e. in the
build container, or on the GPU box where the prebuilt binaries travel)."""
import numpy as np
import pytest


def _need_ref(O):
    if not O.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("kernel", [0, 1, 2, 3, 4, 5, 7])
def test_ts0_dump(oracle, kernel, dt):
    _need_ref(oracle)
    st = (17, 11, 13)
    U1, r, nte = oracle.ref_dump(kernel, st, 7, dt, ts=0)
    pb = oracle.make_problem(kernel, st, dt)
    oracle.run_naive(pb, 7)
    assert U1.tobytes() == pb.U1.tobytes()


@pytest.mark.parametrize("kernel,t_dim,st", [(1, 1, (16, 8, 9)), (1, 3, (10, 16, 11)),
                                             (2, 1, (9, 12, 8)), (0, 1, (10, 16, 13))])
def test_ts2_dump(oracle, kernel, t_dim, st):
    _need_ref(oracle)
    U1, r, nte = oracle.ref_dump(kernel, st, 9, np.float64, ts=2, threads=1,
                                 extra=("--t-dim", t_dim, "--thread-group-size", 1, "--num-wavefronts", 1))
    assert nte == oracle.diamond_round_nt(9, t_dim)
    pb = oracle.make_problem(kernel, st, np.float64)
    oracle.run_steps(pb, nte - 1)
    assert U1.tobytes() == pb.U1.tobytes()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("st,nt,ts", [((17, 11, 13), 5, 0), ((33, 6, 5), 4, 1), ((8, 70, 9), 2, 0)])
def test_solar_dump(oracle, dt, st, nt, ts):
    """table slot 6 through the reference's own solar kernel (src/kernels/solar_spt_blk.ic), whole array compared"""
    _need_ref(oracle)
    U1, r, nte = oracle.ref_dump(6, st, nt, dt, ts=ts)
    pb = oracle.make_problem(6, st, dt)
    oracle.run_naive(pb, nt)
    assert U1.tobytes() == pb.U1.tobytes()


def test_reference_own_verify_passes(oracle):
    """The reference binary built here satisfies its own bit-exact verifier."""
    _need_ref(oracle)
    out = oracle.ref_cli(np.float64, ["--nx", 32, "--ny", 24, "--nz", 20, "--nt", 10, "--target-kernel", 1,
                                      "--target-ts", 0, "--verify", 1, "--verbose", 0,
                                      "--thread-group-size", 2], threads=2)
    assert "eMax:0.000e+00|eL1:0.000e+00-PASSED" in out


def _need_fast_ref(O):
    import os
    _need_ref(O)
    if not os.path.exists(os.path.join(O.REF_DIR, "ref_dump_dp_fast")):
        pytest.skip("oracle/_ref/ref_dump_*_fast not built")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("kernel", [0, 1, 2, 3, 4, 5, 7])
def test_contracted_oracle_equals_reference_built_with_fma(oracle, kernel, dt):
    """oracle_*_{dpf,spf} (gcc -mfma -ffp-contract=fast) == the reference compiled with the same flags,
    for the single-step stepper and (radius 1) the diamond stepper; and both differ from strict."""
    _need_fast_ref(oracle)
    st = (17, 12, 13)
    U1, r, nte = oracle.ref_dump(kernel, st, 7, dt, ts=0, fast=True)
    pb = oracle.make_problem(kernel, st, dt)
    oracle.run_naive(pb, 7, contract=True)
    assert U1.tobytes() == pb.U1.tobytes()
    strict = oracle.make_problem(kernel, st, dt)
    oracle.run_naive(strict, 7)
    assert strict.U1.tobytes() != pb.U1.tobytes()
    if r == 1:
        st = (17, 16, 13)
        U1, r, nte = oracle.ref_dump(kernel, st, 9, dt, ts=2, threads=1, fast=True,
                                     extra=("--t-dim", 1, "--thread-group-size", 1, "--num-wavefronts", 1))
        pb = oracle.make_problem(kernel, st, dt)
        oracle.run_steps(pb, nte - 1, contract=True)
        assert U1.tobytes() == pb.U1.tobytes()
