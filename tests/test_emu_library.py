"""The complete C-ABI library (girih_b200/csrc/girih_cuda.cu: context, transfers, pass and exchange
schedules, steppers, options) compiled for the CPU SIMT emulator, driven through the same Python mirror
and the SAME test functions as the GPU parity suite (tests/test_gpu_parity.py).  Ranks of a z-slab run
are host threads; NCCL is an in-process mailbox (tests/cuda_emu/emu_nccl.cpp).

This is a checker for the host-side logic and the kernel logic together -- it never stands in for the
GPU run: the product loads libgirih_cuda.so only, and nothing under girih_b200/ knows the emulator
exists.  Stream/event ordering is not modelled (every operation completes in program order)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cuda_emu as E
import girih_b200 as G
from girih_b200 import lib as L
import test_gpu_parity as P

_emu = None


def emu_library():
    global _emu
    if _emu is None:
        E.lib()   # builds if needed
        _emu = L.declare(C.CDLL(E.LIB_PATH))
    return _emu


def emu_cli(dtype, args, timeout=3600, env=None):
    """girih_b200/host (the CLI) linked against the emulator library: tests/cuda_emu/_build/mwd_kernel_emu_*"""
    emu_library()
    exe = os.path.join(os.path.dirname(E.LIB_PATH), "mwd_kernel_emu_" + ("dp" if np.dtype(dtype).itemsize == 8 else "sp"))
    out = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout,
                         env=dict(os.environ, **(env or {})))
    return out.returncode, out.stdout, out.stderr


@pytest.fixture(autouse=True)
def _route_mirror_to_emulator(monkeypatch):
    monkeypatch.setattr(G.GpuStepper, "_load", staticmethod(emu_library))
    monkeypatch.setattr(G, "gpu_count", lambda: 5)
    monkeypatch.setattr(G, "run_reference_cli", emu_cli)
    monkeypatch.setenv("GIRIH_RUN_UNVALIDATED", "1")   # tests that wait for their first hardware run do run here
    monkeypatch.delenv("CUDA_EMU_SCHED", raising=False)


def test_every_abi_symbol_is_exported_by_the_emulator_build():
    lib = emu_library()
    for name in L.ABI_SYMBOLS:
        assert hasattr(lib, name), name


# the GPU parity suite, unchanged (function objects carry their own parametrisation; the `gpu` marker of
# that module does not apply here)
test_golden_small = P.test_golden_small
test_golden_checksums = P.test_golden_checksums
test_contracted_golden_small = P.test_contracted_golden_small
test_contracted_golden_checksums = P.test_contracted_golden_checksums
test_contracted_matrix = P.test_contracted_matrix
test_contracted_fused_depths = P.test_contracted_fused_depths
test_verification_std_matrix = P.test_verification_std_matrix
test_verification_idiam_matrix = P.test_verification_idiam_matrix
test_fused_depths = P.test_fused_depths
test_tiles_and_z_chunks = P.test_tiles_and_z_chunks
test_schedule_variant_tiles = P.test_schedule_variant_tiles
test_exact_tiles_match_oracle = P.test_exact_tiles_match_oracle
test_halo_copy_matches_global_oracle = P.test_halo_copy_matches_global_oracle
test_halo_copy_on_thin_slabs_matches_global_oracle = P.test_halo_copy_on_thin_slabs_matches_global_oracle
test_z_wavefront_through_l2_matches_oracle = P.test_z_wavefront_through_l2_matches_oracle
test_trapezoid_skip_variable_coefficients = P.test_trapezoid_skip_variable_coefficients
test_marching_kernel_tiles = P.test_marching_kernel_tiles
test_radius4_tile_seams = P.test_radius4_tile_seams
test_box_kernel_tile_seams = P.test_box_kernel_tile_seams
test_naive_variant_and_edge_sizes = P.test_naive_variant_and_edge_sizes
test_step_box_is_the_operator_contract = P.test_step_box_is_the_operator_contract
test_repeated_runs_keep_evolving_like_the_reference = P.test_repeated_runs_keep_evolving_like_the_reference
test_frame_mismatch_is_reported = P.test_frame_mismatch_is_reported
test_frame_mismatch_after_field_only_uploads_is_reported = P.test_frame_mismatch_after_field_only_uploads_is_reported
test_scan_counts_nan_and_zero = P.test_scan_counts_nan_and_zero
test_z_slabs_match_global_oracle = P.test_z_slabs_match_global_oracle
test_xy_topologies_match_global_oracle = P.test_xy_topologies_match_global_oracle
test_autotune_then_results_are_unchanged = P.test_autotune_then_results_are_unchanged
test_overlap_with_uneven_slabs_takes_one_schedule_on_every_rank = P.test_overlap_with_uneven_slabs_takes_one_schedule_on_every_rank
test_pipelined_transfers_keep_jobs_apart = P.test_pipelined_transfers_keep_jobs_apart
test_halo_push_matches_global_oracle = P.test_halo_push_matches_global_oracle
test_solar_matches_the_reference = P.test_solar_matches_the_reference
test_solar_schedules_chunks_and_boxes_match_oracle = P.test_solar_schedules_chunks_and_boxes_match_oracle
test_solar_limits_are_reported = P.test_solar_limits_are_reported
test_cli_verify = P.test_cli_verify
test_cli_verify_contracted = P.test_cli_verify_contracted
test_cli_autotune_prints_reference_prefix = P.test_cli_autotune_prints_reference_prefix


# ------------------------------------------------------------------------------------------------
# --npx / --npy / --npz topologies (src/mpi_utils.c:63-170): every rank a host thread on the emulator, the
# assembled sub-domains against the serial oracle on the global domain, as the reference itself
# verifies decomposed runs (src/verification.c:955-1040)
# ------------------------------------------------------------------------------------------------
def _topology_run(kernel, gst, dt, dims, fn):
    import threading
    nranks = dims[0] * dims[1] * dims[2]
    uid = G.GpuStepper.comm_unique_id()
    out, errs = [None] * nranks, []

    def work(rank):
        try:
            pb = G.make_problem(kernel, gst, dt, rank=rank, nranks=nranks, topology=dims)
            s = G.GpuStepper(kernel, pb.stencil, pb.shape, dt, device=rank % 4, rank=rank, nranks=nranks)
            s.set_topology(pb.dims, pb.coords)
            s.comm_init(uid)
            s.upload(pb)
            fn(s)
            s.download(pb.U1, pb.U2)
            s.close()
            out[rank] = pb
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]
    return out


def _assert_subdomains(slabs, ob, whole_halo=True):
    r = ob.r
    for pb in slabs:
        x0, y0, z0 = pb.gb
        nx, ny, nz = pb.stencil
        for mine, ref in ((pb.U1, ob.U1), (pb.U2, ob.U2)):
            assert np.array_equal(mine[r:r + nz, r:r + ny, r:r + nx],
                                  ref[z0 + r:z0 + r + nz, y0 + r:y0 + r + ny, x0 + r:x0 + r + nx])
        if whole_halo:   # halos (faces, edges and corners) of both arrays are current after the run
            for mine, ref in ((pb.U1, ob.U1), (pb.U2, ob.U2)):
                assert np.array_equal(mine[:, :, :nx + 2 * r], ref[z0:z0 + nz + 2 * r, y0:y0 + ny + 2 * r, x0:x0 + nx + 2 * r])


@pytest.mark.parametrize("dims", [(2, 1, 1), (1, 2, 1), (2, 2, 1), (2, 1, 2), (1, 3, 2), (2, 2, 2), (3, 1, 1)])
@pytest.mark.parametrize("kernel", [0, 1, 4, 5, 7])
def test_xyz_topologies_match_global_oracle(oracle, kernel, dims):
    gst, nsteps = (37, 29, 23), 6
    for dt in (np.float64, np.float32):
        for ts in (0, 1):
            slabs = _topology_run(kernel, gst, dt, dims, lambda s: s.run_single(nsteps, overlap=bool(ts)))
            ob = oracle.make_problem(kernel, gst, dt)
            oracle.run_steps(ob, nsteps)
            _assert_subdomains(slabs, ob)


def test_fused_stepper_falls_back_to_single_steps_on_xy_topologies(oracle):
    gst, nsteps = (40, 32, 24), 9
    infos = []

    def fn(s):
        s.run_fused(nsteps, 4)
        infos.append(s.launch_info())

    slabs = _topology_run(1, gst, np.float64, (2, 2, 1), fn)
    assert all(i["tfuse"] == 1 and i["steps"] == nsteps for i in infos)
    ob = oracle.make_problem(1, gst, np.float64)
    oracle.run_steps(ob, nsteps)
    _assert_subdomains(slabs, ob)


def test_topology_argument_checks():
    pb = G.make_problem(1, (16, 16, 16), np.float64, rank=1, nranks=4, topology=(2, 2, 1))
    assert pb.coords == (0, 1, 0) and pb.stencil == (8, 8, 16) and pb.gb == (0, 8, 0)
    s = G.GpuStepper(1, pb.stencil, pb.shape, np.float64, device=0, rank=1, nranks=4)
    with pytest.raises(G.GirihError):
        s.set_topology((2, 2, 2), (0, 1, 0))      # product != nranks
    with pytest.raises(G.GirihError):
        s.set_topology((2, 2, 1), (1, 0, 0))      # coordinates of another rank
    s.set_topology((2, 2, 1), (0, 1, 0))
    s.close()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("top", [(2, 1, 1), (1, 2, 1), (2, 2, 2), (1, 1, 3), (3, 2, 1)])
@pytest.mark.parametrize("ts,kernel", [(0, 1), (1, 0), (0, 4), (1, 7)])
def test_cli_verify_topologies(ts, kernel, top, dt):
    """mwd_kernel --npx/--npy/--npz --verify 1: the reference's verdict line (src/verification.c:851-852) from
    one rank thread per (emulated) GPU"""
    rc, out, err = emu_cli(dt, ["--nx", 48, "--ny", 36, "--nz", 30, "--nt", 10, "--target-ts", ts, "--target-kernel", kernel,
                                "--verify", 1, "--verbose", 0, "--npx", top[0], "--npy", top[1], "--npz", top[2]])
    assert rc == 0, out + err
    assert "eMax:0.000e+00|eL1:0.000e+00-PASSED" in out
    assert "top:(%d,%d,%d)" % top in out


def test_cli_diamond_rejects_xy_topologies():
    rc, out, err = emu_cli(np.float64, ["--nx", 48, "--ny", 32, "--nz", 40, "--nt", 20, "--target-ts", 2, "--target-kernel", 1,
                                        "--t-dim", 3, "--verify", 1, "--npx", 2])
    assert rc == 1 and "ERROR: the Diamond stepper of this build decomposes the domain across the Z direction only" in err


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_cli_halo_push(dt):
    """mwd_kernel --npz 3 --gpu-push 1: rank threads map each other's arrays, the Diamond stepper pushes its halos"""
    rc, out, err = emu_cli(dt, ["--nx", 70, "--ny", 32, "--nz", 50, "--nt", 30, "--target-ts", 2, "--target-kernel", 1,
                                "--t-dim", 3, "--verify", 1, "--npz", 3, "--gpu-push", 1, "--verbose", 0])
    assert rc == 0, out + err
    assert "eMax:0.000e+00|eL1:0.000e+00-PASSED" in out


# ------------------------------------------------------------------------------------------------
# bench.py's parity_check cases at the rank count of the driver's largest run: 8 z-slabs as 8 host threads, halos
# moved by the halo-copy schedule.  With 8 ranks the slabs are 5-12 planes thin, i.e. the passes mix the copy schedule
# and the blocking exchange (run_passes, girih_cuda.cu) -- the situation in which the emulator's fuzzer found a race.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel,dt,gst,nsteps,tfuse", [(1, np.float64, (96, 64, 96), 18, 4), (1, np.float32, (70, 41, 41), 10, 3),
                                                        (0, np.float32, (64, 40, 64), 8, 1), (5, np.float64, (70, 41, 48), 11, 3)])
def test_bench_parity_cases_on_eight_ranks_with_halo_copy(kernel, dt, gst, nsteps, tfuse):
    from oracle import girih_oracle as O

    def fn(s):
        if tfuse == 1:
            s.run_single(nsteps, overlap=True)
        else:
            s.run_fused(nsteps, tfuse)

    slabs = P._peer_linked_run(kernel, gst, dt, 8, fn, option="halo_copy")
    ob = O.make_problem(kernel, gst, dt)
    O.run_steps(ob, nsteps)
    r = ob.r
    for pb in slabs:
        z0, lnz = pb.gb[2], pb.stencil[2]
        assert np.array_equal(pb.U1[r:r + lnz], ob.U1[z0 + r:z0 + r + lnz])
        assert np.array_equal(pb.U2[r:r + lnz], ob.U2[z0 + r:z0 + r + lnz])
