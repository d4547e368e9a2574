"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/girih_cuda.h declares, validates arguments, and refuses to run without a GPU (no CPU
fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import girih_b200 as G
from girih_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "girih_cuda.h")).read()
    declared = set(re.findall(r"\b(girih_(?:gpu|kernel|plan)_\w+)\s*\(", hdr))
    assert declared == set(L.ABI_SYMBOLS)
    lib = L.cuda()
    for s in declared:
        assert hasattr(lib, s), s


def test_no_nccl_or_libcuda_link_dependency():
    out = subprocess.run(["ldd", L.CUDA_LIB], capture_output=True, text=True).stdout
    assert "libnccl" not in out and "libcuda.so" not in out and "libcudart" not in out


def test_kernel_table_matches_reference_list():
    # stencil_info_list[], src/kernels/stencils.c:260-271
    want = [("star", 4, 2, 3, 0), ("star", 1, 1, 2, 0), ("star", 1, 1, 4, 1), ("star", 1, 1, 6, 2),
            ("star", 4, 1, 15, 2), ("star", 1, 1, 9, 3), ("star", 1, 1, 40, 4), ("box", 1, 1, 2, 0)]
    assert L.cuda().girih_kernel_count() == 8
    for k, w in enumerate(want):
        d = G.kernel_info(k)
        assert (d.name, d.r, d.time_order, d.nd, d.coeff) == w
    assert all(G.kernel_info(k).gpu_supported for k in range(8))   # the solar slot (6) has kernels since round 2
    # algorithmic words per LUP; solar: 24 field reals in + out, 56 coefficient reals
    assert [G.kernel_info(k).words_per_lup for k in (0, 1, 2, 3, 4, 5, 6)] == [4, 2, 4, 6, 15, 9, 104]
    assert [G.kernel_info(k).max_tfuse for k in (0, 1, 2, 3, 4, 5, 7)] == [1, 4, 3, 3, 1, 3, 1]


def _create(kernel, es, st, ds, rank=0, nranks=1):
    ctx = C.c_void_p()
    i3 = lambda v: (C.c_int * 3)(*v)
    return L.cuda().girih_gpu_create(C.byref(ctx), 0, kernel, es, i3(st), i3(ds), rank, nranks)


def test_argument_validation_and_no_cpu_fallback():
    assert _create(9, 8, (8, 8, 8), (16, 10, 10)) == 1            # GIRIH_ERR_ARG
    assert _create(1, 2, (8, 8, 8), (16, 10, 10)) == 1
    assert _create(1, 8, (8, 8, 8), (16, 10, 11)) == 1            # nnz != nz + 2r
    assert _create(1, 8, (8, 8, 8), (16, 10, 10), 2, 2) == 1
    assert _create(6, 8, (8, 8, 8), (10, 10, 10), 0, 2) == 4      # solar on more than one rank: GIRIH_ERR_UNSUPPORTED
    assert _create(6, 8, (8, 8, 8), (16, 10, 10)) == 1            # solar arrays are not padded (src/utils.c:359-361)
    if G.gpu_count() == 0:
        assert _create(1, 8, (8, 8, 8), (16, 10, 10)) == 2        # GIRIH_ERR_NO_DEVICE
        with pytest.raises(G.GirihError):
            G.GpuStepper.for_problem(G.make_problem(1, (8, 8, 8)))


def test_strerror():
    assert b"no CPU fallback" in L.cuda().girih_gpu_strerror(2)
    assert L.cuda().girih_gpu_strerror(0) == b"success"


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("kernel", [0, 1, 2, 3, 4, 5, 7])
def test_host_init_matches_oracle(oracle, kernel, dt):
    """C host init()/init_coeff()/domain_data_fill() == the oracle's restatement of src/utils.c."""
    for st in ((20, 12, 14), (33, 9, 17)):
        pb, ob = G.make_problem(kernel, st, dt), oracle.make_problem(kernel, st, dt)
        assert pb.shape == ob.shape
        assert pb.U1.tobytes() == ob.U1.tobytes() and pb.U2.tobytes() == ob.U2.tobytes()
        assert pb.coef.tobytes() == ob.coef.tobytes()
        if pb.U3 is not None:
            assert pb.U3.tobytes() == ob.U3.tobytes()


def test_host_slab_decomposition_matches_oracle(oracle):
    for nranks in (2, 3, 8):
        for rank in range(nranks):
            pb = G.make_problem(0, (12, 10, 37), np.float32, rank=rank, nranks=nranks)
            lnz, gbz = oracle.decompose(37, nranks, rank)
            ob = oracle.make_problem(0, (12, 10, lnz), np.float32, gstencil=(12, 10, 37), gb=(0, 0, gbz),
                                     first=(1, 1, int(rank == 0)), last=(1, 1, int(rank == nranks - 1)))
            assert pb.stencil == (12, 10, lnz) and pb.gb == (0, 0, gbz)
            assert pb.U1.tobytes() == ob.U1.tobytes() and pb.U3.tobytes() == ob.U3.tobytes()


def test_diamond_nt_rounding(oracle):
    for nt in range(2, 140, 7):
        for td in (1, 3, 5, 7):
            assert G.diamond_nt(nt, td) == oracle.diamond_round_nt(nt, td)
    assert G.diamond_nt(100, 7) == 114      # SURVEY.md 8(a) a12 [probe]


def test_cli_list_and_help_match_reference_behaviour(oracle):
    rc, out, err = G.run_reference_cli(np.float64, ["--list"])
    assert rc == 0
    if oracle.have_ref():
        assert out == oracle.ref_cli(np.float64, ["--list"])
    rc, out, err = G.run_reference_cli(np.float32, ["--help"])
    assert rc == 0 and "--target-kernel" in out
    rc, out, err = G.run_reference_cli(np.float32, ["--bogus-flag"])
    assert rc == 0 and "Invalid arguments" in err          # src/utils.c:1308-1323
    rc, out, err = G.run_reference_cli(np.float32, ["--target-kernel", 6, "--target-ts", 2, "--verbose", 0])
    assert rc == 1 and "unsupported configuration" in out  # solar + default wavefront: src/kernels/stencils.h:40-47
    rc, out, err = G.run_reference_cli(np.float32, ["--target-ts", 2, "--target-kernel", 0, "--mwd-type", 2,
                                                    "--nx", 32, "--ny", 32, "--nz", 32, "--t-dim", 1])
    assert rc == 1 and "Relaxed synchronization" in err     # diamond_utils.c:859
    rc, out, err = G.run_reference_cli(np.float32, ["--target-ts", 2, "--target-kernel", 1, "--t-dim", 2,
                                                    "--nx", 32, "--ny", 32, "--nz", 32])
    assert rc == 1 and "even time unrolling" in err         # diamond_utils.c:1029
    rc, out, err = G.run_reference_cli(np.float32, ["--npx", 2, "--target-ts", 2, "--target-kernel", 1, "--t-dim", 1,
                                                    "--nx", 32, "--ny", 32, "--nz", 32])   # Diamond: z-slabs only
    assert rc == 1 and "Z direction only" in err
    if G.gpu_count() == 0:
        rc, out, err = G.run_reference_cli(np.float32, ["--nx", 16, "--ny", 16, "--nz", 16, "--verbose", 0])
        assert rc == 1 and "no CUDA device" in err
