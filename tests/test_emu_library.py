"""The complete C-ABI library (girih_b200/csrc/girih_cuda.cu: context, transfers, pass and exchange
schedules, steppers, options) compiled for the CPU SIMT emulator, driven through the same Python mirror
and the SAME test functions as the GPU parity suite (tests/test_gpu_parity.py).  Ranks of a z-slab run
are host threads; NCCL is an in-process mailbox (tests/cuda_emu/emu_nccl.cpp).

This is a checker for the host-side logic and the kernel logic together -- it never stands in for the
GPU run: the product loads libgirih_cuda.so only, and nothing under girih_b200/ knows the emulator
exists.  Stream/event ordering is not modelled (every operation completes in program order)."""
import ctypes as C

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


@pytest.fixture(autouse=True)
def _route_mirror_to_emulator(monkeypatch):
    monkeypatch.setattr(G.GpuStepper, "_load", staticmethod(emu_library))
    monkeypatch.setattr(G, "gpu_count", lambda: 4)
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
test_marching_kernel_tiles = P.test_marching_kernel_tiles
test_radius4_tile_seams = P.test_radius4_tile_seams
test_box_kernel_tile_seams = P.test_box_kernel_tile_seams
test_naive_variant_and_edge_sizes = P.test_naive_variant_and_edge_sizes
test_step_box_is_the_operator_contract = P.test_step_box_is_the_operator_contract
test_repeated_runs_keep_evolving_like_the_reference = P.test_repeated_runs_keep_evolving_like_the_reference
test_frame_mismatch_is_reported = P.test_frame_mismatch_is_reported
test_scan_counts_nan_and_zero = P.test_scan_counts_nan_and_zero
test_z_slabs_match_global_oracle = P.test_z_slabs_match_global_oracle
