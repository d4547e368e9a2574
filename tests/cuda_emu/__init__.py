"""CPU SIMT emulator of the library's CUDA kernels -- TEST INFRASTRUCTURE (see include/cuda_runtime.h).

`EmuStepper` mirrors the part of girih_b200.GpuStepper the parity tests use, but runs the kernel and
launcher sources (girih_b200/csrc/*.cu, *.cuh, compiled unchanged as C++) on host memory.  Nothing in the
product imports this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libgirih_emu.so")
_lib = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "-j", str(min(8, os.cpu_count() or 1))])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()   # make is a no-op when the library is newer than every kernel source
        _lib = C.CDLL(LIB_PATH)
        _lib.emu_create.restype = C.c_void_p
        _lib.emu_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int,
                                    C.c_int, C.c_int]
        _lib.emu_destroy.argtypes = [C.c_void_p]
        _lib.emu_upload.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int]
        _lib.emu_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.emu_pass.argtypes = [C.c_void_p] + [C.c_int] * 10
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class EmuStepper:
    """One slab on the emulator.  `desc` is the library's own operator-table row (girih_kernel_info)."""

    def __init__(self, kernel, stencil, shape, dtype, desc):
        self.kernel, self.dtype, self.desc = kernel, np.dtype(dtype), desc
        self.stencil = tuple(int(s) for s in stencil)
        i3 = lambda v: (C.c_int * 3)(*[int(x) for x in v])
        self._h = lib().emu_create(kernel, self.dtype.itemsize, i3(stencil), i3(shape), desc.r, desc.time_order,
                                   desc.n_coef_arrays, desc.max_tfuse)
        self.cur = 1          # array holding the newest level: U2 after upload (U1 is written first)
        self.tile = self.variant = self.contract = self.zchunk = 0
        self.launches = 0

    def close(self):
        if self._h:
            lib().emu_destroy(self._h)
            self._h = None

    def upload(self, pb):
        lib().emu_upload(self._h, _p(pb.U1), _p(pb.U2), _p(pb.U3), _p(pb.coef), int(pb.coef.size))
        self.cur = 1

    def download(self, U1, U2):
        lib().emu_download(self._h, _p(U1), _p(U2))

    def one_pass(self, T, zb=0, ze=None, zb1=0, ze1=0):
        ze = self.stencil[2] if ze is None else ze
        rc = lib().emu_pass(self._h, T, self.cur, zb, ze, zb1, ze1, self.tile, self.variant, self.contract,
                            self.zchunk)
        if rc != 0:
            raise RuntimeError(f"emulated launch failed with cudaError {rc}")
        self.launches += 1

    def run_passes(self, sizes):
        """Full-slab passes of the given depths, ping-ponging between the two arrays."""
        for T in sizes:
            self.one_pass(T)
            self.cur ^= 1
