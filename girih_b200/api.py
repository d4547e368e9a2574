"""ctypes mirror of include/girih_cuda.h and of the C host's initialisation.

Names follow the reference's domain: a *problem* is the set of host arrays of one z-slab
(U1, U2, U3 = roc2, coef) exactly as GIRIH's arrays_allocate / init_coeff / domain_data_fill produce
them (src/utils.c:153-218, 436-496, 605-697); a *stepper* owns the device copy of one slab and runs
the time steppers of the reference's TSList[] (src/wrappers.h:29-37).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from . import lib as _lib


class GirihError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        msg = _lib.cuda().girih_gpu_strerror(status).decode()
        super().__init__(f"{what}: {msg}" + (f" ({detail})" if detail else ""))


@dataclass(frozen=True)
class KernelDesc:
    name: str
    r: int
    time_order: int
    nd: int
    shape: int
    coeff: int
    n_coef_arrays: int
    n_coef_scalars: int
    words_per_lup: int
    max_tfuse: int
    gpu_supported: bool


def kernel_info(k: int) -> KernelDesc:
    d = _lib.KernelDescC()
    rc = _lib.cuda().girih_kernel_info(k, C.byref(d))
    if rc:
        raise GirihError(rc, "girih_kernel_info")
    return KernelDesc(d.name.decode(), d.r, d.time_order, d.nd, d.shape, d.coeff, d.n_coef_arrays,
                      d.n_coef_scalars, d.words_per_lup, d.max_tfuse, bool(d.gpu_supported))


def gpu_count() -> int:
    n = C.c_int(0)
    _lib.cuda().girih_gpu_count(C.byref(n))
    return n.value


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def diamond_nt(nt: int, t_dim: int) -> int:
    """nt after the Diamond stepper's rounding (src/kernels/diamond_utils.c:1042-1056)."""
    return _lib.host(8).girih_host_diamond_nt(nt, t_dim)


@dataclass
class HostProblem:
    kernel: int
    dtype: np.dtype
    gstencil: tuple      # global interior
    stencil: tuple       # this slab's interior
    shape: tuple         # this slab's host array shape (nnx, nny, nnz)
    gb: tuple            # global begin of this slab
    rank: int
    nranks: int
    r: int
    dims: tuple          # process topology (npx, npy, npz); z-slabs: (1, 1, nranks)
    coords: tuple        # this rank's position in it
    U1: np.ndarray       # [nnz, nny, nnx]; solar slot: [12, nnz, nny, nnx, 2]
    U2: np.ndarray | None  # None for the solar slot
    U3: np.ndarray | None
    coef: np.ndarray

    def interior(self, a=None):
        a = self.U1 if a is None else a
        r = self.r
        nx, ny, nz = self.stencil
        return a[r:r + nz, r:r + ny, r:r + nx]


def make_problem(kernel, gstencil, dtype=np.float64, rank=0, nranks=1, alignment=8, padding=True,
                 pinned=False, topology=None) -> HostProblem:
    """Host arrays of sub-domain `rank`: the C host's init() + init_coeff() + domain_data_fill().
    topology = (npx, npy, npz) as --npx/--npy/--npz; default: z-slabs, (1, 1, nranks)."""
    dtype = np.dtype(dtype)
    es = dtype.itemsize
    h = _lib.host(es)
    info = kernel_info(kernel)
    dims = (1, 1, nranks) if topology is None else tuple(int(d) for d in topology)
    if dims[0] * dims[1] * dims[2] != nranks:
        raise ValueError("topology does not match nranks")
    solar = kernel == 6   # 12 complex fields in ONE array [f][z][y][x][re, im], no U2, no x padding (src/utils.c:168-172, 359-361)
    if solar:
        padding = False
    ls, ds, gb, co = (C.c_int * 3)(), (C.c_int * 3)(), (C.c_int * 3)(), (C.c_int * 3)()
    h.girih_host_shapes_topo(kernel, _i3(gstencil), rank, _i3(dims), alignment, int(padding), ls, ds, gb, co)
    shape = tuple(ds)
    zyx = (shape[2], shape[1], shape[0])
    ncoef = int(h.girih_host_coef_size_topo(kernel, _i3(gstencil), rank, _i3(dims), alignment, int(padding)))

    def alloc(sh):
        if pinned:
            import torch
            t = torch.empty(sh, dtype=torch.float64 if es == 8 else torch.float32).pin_memory()
            return t.numpy()
        return np.empty(sh, dtype)

    if solar:
        U1, U2 = alloc((12,) + zyx + (2,)), None
    else:
        U1, U2 = alloc(zyx), alloc(zyx)
    U3 = alloc(zyx) if info.time_order == 2 else None
    coef = np.zeros(ncoef, dtype)
    rc = h.girih_host_fill_topo(kernel, _i3(gstencil), rank, _i3(dims), alignment, int(padding),
                                U1.ctypes.data, U2.ctypes.data if U2 is not None else None,
                                U3.ctypes.data if U3 is not None else None,
                                coef.ctypes.data)
    if rc:
        raise RuntimeError("girih_host_fill failed")
    return HostProblem(kernel, dtype, tuple(gstencil), tuple(ls), shape, tuple(gb), rank, nranks, info.r,
                       dims, tuple(co), U1, U2, U3, coef)


class GpuStepper:
    """Device state of one z-slab + the time steppers (girih_gpu_ctx)."""

    # the native library behind this class: libgirih_cuda.so (raises when it has not been built)
    _load = staticmethod(_lib.cuda)

    def __init__(self, kernel, stencil, shape, dtype=np.float64, device=0, rank=0, nranks=1):
        self._lib = self._load()
        self._ctx = C.c_void_p()
        self.kernel, self.dtype = kernel, np.dtype(dtype)
        self.stencil, self.shape = tuple(stencil), tuple(shape)
        self.rank, self.nranks = rank, nranks
        rc = self._lib.girih_gpu_create(C.byref(self._ctx), device, kernel, self.dtype.itemsize,
                                        _i3(stencil), _i3(shape), rank, nranks)
        if rc:
            self._ctx = C.c_void_p()
            raise GirihError(rc, "girih_gpu_create")

    @classmethod
    def for_problem(cls, pb: HostProblem, device=0):
        s = cls(pb.kernel, pb.stencil, pb.shape, pb.dtype, device, pb.rank, pb.nranks)
        s.upload(pb)
        return s

    def _check(self, rc, what):
        if rc:
            raise GirihError(rc, what, self._lib.girih_gpu_last_error(self._ctx).decode())

    def close(self):
        if self._ctx:
            self._lib.girih_gpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- communicator -------------------------------------------------------------------------
    @classmethod
    def comm_unique_id(cls) -> bytes:
        buf = C.create_string_buffer(128)
        rc = cls._load().girih_gpu_comm_unique_id(buf, 128)
        if rc:
            raise GirihError(rc, "girih_gpu_comm_unique_id")
        return buf.raw

    def set_topology(self, dims, coords):
        """(npx, npy, npz) and this rank's position; before comm_init.  Default: z-slabs."""
        self._check(self._lib.girih_gpu_set_topology(self._ctx, _i3(dims), _i3(coords)), "girih_gpu_set_topology")

    # -- halo push over peer memory (include/girih_cuda.h) ---------------------------------------
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(256)
        self._check(self._lib.girih_gpu_peer_export(self._ctx, buf, 256), "girih_gpu_peer_export")
        return buf.raw

    def peer_attach(self, which: int, blob: bytes):
        self._check(self._lib.girih_gpu_peer_attach(self._ctx, which, blob, len(blob)), "girih_gpu_peer_attach")

    def peer_detach(self):
        self._check(self._lib.girih_gpu_peer_detach(self._ctx), "girih_gpu_peer_detach")

    def comm_init(self, uid: bytes):
        self._check(self._lib.girih_gpu_comm_init(self._ctx, uid, len(uid)), "girih_gpu_comm_init")

    # -- transfers ----------------------------------------------------------------------------
    def upload(self, pb: HostProblem):
        for a in (pb.U1, pb.U2, pb.U3, pb.coef):
            assert a is None or (a.dtype == self.dtype and a.flags.c_contiguous)
        self._check(self._lib.girih_gpu_upload(
            self._ctx, pb.U1.ctypes.data, pb.U2.ctypes.data if pb.U2 is not None else None,
            pb.U3.ctypes.data if pb.U3 is not None else None, pb.coef.ctypes.data), "girih_gpu_upload")

    def upload_fields(self, U1, U2):
        self._check(self._lib.girih_gpu_upload_fields(
            self._ctx, U1.ctypes.data if U1 is not None else None,
            U2.ctypes.data if U2 is not None else None), "girih_gpu_upload_fields")

    def download(self, U1=None, U2=None):
        self._check(self._lib.girih_gpu_download(
            self._ctx, U1.ctypes.data if U1 is not None else None,
            U2.ctypes.data if U2 is not None else None), "girih_gpu_download")

    # pipelined transfers for a stream of independent jobs (see include/girih_cuda.h)
    def prefetch_fields(self, U1, U2):
        self._check(self._lib.girih_gpu_prefetch_fields(
            self._ctx, U1.ctypes.data if U1 is not None else None,
            U2.ctypes.data if U2 is not None else None), "girih_gpu_prefetch_fields")

    def commit_fields(self):
        self._check(self._lib.girih_gpu_commit_fields(self._ctx), "girih_gpu_commit_fields")

    def download_async(self, U1=None, U2=None):
        self._check(self._lib.girih_gpu_download_async(
            self._ctx, U1.ctypes.data if U1 is not None else None,
            U2.ctypes.data if U2 is not None else None), "girih_gpu_download_async")

    def sync_transfers(self):
        self._check(self._lib.girih_gpu_sync_transfers(self._ctx), "girih_gpu_sync_transfers")

    # -- steppers -----------------------------------------------------------------------------
    def run_single(self, nsteps, overlap=False):
        self._check(self._lib.girih_gpu_run_single(self._ctx, nsteps, int(overlap)), "girih_gpu_run_single")

    def run_fused(self, nsteps, tfuse=0):
        self._check(self._lib.girih_gpu_run_fused(self._ctx, nsteps, tfuse), "girih_gpu_run_fused")

    def run_ts(self, ts, nt, t_dim=1, tfuse=0):
        """The reference CLI's semantics for --target-ts / --nt (returns nt after rounding)."""
        if ts in (0, 1):
            self.run_single((nt + 1) // 2 * 2, overlap=(ts == 1))
            return nt
        nt = diamond_nt(nt, t_dim)
        self.run_fused(nt - 1, tfuse)
        return nt

    def step_box(self, dst, box):
        self._check(self._lib.girih_gpu_step_box(self._ctx, dst, *[int(b) for b in box]), "girih_gpu_step_box")

    def set_option(self, key, value):
        self._check(self._lib.girih_gpu_set_option(self._ctx, key.encode(), int(value)), "girih_gpu_set_option")

    def autotune(self, fused=True, verbose=False):
        """on-device search over fusion depth and tiles; returns (tfuse, tile, MLUP/s).  The fields evolve
        while it measures: upload again before a run whose result matters."""
        t, tile, perf = C.c_int(), C.c_int(), C.c_double()
        self._check(self._lib.girih_gpu_autotune(self._ctx, int(fused), int(verbose), C.byref(t), C.byref(tile),
                                                 C.byref(perf)), "girih_gpu_autotune")
        return t.value, tile.value, perf.value

    def time_pass(self, tfuse, reps=10):
        """average device milliseconds of one fused pass (= one kernel launch)"""
        ms = C.c_double()
        self._check(self._lib.girih_gpu_time_pass(self._ctx, tfuse, reps, C.byref(ms)), "girih_gpu_time_pass")
        return ms.value

    # -- accounting ---------------------------------------------------------------------------
    def elapsed_ms(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._lib.girih_gpu_last_elapsed_ms(self._ctx, C.byref(a), C.byref(b), C.byref(c))
        return {"compute": a.value, "comm": b.value, "total": c.value}

    def launch_info(self):
        v = [C.c_int() for _ in range(4)]
        self._lib.girih_gpu_last_launch_info(self._ctx, *[C.byref(x) for x in v])
        return {"kernels": v[0].value, "passes": v[1].value, "steps": v[2].value, "tfuse": v[3].value}

    def stat(self, key: str) -> int:
        v = C.c_longlong()
        self._check(self._lib.girih_gpu_get_stat(self._ctx, key.encode(), C.byref(v)), "girih_gpu_get_stat")
        return v.value

    def scan_u1(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._lib.girih_gpu_scan_u1(self._ctx, C.byref(a), C.byref(b)), "girih_gpu_scan_u1")
        return a.value, b.value


def plan_fused_passes(nsteps: int, tfuse: int) -> list:
    """steps per pass that girih_gpu_run_fused executes (host-only)"""
    n = C.c_int()
    buf = (C.c_int * (nsteps + 4))()
    rc = _lib.cuda().girih_plan_fused_passes(nsteps, tfuse, buf, nsteps + 4, C.byref(n))
    if rc:
        raise GirihError(rc, "girih_plan_fused_passes")
    return list(buf[:n.value])


def plan_fused_exchanges(nsteps: int, tfuse: int, r: int, halo_cap: int, group: int) -> list:
    """halo planes exchanged before each pass of plan_fused_passes(nsteps, tfuse); 0 = no exchange (host-only)"""
    n = C.c_int()
    buf = (C.c_int * (nsteps + 4))()
    rc = _lib.cuda().girih_plan_fused_exchanges(nsteps, tfuse, r, halo_cap, group, buf, nsteps + 4, C.byref(n))
    if rc:
        raise GirihError(rc, "girih_plan_fused_exchanges")
    return list(buf[:n.value])


def plan_halo_exchange(nz: int, depth: int, rank: int, nranks: int) -> dict:
    """first local plane of the send/recv blocks of one z exchange (host-only)"""
    v = [C.c_int() for _ in range(4)]
    rc = _lib.cuda().girih_plan_halo_exchange(nz, depth, rank, nranks, *[C.byref(x) for x in v])
    if rc:
        raise GirihError(rc, "girih_plan_halo_exchange")
    return dict(zip(("send_down", "recv_down", "send_up", "recv_up"), (x.value for x in v)))


def run_reference_cli(dtype, args, timeout=3600, env=None):
    """Run this repo's own mwd_kernel executable (build/ = fp32, build_dp/ = fp64); returns
    (returncode, stdout, stderr)."""
    exe = os.path.join(_lib.ROOT, "build_dp" if np.dtype(dtype).itemsize == 8 else "build", "mwd_kernel")
    out = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout,
                         env=dict(os.environ, **(env or {})))
    return out.returncode, out.stdout, out.stderr
