"""ctypes binding of oracle/girih_oracle.c (the CPU restatement) plus helpers that drive the
real reference binaries in oracle/_ref/ when they exist.

TEST INFRASTRUCTURE -- not product code (see oracle/__init__.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libgirih_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_lib = None


def build(force: bool = False) -> None:
    """Compile the C restatement (and the reference, where /root/reference exists)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/src") and (force or not have_ref()):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "-j", "8"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_coef_size.restype = C.c_uint64
        _lib.oracle_coef_size.argtypes = [C.c_int, C.c_uint64]
        _lib.oracle_diamond_round_nt.restype = C.c_int
    return _lib


def have_ref() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f))
               for f in ("mwd_kernel_sp", "mwd_kernel_dp", "ref_dump_sp", "ref_dump_dp"))


@dataclass(frozen=True)
class KernelInfo:
    r: int
    time_order: int
    nd: int
    coeff: int      # 0 const, 1 var, 2 axsym, 3 nosym, 4 solar
    is_box: bool


def kernel_info(k: int) -> KernelInfo:
    out = (C.c_int * 5)()
    if lib().oracle_kernel_info(k, out) != 0:
        raise ValueError(f"bad kernel {k}")
    return KernelInfo(out[0], out[1], out[2], out[3], bool(out[4]))


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def _sfx(dtype) -> str:
    return "dp" if np.dtype(dtype) == np.float64 else "sp"


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def domain_shape(lstencil, r, alignment=8, padding=True):
    out = (C.c_int * 3)()
    lib().oracle_domain_shape(_i3(lstencil), r, alignment, int(padding), out)
    return tuple(out)


def decompose(n, nparts, coord):
    ln, gb = C.c_int(), C.c_int()
    lib().oracle_decompose(n, nparts, coord, C.byref(ln), C.byref(gb))
    return ln.value, gb.value


def diamond_round_nt(nt, t_dim):
    return lib().oracle_diamond_round_nt(nt, t_dim)


@dataclass
class Problem:
    """Host arrays of one (sub)domain in the reference's layout ([z][y][x], x fastest)."""
    kernel: int
    dtype: np.dtype
    stencil: tuple          # local interior (nx, ny, nz)
    shape: tuple            # local domain (nnx, nny, nnz)
    r: int
    U1: np.ndarray
    U2: np.ndarray
    U3: np.ndarray | None
    coef: np.ndarray

    def interior(self, a=None):
        a = self.U1 if a is None else a
        r = self.r
        nx, ny, nz = self.stencil
        return a[r:r + nz, r:r + ny, r:r + nx]


def make_problem(kernel, stencil, dtype=np.float64, alignment=8, padding=True,
                 gstencil=None, gb=(0, 0, 0), first=(1, 1, 1), last=(1, 1, 1)) -> Problem:
    """allocate + init_coeff + domain_data_fill (src/performance.c:50-52) for one subdomain."""
    info = kernel_info(kernel)
    dtype = np.dtype(dtype)
    gstencil = tuple(stencil) if gstencil is None else tuple(gstencil)
    if kernel == 6:
        # solar: no array padding (src/utils.c:359-361), ONE array of 12 complex fields [f][z][y][x][re,im], U2 = 0
        # (src/utils.c:168-172), 28 complex coefficient arrays (:199-201)
        shape = domain_shape(stencil, info.r, alignment, False)
        n = shape[0] * shape[1] * shape[2]
        U1 = np.empty((12, shape[2], shape[1], shape[0], 2), dtype)
        coef = np.zeros(n * 56, dtype)
        getattr(lib(), "oracle_solar_init_coeff_" + _sfx(dtype))(C.c_uint64(n), _p(coef))
        getattr(lib(), "oracle_solar_fill_" + _sfx(dtype))(_i3(shape), _i3(gstencil), _i3(gb), _p(U1))
        return Problem(kernel, dtype, tuple(stencil), shape, info.r, U1, None, None, coef)
    shape = domain_shape(stencil, info.r, alignment, padding)
    n = shape[0] * shape[1] * shape[2]
    zyx = (shape[2], shape[1], shape[0])
    U1 = np.empty(zyx, dtype)
    U2 = np.empty(zyx, dtype)
    U3 = np.empty(zyx, dtype) if info.time_order == 2 else None
    coef = np.zeros(int(lib().oracle_coef_size(kernel, n)), dtype)
    getattr(lib(), "oracle_init_coeff_" + _sfx(dtype))(kernel, C.c_uint64(n), _p(coef))
    getattr(lib(), "oracle_fill_" + _sfx(dtype))(
        _i3(shape), _i3(stencil), _i3(gstencil), _i3(gb), _i3(first), _i3(last), info.r,
        _p(U1), _p(U2), _p(U3))
    return Problem(kernel, dtype, tuple(stencil), shape, info.r, U1, U2, U3, coef)


def _fsfx(dtype, contract) -> str:
    """contract=True selects the copy of the oracle compiled with gcc's FMA contraction (_dpf/_spf)."""
    return _sfx(dtype) + ("f" if contract else "")


def step(kernel, shape, box, coef, u, v, roc2, contract=False):
    """u <- one stencil application of v over box=(xb,yb,zb,xe,ye,ze)."""
    getattr(lib(), "oracle_step_" + _fsfx(u.dtype, contract))(
        kernel, _i3(shape), *[int(b) for b in box], _p(coef), _p(u), _p(v), _p(roc2))


def run_naive(pb: Problem, nt: int, contract=False) -> None:
    """The reference's ts 0 loop (nb_naive_ts.c:187-203): nt rounded up to even steps."""
    if pb.kernel == 6:
        return run_steps(pb, nt + (nt & 1), contract)
    getattr(lib(), "oracle_run_naive_" + _fsfx(pb.dtype, contract))(
        pb.kernel, _i3(pb.shape), pb.stencil[0], nt, _p(pb.coef), _p(pb.U1), _p(pb.U2), _p(pb.U3))


def run_steps(pb: Problem, nsteps: int, contract=False) -> None:
    """Exactly nsteps steps, odd steps writing U1 (what ts 2 leaves: nsteps = nt-1)."""
    if pb.kernel == 6:   # solar: every step updates the one array in place
        getattr(lib(), "oracle_solar_run_" + _fsfx(pb.dtype, contract))(
            _i3(pb.shape), pb.stencil[0], nsteps, _p(pb.coef), _p(pb.U1))
        return
    getattr(lib(), "oracle_run_steps_" + _fsfx(pb.dtype, contract))(
        pb.kernel, _i3(pb.shape), pb.stencil[0], nsteps, _p(pb.coef), _p(pb.U1), _p(pb.U2), _p(pb.U3))


def compare(ref_full, target_interior, stencil, r):
    """(passed, max_err, l1_err, max_ref) per src/verification.c:823-860."""
    me, l1, mr = C.c_double(), C.c_double(), C.c_double()
    target_interior = np.ascontiguousarray(target_interior)
    rc = getattr(lib(), "oracle_compare_" + _sfx(ref_full.dtype))(
        _p(ref_full), _p(target_interior), *[int(s) for s in stencil], r,
        C.byref(me), C.byref(l1), C.byref(mr))
    return rc == 0, me.value, l1.value, mr.value


# ----------------------------------------------------------------------------------------------
# the real reference (oracle/_ref), when present
# ----------------------------------------------------------------------------------------------
def ref_dump(kernel, stencil, nt, dtype=np.float64, ts=0, extra=(), threads=2, fast=False, timeout=600):
    """Run the unmodified reference stepper and return (U1 full domain [z,y,x], r, nt_effective).
    fast=True uses the build with FMA contraction (-O3 -mfma -ffp-contract=fast)."""
    exe = os.path.join(REF_DIR, "ref_dump_" + _sfx(dtype) + ("_fast" if fast else ""))
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        path = f.name
    try:
        env = dict(os.environ, GIRIH_REF_DUMP=path, OMP_NUM_THREADS=str(threads))
        cmd = [exe, "--nx", str(stencil[0]), "--ny", str(stencil[1]), "--nz", str(stencil[2]),
               "--nt", str(nt), "--target-kernel", str(kernel), "--target-ts", str(ts),
               "--verbose", "0"]
        if ts != 2:
            cmd += ["--thread-group-size", str(threads)]
        cmd += [str(e) for e in extra]
        subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL, timeout=timeout)
        raw = open(path, "rb").read()
    finally:
        os.unlink(path)
    hdr = np.frombuffer(raw[:32], np.int32)
    assert hdr[0] == 0x47495249 and hdr[1] == np.dtype(dtype).itemsize
    nnx, nny, nnz, r, nt_eff = (int(x) for x in hdr[2:7])
    if kernel == 6:   # solar: 12 complex fields in one array
        U1 = np.frombuffer(raw[32:], dtype).reshape(12, nnz, nny, nnx, 2).copy()
    else:
        U1 = np.frombuffer(raw[32:], dtype).reshape(nnz, nny, nnx).copy()
    return U1, r, nt_eff


def ref_cli(dtype, args, threads=None, fast=False, timeout=3600):
    """Run the reference's own mwd_kernel CLI; returns stdout."""
    name = "mwd_kernel_" + _sfx(dtype) + ("_fast" if fast else "")
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    out = subprocess.run([os.path.join(REF_DIR, name)] + [str(a) for a in args], env=env,
                         check=True, capture_output=True, text=True, timeout=timeout)
    return out.stdout + out.stderr
