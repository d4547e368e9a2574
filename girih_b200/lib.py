"""Loads the native libraries.  Fails loudly when they have not been built (`make` at the repo root,
or __graft_entry__.build())."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CUDA_LIB = os.path.join(PKG, "libgirih_cuda.so")
HOST_LIBS = {4: os.path.join(PKG, "libgirih_host_sp.so"), 8: os.path.join(PKG, "libgirih_host_dp.so")}

# every symbol include/girih_cuda.h declares
ABI_SYMBOLS = [
    "girih_kernel_count", "girih_kernel_info", "girih_gpu_count", "girih_gpu_create", "girih_gpu_destroy",
    "girih_gpu_comm_unique_id", "girih_gpu_comm_init", "girih_gpu_set_topology", "girih_gpu_peer_export", "girih_gpu_peer_attach", "girih_gpu_peer_detach", "girih_gpu_upload", "girih_gpu_download",
    "girih_gpu_upload_fields", "girih_gpu_prefetch_fields", "girih_gpu_commit_fields", "girih_gpu_download_async",
    "girih_gpu_sync_transfers", "girih_gpu_run_single", "girih_gpu_run_fused", "girih_gpu_step_box",
    "girih_gpu_time_pass", "girih_gpu_last_elapsed_ms", "girih_gpu_last_launch_info", "girih_gpu_get_stat", "girih_gpu_scan_u1", "girih_gpu_set_option",
    "girih_gpu_autotune",
    "girih_gpu_strerror", "girih_gpu_last_error", "girih_plan_fused_passes", "girih_plan_halo_exchange", "girih_plan_fused_exchanges",
]


class KernelDescC(C.Structure):
    _fields_ = [("name", C.c_char_p), ("r", C.c_int), ("time_order", C.c_int), ("nd", C.c_int),
                ("shape", C.c_int), ("coeff", C.c_int), ("n_coef_arrays", C.c_int),
                ("n_coef_scalars", C.c_int), ("words_per_lup", C.c_int), ("max_tfuse", C.c_int),
                ("gpu_supported", C.c_int)]


_cuda = None
_host = {}


def cuda() -> C.CDLL:
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB):
            raise ImportError(f"{CUDA_LIB} is missing: build it with `make` (nvcc, sm_100a). "
                              "There is no Python or CPU fallback.")
        _cuda = declare(C.CDLL(CUDA_LIB, mode=C.RTLD_GLOBAL))
    return _cuda


def declare(lib: C.CDLL) -> C.CDLL:
    """ctypes prototypes of every entry point of include/girih_cuda.h on a loaded library."""
    P, I = C.c_void_p, C.c_int
    lib.girih_kernel_info.argtypes = [I, C.POINTER(KernelDescC)]
    lib.girih_gpu_count.argtypes = [C.POINTER(I)]
    lib.girih_gpu_create.argtypes = [C.POINTER(P), I, I, I, C.POINTER(I), C.POINTER(I), I, I]
    lib.girih_gpu_destroy.argtypes = [P]
    lib.girih_gpu_destroy.restype = None
    lib.girih_gpu_comm_unique_id.argtypes = [P, C.c_size_t]
    lib.girih_gpu_comm_init.argtypes = [P, P, C.c_size_t]
    lib.girih_gpu_set_topology.argtypes = [P, C.POINTER(I), C.POINTER(I)]
    lib.girih_gpu_peer_export.argtypes = [P, P, C.c_size_t]
    lib.girih_gpu_peer_attach.argtypes = [P, I, P, C.c_size_t]
    lib.girih_gpu_peer_detach.argtypes = [P]
    lib.girih_gpu_upload.argtypes = [P, P, P, P, P]
    lib.girih_gpu_download.argtypes = [P, P, P]
    lib.girih_gpu_upload_fields.argtypes = [P, P, P]
    lib.girih_gpu_prefetch_fields.argtypes = [P, P, P]
    lib.girih_gpu_commit_fields.argtypes = [P]
    lib.girih_gpu_download_async.argtypes = [P, P, P]
    lib.girih_gpu_sync_transfers.argtypes = [P]
    lib.girih_gpu_run_single.argtypes = [P, I, I]
    lib.girih_gpu_run_fused.argtypes = [P, I, I]
    lib.girih_gpu_step_box.argtypes = [P, I, I, I, I, I, I, I]
    lib.girih_gpu_time_pass.argtypes = [P, I, I, C.POINTER(C.c_double)]
    lib.girih_gpu_last_elapsed_ms.argtypes = [P] + [C.POINTER(C.c_double)] * 3
    lib.girih_gpu_last_launch_info.argtypes = [P] + [C.POINTER(I)] * 4
    lib.girih_gpu_get_stat.argtypes = [P, C.c_char_p, C.POINTER(C.c_longlong)]
    lib.girih_gpu_scan_u1.argtypes = [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.girih_gpu_set_option.argtypes = [P, C.c_char_p, I]
    lib.girih_gpu_autotune.argtypes = [P, I, I, C.POINTER(I), C.POINTER(I), C.POINTER(C.c_double)]
    lib.girih_plan_fused_passes.argtypes = [I, I, C.POINTER(I), I, C.POINTER(I)]
    lib.girih_plan_fused_exchanges.argtypes = [I, I, I, I, I, C.POINTER(I), I, C.POINTER(I)]
    lib.girih_plan_halo_exchange.argtypes = [I, I, I, I] + [C.POINTER(I)] * 4
    lib.girih_gpu_strerror.argtypes = [I]
    lib.girih_gpu_strerror.restype = C.c_char_p
    lib.girih_gpu_last_error.argtypes = [P]
    lib.girih_gpu_last_error.restype = C.c_char_p
    return lib


def host(elem_size: int) -> C.CDLL:
    if elem_size not in _host:
        path = HOST_LIBS[elem_size]
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with `make`.")
        cuda()   # dependency, resolved by rpath as well
        lib = C.CDLL(path)
        I, P = C.c_int, C.c_void_p
        I3 = C.POINTER(I)
        lib.girih_host_shapes.argtypes = [I, I3, I, I, I, I, I3, I3, I3]
        lib.girih_host_coef_size.argtypes = [I, I3, I, I, I, I]
        lib.girih_host_coef_size.restype = C.c_ulonglong
        lib.girih_host_fill.argtypes = [I, I3, I, I, I, I, P, P, P, P]
        lib.girih_host_diamond_nt.argtypes = [I, I]
        lib.girih_host_shapes_topo.argtypes = [I, I3, I, I3, I, I, I3, I3, I3, I3]
        lib.girih_host_coef_size_topo.argtypes = [I, I3, I, I3, I, I]
        lib.girih_host_coef_size_topo.restype = C.c_ulonglong
        lib.girih_host_fill_topo.argtypes = [I, I3, I, I3, I, I, P, P, P, P]
        _host[elem_size] = lib
    return _host[elem_size]
