"""girih_b200 -- B200 (sm_100a) implementation of GIRIH's star-stencil time stepper.

The product is native: hand-written CUDA kernels behind a C ABI (include/girih_cuda.h, built into
girih_b200/libgirih_cuda.so) and a C host (girih_b200/host/, the `mwd_kernel` executables plus
libgirih_host_{sp,dp}.so).  This Python package is a thin ctypes mirror of that interface for tests
and bench.py -- it contains no arithmetic and NO fallback: if the native libraries are missing the
import of girih_b200.lib raises, and without a CUDA device every stepper call raises GirihError.
"""
from .api import (GirihError, GpuStepper, HostProblem, KernelDesc, diamond_nt, gpu_count,  # noqa: F401
                  kernel_info, make_problem, plan_fused_exchanges, plan_fused_passes, plan_halo_exchange,
                  run_reference_cli)

__all__ = ["GirihError", "GpuStepper", "HostProblem", "KernelDesc", "diamond_nt", "gpu_count",
           "kernel_info", "make_problem", "plan_fused_exchanges", "plan_fused_passes", "plan_halo_exchange",
           "run_reference_cli"]
