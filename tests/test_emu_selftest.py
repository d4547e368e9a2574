"""Known-answer tests of the CPU SIMT emulator itself (tests/cuda_emu): if the emulator did not
implement warp shuffles, votes, barriers and cp.async groups with CUDA's semantics -- or could not expose
a missing barrier -- the kernel tests of test_emu_kernels.py would prove nothing."""
import ctypes as C

import numpy as np
import pytest

import cuda_emu as E


def selftest(which, out, src=None, rounds=0):
    f = E.lib().emu_selftest
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    return f(which, out.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p) if src is not None else None, rounds)


@pytest.mark.parametrize("order", ["0", "1", "2"])
def test_shuffles_and_votes(monkeypatch, order):
    monkeypatch.setenv("CUDA_EMU_SCHED", order)
    out = np.zeros((128, 8), np.int32)
    assert selftest(0, out) == 0
    for b in range(2):
        for t in range(64):
            lane, warp = t % 32, t // 32
            v = 100 * warp + lane
            o = out[b * 64 + t]
            assert o[0] == (v - 1 if lane >= 1 else v)            # shfl_up keeps the own value at the low end
            assert o[1] == (v + 3 if lane + 3 <= 31 else v)       # shfl_down at the high end
            assert o[2] == 2 * (v + 1 if lane < 31 else v) + 1    # 64-bit payload
            assert np.uint32(o[3]) == np.uint32(sum(1 << l for l in range(32) if l % 3 == 0))
            assert o[4] == (1 if warp == 1 else 0) and o[5] == 1
            assert o[6] == 100 * warp + 5 and o[7] == 2


def _exchange_expected(n, rounds):
    return np.array([sum(r * 1000 + (t + 37) % n for r in range(rounds)) for t in range(n)] * 3, np.int32)


@pytest.mark.parametrize("order", ["0", "1", "2"])
def test_barrier_makes_exchange_order_independent(monkeypatch, order):
    monkeypatch.setenv("CUDA_EMU_SCHED", order)
    out = np.zeros(3 * 96, np.int32)
    assert selftest(1, out, rounds=6) == 0
    assert (out == _exchange_expected(96, 6)).all()


def test_missing_barrier_is_exposed(monkeypatch):
    """the same kernel without its __syncthreads: at least one scheduling order must give wrong results"""
    wrong = 0
    for order in ("0", "1", "2"):
        monkeypatch.setenv("CUDA_EMU_SCHED", order)
        out = np.zeros(3 * 96, np.int32)
        assert selftest(2, out, rounds=6) == 0
        wrong += int((out != _exchange_expected(96, 6)).any())
    assert wrong >= 1


def test_cp_async_groups():
    src = np.arange(256, dtype=np.float64) + 0.25
    out = np.zeros(64, np.int32)
    assert selftest(3, out, src) == 0
    # stale before any wait, first group after wait<1>, second still pending, second after wait<0>,
    # neighbour's data visible after the barrier
    assert (out == 31).all()


def test_early_exit_and_bad_configuration():
    out = np.zeros(64, np.int32)
    assert selftest(4, out) == 0
    assert out[:32].all() and not out[32:].any()
    assert selftest(5, out) != 0      # 2048 threads per CTA: cudaErrorInvalidValue, nothing runs
