"""The library's CUDA kernel and launcher sources, compiled unchanged for the CPU SIMT emulator
(tests/cuda_emu: every CUDA thread a fiber; __syncthreads, warp shuffles/votes and cp.async with their
CUDA semantics), against the oracle and the committed outputs of the unmodified reference.

What this tier proves without a GPU: tiling, overlap/halo geometry, masks for the Dirichlet frame and
partial tiles, register-plane rotation, shared-memory exchange and barrier placement, z chunking and
two-range launches of every kernel are right -- bit for bit -- for shapes that exercise partial tiles
in x, y and z.  What it cannot prove: anything about speed, and hardware-only hazards (memory-model
races between fibers that the three scheduling orders below do not expose).  The `-m gpu` tests remain
the parity gate for the CUDA build itself; the emulator is test infrastructure, never a fallback."""
import os
import re

import numpy as np
import pytest

import girih_b200 as G
import cuda_emu as E

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = np.load(os.path.join(HERE, "golden", "small.npz"))
SMALL_FMA = np.load(os.path.join(HERE, "golden", "small_fma.npz"))
KEY = re.compile(r"k(\d)_(\d+)x(\d+)x(\d+)_nt(\d+)_ts(\d)_td(\d)_(sp|dp)$")
F32, F64 = np.dtype(np.float32), np.dtype(np.float64)


@pytest.fixture(scope="module", autouse=True)
def _emu_built():
    E.lib()


@pytest.fixture(autouse=True)
def _default_order(monkeypatch):
    monkeypatch.delenv("CUDA_EMU_SCHED", raising=False)


def emu_run(O, kernel, st, dt, sizes, contract=False, **opt):
    """Run full-slab passes of the given depths on the emulator; returns the host problem."""
    pb = O.make_problem(kernel, st, dt)
    s = E.EmuStepper(kernel, st, pb.shape, dt, G.kernel_info(kernel))
    s.contract = int(contract)
    for k, v in opt.items():
        setattr(s, k, v)
    s.upload(pb)
    s.run_passes(sizes)
    s.download(pb.U1, pb.U2)
    s.close()
    return pb


def oracle_steps(O, kernel, st, dt, nsteps, contract=False):
    ob = O.make_problem(kernel, st, dt)
    O.run_steps(ob, nsteps, contract=contract)
    return ob


def same(pb, ob):
    assert pb.U1.tobytes() == ob.U1.tobytes(), "U1 differs from the oracle (bit-exact expected)"
    assert pb.U2.tobytes() == ob.U2.tobytes(), "U2 differs from the oracle (bit-exact expected)"


# shapes: one partial tile in every direction / nx not a multiple of the vector width / several tiles
SHAPES_R1 = [(70, 30, 20), (37, 9, 5), (131, 53, 11)]
SHAPES_R4 = [(70, 30, 20), (37, 9, 9), (131, 21, 12)]


# ------------------------------------------------------------------------------------------------
# single-step kernels: k_naive, k_r1_march, k_r4_async / k_r4 / k_r4_strip, k_box_march
# ------------------------------------------------------------------------------------------------
SINGLE_TILES = {0: (0, 8, 16, 116), 4: (0, 8, 16), 7: (0, 4), 1: (0, 108, 208, 404, 408),
                2: (0, 108, 404), 3: (0, 208), 5: (0, 408)}


@pytest.mark.parametrize("dt", [F32, F64], ids=["sp", "dp"])
@pytest.mark.parametrize("kernel", [0, 1, 2, 3, 4, 5, 7])
def test_single_step_kernels(oracle, kernel, dt):
    shapes = SHAPES_R4 if G.kernel_info(kernel).r == 4 else SHAPES_R1
    for st in shapes:
        ob = oracle_steps(oracle, kernel, st, dt, 3)
        same(emu_run(oracle, kernel, st, dt, [1, 1, 1], variant=1), ob)          # one thread per site
        for tile in SINGLE_TILES[kernel]:
            for zchunk in (0, 4):
                same(emu_run(oracle, kernel, st, dt, [1, 1, 1], tile=tile, zchunk=zchunk), ob)


@pytest.mark.parametrize("dt", [F32, F64], ids=["sp", "dp"])
@pytest.mark.parametrize("kernel", [0, 1, 2, 3, 4, 5, 7])
def test_single_step_kernels_contracted(oracle, kernel, dt):
    """the FMA-contracted arithmetic mode against the oracle compiled with gcc's contraction"""
    st = (SHAPES_R4 if G.kernel_info(kernel).r == 4 else SHAPES_R1)[0]
    ob = oracle_steps(oracle, kernel, st, dt, 2, contract=True)
    same(emu_run(oracle, kernel, st, dt, [1, 1], contract=True), ob)
    same(emu_run(oracle, kernel, st, dt, [1, 1], contract=True, variant=1), ob)


# ------------------------------------------------------------------------------------------------
# fused sweep k_r1: every depth and tile shape the launcher offers
# ------------------------------------------------------------------------------------------------
FUSED_TILES = {1: (0, 216, 408, 312, 310, 316, 5408, 5216, 7408, 7216, 9408, 9216), 2: (0, 216, 408, 9216, 9408), 3: (0, 216, 408, 9216, 9408), 5: (0, 216, 408, 9216, 9408)}


@pytest.mark.parametrize("dt", [F32, F64], ids=["sp", "dp"])
@pytest.mark.parametrize("kernel", [1, 2, 3, 5])
def test_fused_sweep(oracle, kernel, dt):
    tmax = G.kernel_info(kernel).max_tfuse
    for st in SHAPES_R1:
        for T in range(1, tmax + 1):
            nsteps = 2 * T + 1
            sizes = G.plan_fused_passes(nsteps, T)
            assert sum(sizes) == nsteps and sizes[-1] == 1
            ob = oracle_steps(oracle, kernel, st, dt, nsteps)
            for tile in FUSED_TILES[kernel]:
                for zchunk in (0, 7):
                    same(emu_run(oracle, kernel, st, dt, sizes, tile=tile, zchunk=zchunk, variant=2), ob)


@pytest.mark.parametrize("dt", [F32, F64], ids=["sp", "dp"])
def test_fused_sweep_contracted(oracle, dt):
    for kernel in (1, 2, 3, 5):
        T = G.kernel_info(kernel).max_tfuse
        sizes = G.plan_fused_passes(2 * T + 1, T)
        ob = oracle_steps(oracle, kernel, SHAPES_R1[0], dt, sum(sizes), contract=True)
        same(emu_run(oracle, kernel, SHAPES_R1[0], dt, sizes, contract=True), ob)


@pytest.mark.parametrize("T", [1, 2, 3, 4])
def test_fused_sweep_two_ranges(oracle, T):
    """one launch sweeps the two outer parts of a slab, a second one the middle (the halo-first
    overlap schedule of run_passes): together they must equal a full pass"""
    st, dt, kernel = (70, 30, 24), F64, 1
    pb = oracle.make_problem(kernel, st, dt)
    s = E.EmuStepper(kernel, st, pb.shape, dt, G.kernel_info(kernel))
    s.variant = 2
    s.upload(pb)
    zq = 7
    s.one_pass(T, 0, zq, st[2] - zq, st[2])
    s.one_pass(T, zq, st[2] - zq)
    s.cur ^= 1
    s.download(pb.U1, pb.U2)
    s.close()
    ob = oracle_steps(oracle, kernel, st, dt, T)
    # T steps from U2 (level 0) land in U1 = level T; the oracle leaves level T in U1 (T odd) or U2 (T even)
    want = ob.U1 if T % 2 == 1 else ob.U2
    assert pb.U1.tobytes() == want.tobytes()


# ------------------------------------------------------------------------------------------------
# scheduling-order independence: a missing barrier or a buffer reused too early shows up as a
# result that depends on the order in which fibers reach their synchronisation points
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [1, 2], ids=["reverse", "random"])
def test_schedule_order_independence(oracle, monkeypatch, order):
    monkeypatch.setenv("CUDA_EMU_SCHED", str(order))
    cases = [(1, F64, [4, 4, 1], dict(variant=2)), (1, F32, [4, 4, 1], dict(variant=2, tile=216)),
             (1, F64, [4, 4, 1], dict(variant=2, tile=5408)), (1, F32, [3, 3, 1], dict(variant=2, tile=5216)),
             (1, F64, [2, 2, 1], dict(variant=2, tile=5408, zchunk=6, contract=True)),
             (1, F64, [4, 4, 1], dict(variant=2, tile=7408)), (1, F32, [3, 3, 1], dict(variant=2, tile=7216, zchunk=5)),
             (1, F64, [2, 2, 1], dict(variant=2, tile=7408, contract=True)),
             (1, F64, [4, 4, 1], dict(variant=2, tile=9408)), (1, F32, [4, 4, 1], dict(variant=2, tile=9216, zchunk=5)),
             (1, F64, [4, 2, 1], dict(variant=2, tile=9216, contract=True)),
             (2, F64, [3, 3, 1], dict(variant=2)), (5, F32, [2, 2, 1], dict(variant=2)),
             (0, F64, [1, 1], {}), (0, F32, [1, 1], dict(tile=16)), (0, F32, [1, 1], dict(tile=116)),
             (4, F64, [1, 1], {}), (4, F32, [1, 1], dict(tile=8)), (7, F64, [1, 1], {}), (1, F64, [1, 1], {})]
    for kernel, dt, sizes, opt in cases:
        st = (70, 30, 20)
        same(emu_run(oracle, kernel, st, dt, sizes, **opt),
             oracle_steps(oracle, kernel, st, dt, sum(sizes), contract=bool(opt.get("contract", False))))


# ------------------------------------------------------------------------------------------------
# the emulated kernels against outputs of the unmodified reference (tests/golden, made by
# tests/golden/make_golden.py from oracle/_ref/ref_dump_*): same schedules as the host steppers
# ------------------------------------------------------------------------------------------------
def _golden_sizes(kernel, nt, ts, t_dim):
    if ts == 2:
        nt_eff = G.diamond_nt(nt, t_dim)
        T = min(G.kernel_info(kernel).max_tfuse, 4)
        return G.plan_fused_passes(nt_eff - 1, T), nt_eff
    return [1] * (2 * ((nt + 1) // 2)), nt   # nb_naive_ts.c:187-203: two steps per iteration


def _golden_case(O, key, table, contract):
    m = KEY.match(key)
    k, nx, ny, nz, nt, ts, td = (int(g) for g in m.groups()[:7])
    dt = F32 if m.group(8) == "sp" else F64
    sizes, nt_eff = _golden_sizes(k, nt, ts, td)
    if ts == 2:
        assert nt_eff == int(table[key + "_nteff"])
    pb = emu_run(O, k, (nx, ny, nz), dt, sizes, contract=contract)
    assert pb.interior().tobytes() == table[key].tobytes()


@pytest.mark.parametrize("key", sorted(k for k in SMALL.files if not k.endswith("_nteff")))
def test_reference_golden(oracle, key):
    _golden_case(oracle, key, SMALL, False)


@pytest.mark.parametrize("key", sorted(k for k in SMALL_FMA.files if not k.endswith("_nteff")))
def test_reference_golden_contracted(oracle, key):
    _golden_case(oracle, key, SMALL_FMA, True)
