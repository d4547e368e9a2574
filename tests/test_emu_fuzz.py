"""A short fixed-seed slice of the emulator fuzzers (tests/cuda_emu/fuzz.py): random operators, shapes, depths,
tiles, z chunkings, arithmetic modes, topologies and stepper options, every result bit-exact vs the oracle."""
import pytest

from cuda_emu import fuzz


@pytest.mark.parametrize("seed", [101, 102])
def test_fuzz_kernels(oracle, seed, monkeypatch):
    monkeypatch.setenv("CUDA_EMU_SCHED", str(seed % 3))
    msgs = []
    n, bad = fuzz.fuzz_kernels(seed, seconds=60, max_cases=40, log=lambda *a: msgs.append(a))
    assert n > 0 and bad == 0, msgs


@pytest.mark.parametrize("seed", [201])
def test_fuzz_library(oracle, seed):
    msgs = []
    n, bad = fuzz.fuzz_library(seed, seconds=60, max_cases=25, log=lambda *a: msgs.append(a))
    assert n > 0 and bad == 0, msgs
