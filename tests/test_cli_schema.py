"""stdout of build*/mwd_kernel against the field rules of the reference's result parser
(scripts/parse.py:95-245, `get_summary`): every `Key: value` line that parser extracts must be present with the
type it converts to -- float fields via float(), int fields via int(), the domain / topology lines by position.
The parser below is a Python-3 restatement of those rules written for this test (the reference script is
Python 2 and is not imported)."""
import re

import numpy as np
import pytest

import girih_b200 as G

pytestmark = pytest.mark.gpu

FLOAT_FIELDS = ("RANK0 MStencil/s  MIN", "RANK0 MStencil/s  AVG", "RANK0 MStencil/s  MAX",
                "MWD main-loop RANK0 MStencil/s MIN", "MWD main-loop RANK0 MStencil/s MAX",
                "Total RANK0 MStencil/s MIN", "Total RANK0 MStencil/s MAX",
                "RANK0 Total", "RANK0 Computation", "RANK0 Communication", "RANK0 Waiting", "RANK0 Other",
                "MAX Total", "MAX Computation", "MAX Communication", "MAX Waiting", "MAX Other",
                "MIN Total", "MIN Computation", "MIN Communication", "MIN Waiting", "MIN Other",
                "MEAN Total", "MEAN Computation", "MEAN Communication", "MEAN Waiting", "MEAN Other",
                "RANK0 ts main loop")
STR_FIELDS = ("Time stepper name", "Stencil Kernel name", "Stencil Kernel coefficients", "Precision",
              "Wavefront parallel strategy")
INT_FIELDS = ("Number of time steps", "Alignment size", "Number of tests", "Verify", "Time unroll",
              "Intra-diamond width", "OpenMP Threads", "MPI size", "Stencil Kernel semi-bandwidth",
              "Multi-wavefront updates", "Thread group size", "Intra-diamond prologue/epilogue MStencils",
              "Block size in X", "User set thread group size")


def parse_summary(text):
    """the extraction rules of get_summary(): `^field:` -> first token after the first colon"""
    out = {}
    for line in text.splitlines():
        for f in FLOAT_FIELDS:
            if re.match("^" + re.escape(f) + ":", line):
                out[f] = float(line.split(":")[1].split()[0])
        for f in INT_FIELDS:
            if re.match("^" + re.escape(f) + ":", line):
                out[f] = int(line.split(":")[1].split()[0])
        for f in STR_FIELDS:
            if re.match("^" + re.escape(f) + ":", line):
                out[f] = line.split(":")[1].strip()
        if "Assumed usable cache size" in line:
            out["cache size"] = int(line.split(":")[1].strip().split("K")[0])
        if "Global domain" in line:
            d = line.split()[3:6]
            out["Global N"] = tuple(int(x.split(":")[1]) for x in d)
        if "Rank 0 domain" in line:
            d = line.split()[4:7]
            out["Local N"] = tuple(int(x.split(":")[1]) for x in d)
        if "Processors topology" in line:
            out["topology"] = tuple(int(x) for x in line.split(":")[1].strip().split(","))
    return out


def test_diamond_run_parses_like_the_reference_output():
    rc, out, err = G.run_reference_cli(np.float64, ["--nx", 128, "--ny", 128, "--nz", 96, "--nt", 52,
                                                    "--target-ts", 2, "--target-kernel", 1, "--t-dim", 7,
                                                    "--n-tests", 2])
    assert rc == 0, out + err
    s = parse_summary(out)
    assert s["Time stepper name"] == "Diamond" and s["Precision"] == "DP"
    assert s["Number of time steps"] == 66 and s["Number of tests"] == 2 and s["Verify"] == 0
    assert s["Global N"] == (128, 128, 96) and s["Local N"] == (128, 128, 96) and s["topology"] == (1, 1, 1)
    assert s["Stencil Kernel semi-bandwidth"] == 1 and s["Time unroll"] == 7 and s["Intra-diamond width"] == 16
    assert s["MPI size"] == 1 and "cache size" in s
    # the Diamond stepper's result block (src/utils.c:906-912): totals, main loop, no per-phase breakdown
    for f in ("Total RANK0 MStencil/s MIN", "Total RANK0 MStencil/s MAX", "MWD main-loop RANK0 MStencil/s MIN",
              "MWD main-loop RANK0 MStencil/s MAX", "RANK0 ts main loop"):
        assert s[f] >= 0.0, f
    assert s["Total RANK0 MStencil/s MAX"] > 1000.0          # MStencil/s: a GPU is far above 1 GLUP/s


def test_spatial_blocking_run_parses():
    rc, out, err = G.run_reference_cli(np.float32, ["--nx", 128, "--ny", 64, "--nz", 64, "--nt", 20, "--n-tests", 2,
                                                    "--target-kernel", 0])
    assert rc == 0, out + err
    s = parse_summary(out)
    assert s["Time stepper name"] == "Spatial Blocking" and s["Precision"] == "SP"
    assert s["Stencil Kernel semi-bandwidth"] == 4 and s["Global N"] == (128, 64, 64)
    # the reference prints GStencil/s for ts 0/1 although its parser looks for MStencil/s:
    # the line is kept exactly as the reference binary prints it
    assert re.search(r"^RANK0 GStencil/s    MAX: +\d", out, re.M)
    for f in ("RANK0 Total", "RANK0 Computation", "RANK0 Communication", "RANK0 Waiting", "RANK0 Other",
              "MEAN Total", "MEAN Computation", "MAX Total", "MAX Computation", "MIN Total", "MIN Communication"):
        assert s[f] >= 0.0, f
    assert s["RANK0 Total"] > 0 and s["RANK0 Computation"] > 0
