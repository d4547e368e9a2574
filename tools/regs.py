#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: registers / stack / spills per kernel."""
import re, subprocess, sys
log = open(sys.argv[1]).read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers")
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for name, stack, ss, sl, regs in pat.findall(log):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("void girih::", "")
    if flt in dem:
        print(f"{dem:48s} regs={regs:>3} stack={stack:>4} spill={ss}/{sl}")
