#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: registers / spills / smem per kernel (developer tool).
usage: ptxas_report.py LOG [regex]"""
import re, subprocess, sys
log = open(sys.argv[1]).read()
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
blocks = re.split(r"ptxas info\s+: Compiling entry function '", log)[1:]
for b in blocks:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = dem.replace("dfft::", "").replace("void ", "")
    spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    regs = re.search(r"Used (\d+) registers", b)
    if pat and not pat.search(dem):
        continue
    print(f"regs={regs.group(1):>3} stack={spill.group(1):>4} spill_st={spill.group(2):>4}  {dem[:230]}")
