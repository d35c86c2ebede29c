#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` output: registers / stack / spills per kernel."""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
pat = re.compile(r"Compiling entry function '([^']+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers")
names = [m[0] for m in pat.findall(txt)]
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().split("\n") if names else []
for (name, stack, ss, sl, regs), d in zip(pat.findall(txt), dem):
    m = re.search(r"ps_kernel<(?:\(anonymous namespace\)::)?(ps::)?([A-Za-z0-9_]+(?:<[^>]*>)?)", d)
    short = m.group(2) if m else d.split("(")[0][-60:]
    print("%-40s regs=%-4s stack=%-5s spill=%s/%s" % (short, regs, stack, ss, sl))
