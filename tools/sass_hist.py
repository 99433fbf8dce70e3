#!/usr/bin/env python
"""Histogram SASS opcodes per kernel of a .so/.cubin: python tools/sass_hist.py <file> [name-substring]"""
import re, subprocess, sys, collections
path = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur = None
hist = collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
for k, h in hist.items():
    if filt in k:
        tot = sum(h.values())
        print(f"== {k}  total={tot}")
        print("   " + ", ".join(f"{op}:{n}" for op, n in h.most_common(18)))
