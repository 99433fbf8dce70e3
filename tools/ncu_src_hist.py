#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass` dump: executed warp-instructions by opcode,
for the first kernel in the file.  usage: ncu_src_hist.py src.csv [words_per_launch]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); it = hdr.index("Avg. Threads Executed")
ist = hdr.index("Warp Stall Sampling (All Samples)")
ops = collections.Counter(); stalls = collections.Counter(); tot = 0
lines = []
for r in rows[2:]:
    if len(r) <= ie or r[0].startswith("Kernel Name"):
        if r and r[0].startswith("Kernel Name") and tot: break
        continue
    try: n = int(r[ie])
    except ValueError: continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", r[ia])
    if not m: continue
    op = m.group(1) + ("" if m.group(1) not in ("IMAD","LOP3","SHF") else "".join(x for x in m.group(2).split(".") if x in ("WIDE","MOV","IADD","HI","SHL","X")) and "." + ".".join(x for x in m.group(2).split(".") if x in ("WIDE","MOV","IADD","HI","SHL","X")))
    ops[op] += n; tot += n
    stalls[op] += int(r[ist] or 0)
    lines.append((n, int(r[ist] or 0), r[ia].strip()))
div = float(sys.argv[2]) if len(sys.argv) > 2 else None
print("total warp-inst", tot, ("per warp-word-row %.1f" % (tot/div)) if div else "")
for op, n in ops.most_common(25):
    print(f"{op:14s} {n:14d} {100*n/tot:5.1f}%  stall-samples {stalls[op]}" + (f"  per-unit {n/div:.2f}" if div else ""))
