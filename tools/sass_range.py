"""opcode histogram of an address range of a SASS dump: python tools/sass_range.py file.sass 0x15e0 0x3650"""
import sys, collections
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
ops = collections.Counter(); n = 0
for line in open(sys.argv[1]):
    parts = line.split()
    if len(parts) > 2 and parts[0].startswith("/*") and len(parts[0]) == 8:
        a = int(parts[0][2:6], 16)
        if lo <= a < hi:
            op = parts[1] if not parts[1].startswith("@") else parts[2]
            ops[op.rstrip(";")] += 1; n += 1
print(n, "instructions")
print(", ".join(f"{k} {v}" for k, v in ops.most_common(40)))
