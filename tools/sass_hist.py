#!/usr/bin/env python
"""Opcode histogram of one kernel's SASS (optionally of an address range): tools/sass_hist.py LIB SUBSTR [lo hi]
Used to check instruction mixes before spending GPU time (cuobjdump -sass)."""
import collections
import re
import subprocess
import sys


def main():
    lib, sub = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, hist, n = None, collections.Counter(), 0
    back = []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or sub not in cur:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        t = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\w+,\s*)?0x([0-9a-f]+)", line)
        if t and int(t.group(1), 16) < addr:
            back.append((addr, int(t.group(1), 16)))
        if lo <= addr < hi:
            hist[m.group(2)] += 1
            n += 1
    print(f"{n} instructions in [{lo:#x}, {hi:#x}); backward branches: " + ", ".join(f"{a:#x}->{b:#x}" for a, b in back))
    for op, c in hist.most_common():
        print(f"{c:6d} {op}")


if __name__ == "__main__":
    main()
