#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell path (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), from `cuobjdump -sass` of the built library.  Usage: python scripts/sass_opcodes.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "openhush_b200", "csrc", "libohwhisper.so")
ops = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "UCGABAR"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        order.append(cur)
        continue
    if cur:
        for op in ops:
            if re.search(r"\b" + op + r"\b|\b" + op + r"\.", line):
                counts[cur][op] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a): occurrences per kernel; kernels without any of these opcodes omitted")
print("kernel".ljust(78) + "".join(o.rjust(9) for o in ops))
for k in order:
    if any(counts[k].values()):
        print(k[:77].ljust(78) + "".join(str(counts[k][o] or "").rjust(9) for o in ops))
