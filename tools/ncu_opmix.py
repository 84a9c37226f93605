"""Opcode mix and stall-sample share of a kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME` (development tool)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Instructions Executed" in r)
ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ist = hdr.index("Warp Stall Sampling (All Samples)")
cnt = collections.Counter(); st = collections.Counter(); tot = 0; tots = 0
for r in rows:
    if len(r) <= ia or not r[ia].isdigit():
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
    if not m:
        continue
    op = m.group(2).split('.')[0]
    n = int(r[ia]); s = int(r[ist]) if r[ist].isdigit() else 0
    cnt[op] += n; st[op] += s; tot += n; tots += s
print("total warp-instr", tot, "samples", tots)
for op, n in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:10s} {n:10d} {100 * n / tot:5.1f}%   stall samples {100 * st[op] / max(tots, 1):5.1f}%")
