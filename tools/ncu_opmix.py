"""Aggregates an `ncu --page source --csv` export by SASS opcode: executed warp-instruction issue slots per pixel.
usage: ncu -i rep.ncu-rep --page source --csv | python tools/ncu_opmix.py <pixels> [top] [section]
A report with several kernels prints one section per kernel (each starts with its own header row); `section` (0-based,
default 0) picks one."""
import collections
import csv
import sys

npx = float(sys.argv[1])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(sys.stdin))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if not heads:
    sys.exit("no SASS section in the input")
hi = heads[min(want, len(heads) - 1)]
end = heads[heads.index(hi) + 1] if heads.index(hi) + 1 < len(heads) else len(rows)
hdr = rows[hi]
isrc, iex, ith, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
for r in rows[hi + 1:end]:
    if len(r) <= ith:
        continue
    s = r[isrc].strip().split()
    if not s:
        continue
    op = s[1] if s[0].startswith("@") and len(s) > 1 else s[0]
    op = op.rstrip(";")
    try:
        n = int(r[iex] or 0) * 32
    except ValueError:
        continue
    ops[op] += n; tot += n
    smp[op] += int(r[ismp] or 0); tots += int(r[ismp] or 0)
print(f"total thread-instr slots / px: {tot / npx:.1f}")
for op, n in ops.most_common(top):
    print(f"{op:28s} {n / npx:7.2f} /px   samples {100.0 * smp[op] / max(tots,1):5.1f}%")
