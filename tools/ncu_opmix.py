"""Aggregates an `ncu --page source --csv` export by SASS opcode: executed warp-instructions per pixel.
usage: ncu -i rep.ncu-rep --page source --csv | python tools/ncu_opmix.py <pixels> [top]"""
import csv, sys, collections
npx = float(sys.argv[1]); top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
isrc, iex, ith, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
for r in rows[hi + 1:]:
    if len(r) <= ith: continue
    s = r[isrc].strip().split()
    if not s: continue
    op = s[1] if s[0].startswith("@") and len(s) > 1 else s[0]
    op = op.rstrip(";")
    n = int(r[iex] or 0) * 32
    ops[op] += n; tot += n
    smp[op] += int(r[ismp] or 0); tots += int(r[ismp] or 0)
print(f"total thread-instr slots / px: {tot / npx:.1f}")
for op, n in ops.most_common(top):
    print(f"{op:28s} {n / npx:7.2f} /px   samples {100.0 * smp[op] / max(tots,1):5.1f}%")
