"""Per-source-line thread-instruction counts and stall samples of the first kernel in an `ncu --set full
--import-source on` report.  usage: python tools/ncu_lines.py <report.ncu-rep> <pixels-per-launch> [top]"""
import csv
import io
import subprocess
import sys

rep, npx = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, per = None, None, {}
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 3 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        try:
            per[(cur, int(r[0]))] = (int(r[hdr.index("Thread Instructions Executed")]), int(r[hdr.index("# Samples")]), r[1],
                                     int(r[hdr.index("stall_barrier")]), int(r[hdr.index("stall_long_sb")]))
        except ValueError:
            pass
tot = sum(v[0] for v in per.values())
tot_s = sum(v[1] for v in per.values())
print(f"thread instructions / px: {tot / npx:.1f}; samples {tot_s}; barrier-stall samples {sum(v[3] for v in per.values()) / tot_s:.1%}; "
      f"long-scoreboard {sum(v[4] for v in per.values()) / tot_s:.1%}")
for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f[:14]:14s}{ln:5d} {v[0] / npx:6.2f}/px samples {v[1] / tot_s:5.1%} barrier {v[3] / tot_s:5.1%}  {v[2].strip()[:100]}")
