"""Lists the loops of a kernel's SASS (backward branches) with their instruction mix and spill counts.
usage: python tools/sass_loops.py <object.o> <function-name-substring> [min_len]"""
import re
import subprocess
import sys

obj, fn = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 150
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
out, on = [], False
for line in sass.split("\n"):
    if "Function :" in line:
        on = fn in line
    if on:
        out.append(line)
pat = re.compile(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);")
ins = [(int(m.group(1), 16), m.group(2)) for m in (pat.search(l) for l in out) if m]
amap = {a: i for i, (a, _) in enumerate(ins)}
print(f"{fn}: {len(ins)} instructions")
for i, (a, t) in enumerate(ins):
    if "BRA" in t:
        m = re.search(r"0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in amap:
            s = amap[int(m.group(1), 16)]
            body = [x for _, x in ins[s:i + 1]]
            if min_len < i - s < 4000:
                c = lambda k: sum(k in x for x in body)
                print(f"loop {ins[s][0]:#07x}-{a:#07x} len {i - s:5d} LDG {c('LDG'):3d} LDS {c('LDS'):3d} FFMA {c('FFMA'):3d} FFMA2 {c('FFMA2'):3d} "
                      f"PRMT {c('PRMT'):3d} FSET {c('FSET'):3d} LOP3 {c('LOP3'):3d} ATOMS {c('ATOMS'):3d} spill {c('STL') + c('LDL'):3d} BRA {c('BRA'):3d}")
