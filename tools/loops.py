"""Per-loop instruction mix of one kernel in a `cuobjdump -sass` listing (every backward branch = one loop).
usage: python tools/loops.py listing.sass <substring of the mangled kernel name>"""
import re
import sys
from collections import Counter


def function_lines(path, key):
    out, on = [], False
    for line in open(path):
        if "Function :" in line:
            on = key in line
            continue
        if on:
            out.append(line.rstrip("\n"))
    return out


ins = [l for l in function_lines(sys.argv[1], sys.argv[2]) if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l) and not l.strip().startswith("/* 0x")]
addr = lambda l: int(re.search(r"/\*([0-9a-f]{4,})\*/", l).group(1), 16)
op = lambda l: re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l).group(1)
print(f"{len(ins)} instructions")
for l in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", l)
    if not m:
        continue
    t, a = int(m.group(1), 16), addr(l)
    if t < a:
        body = [x for x in ins if t <= addr(x) <= a]
        c = Counter(op(x).split(".")[0] for x in body)
        main = ("FFMA2", "FMUL2", "FADD2", "MUFU", "LDS")
        rest = Counter({k: v for k, v in c.items() if k not in main})
        print(f"  loop {t:#x}..{a:#x} n={len(body)} " + " ".join(f"{k}={c[k]}" for k in main) + " | " +
              " ".join(f"{k}={v}" for k, v in rest.most_common(8)))
