"""Instruction mix of the hottest loop (the backward branch whose body holds the most FFMA2) of one kernel in a
`cuobjdump -sass` listing.  usage: python tools/sass_loop.py listing.sass <substring of the mangled kernel name>"""
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


def main():
    path, key = sys.argv[1], sys.argv[2]
    ins = [l for l in function_lines(path, key) if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l) and not l.strip().startswith("/* 0x")]
    addr = lambda l: int(re.search(r"/\*([0-9a-f]{4,})\*/", l).group(1), 16)
    op = lambda l: re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l).group(1)
    best = None
    for l in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", l)
        if not m:
            continue
        t, a = int(m.group(1), 16), addr(l)
        if t < a:
            body = [x for x in ins if t <= addr(x) <= a]
            n2 = sum("FFMA2" in x for x in body)
            if best is None or n2 > best[0]:
                best = (n2, t, a, body)
    whole = Counter(op(x).split(".")[0] for x in ins)
    print(f"kernel: {len(ins)} instructions; " + ", ".join(f"{k} {v}" for k, v in whole.most_common(14)))
    if best:
        n2, t, a, body = best
        c = Counter(op(x).split(".")[0] for x in body)
        print(f"hot loop {t:#x}..{a:#x}: {len(body)} instructions; " + ", ".join(f"{k} {v}" for k, v in c.most_common()))


main()
