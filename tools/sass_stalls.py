#!/usr/bin/env python3
"""Decode the scheduling control fields (stall count, yield, barriers) from `cuobjdump -sass` output of one function and print, per
opcode class, the instruction count and the sum of encoded stall cycles = the issue time of one warp running alone, scoreboard waits excluded.
usage: sass_stalls.py file.sass function_substring [start_addr end_addr]"""
import collections
import re
import sys


def parse(path, fn):
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if "Function :" in l and fn in l)
    out = []
    i = start + 1
    while i < len(lines) and "Function :" not in lines[i]:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                txt = m.group(2).strip()
                op = re.sub(r"^@!?U?P\d+\s+", "", txt).split()[0]
                out.append({"addr": int(m.group(1), 16), "op": op, "txt": txt, "stall": (hi >> 41) & 0xF, "yield": (hi >> 45) & 1,
                            "wbar": (hi >> 46) & 7, "rbar": (hi >> 49) & 7, "wait": (hi >> 52) & 0x3F})
                i += 1
        i += 1
    return out


def main():
    ins = parse(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 4:
        a, b = int(sys.argv[3], 16), int(sys.argv[4], 16)
        ins = [x for x in ins if a <= x["addr"] < b]
    by = collections.defaultdict(lambda: [0, 0])
    for x in ins:
        k = "IMAD.WIDE" if x["op"].startswith("IMAD.WIDE") else x["op"].split(".")[0]
        by[k][0] += 1; by[k][1] += x["stall"]
    tot = sum(v[1] for v in by.values())
    print("instructions", len(ins), "sum of stall counts", tot)
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:12]:
        print("  %-10s n=%5d stall_sum=%6d avg=%.2f" % (k, v[0], v[1], v[1] / v[0]))
    if "--dump" in sys.argv:
        for x in ins:
            print("%04x st=%2d y=%d w=%d r=%d wm=%02x  %s" % (x["addr"], x["stall"], x["yield"], x["wbar"], x["rbar"], x["wait"], x["txt"][:90]))


if __name__ == "__main__":
    main()
