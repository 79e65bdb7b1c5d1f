#!/usr/bin/env python3
"""Attribute the `ncu --page source --csv` (SASS view) export of one kernel to the out-of-line device functions inside it.
    cuobjdump -elf lib.so > elf.txt ;  ncu -i rep --page source --csv --kernel-name K > src.csv ;  python tools/ncu_by_function.py src.csv elf.txt K
Per function: share of executed warp instructions, share of stall samples, IMAD.WIDE share inside it, top stall reasons."""
import collections, csv, re, sys


def main(src, elf, kernel):
    syms = []
    for line in open(elf):
        m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+\d+\s+0x[0-9a-f]+\s+\$(\S+?)\$(\S+)", line)
        if m and kernel in m.group(3):
            syms.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
    syms.sort()
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    num = lambda x: int(float(x)) if x not in ("", "-") else 0
    base = None; seen = set()
    by = collections.defaultdict(collections.Counter); tot = collections.Counter()
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[0] in ("Address", "Kernel Name") or r[0] in seen: continue
        seen.add(r[0])
        a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
        if base is None: base = a
        off = a - base
        fn = "(kernel body)"
        for s, sz, name in syms:
            if s <= off < s + sz: fn = name; break
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]]); op = m.group(2) if m else "?"
        n, ex = num(r[ix["# Samples"]]), num(r[ix["Instructions Executed"]])
        c = by[fn]; c["samples"] += n; c["exec"] += ex; c["static"] += 1
        if op.startswith("IMAD.WIDE"): c["wide"] += ex
        elif op.startswith("IMAD"): c["imad_other"] += ex
        elif op.startswith("IADD3"): c["iadd3"] += ex
        elif op.split(".")[0] in ("MOV", "SEL"): c["mov_sel"] += ex
        elif op.split(".")[0] in ("LDS", "STS", "LDL", "STL", "LDG", "STG", "LD", "ST"): c["mem"] += ex
        for s in stalls: c[s] += num(r[ix[s]])
        tot["samples"] += n; tot["exec"] += ex
    print("total warp instructions %d, samples %d" % (tot["exec"], tot["samples"]))
    print("%-44s %7s %7s %6s %6s %6s %6s %6s %6s  stalls" % ("function", "exec%", "time%", "wide%", "imadO%", "iadd3%", "movsel", "mem%", "static"))
    for fn, c in sorted(by.items(), key=lambda kv: -kv[1]["samples"]):
        e = max(c["exec"], 1)
        top = sorted(((s[6:], c[s]) for s in stalls), key=lambda x: -x[1])[:4]
        print("%-44s %7.2f %7.2f %6.1f %6.1f %6.1f %6.1f %6.1f %6d  %s" % (fn[:44], 100 * c["exec"] / tot["exec"], 100 * c["samples"] / tot["samples"], 100 * c["wide"] / e,
              100 * c["imad_other"] / e, 100 * c["iadd3"] / e, 100 * c["mov_sel"] / e, 100 * c["mem"] / e, c["static"], [(s, round(v / max(c["samples"], 1), 2)) for s, v in top]))


if __name__ == "__main__":
    main(*sys.argv[1:4])
