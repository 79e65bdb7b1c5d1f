#!/usr/bin/env python3
"""Aggregate the `ncu --page source --csv` export of one kernel by SASS opcode class: share of executed instructions,
share of stall samples, samples per executed instruction (relative), and the top stall reasons of each class."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ix = {k: i for i, k in enumerate(hdr)}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    by = collections.defaultdict(collections.Counter)
    tot = collections.Counter()
    num = lambda x: int(float(x)) if x not in ("", "-") else 0
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[0] in ("Address", "Kernel Name"):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = m.group(2) if m else "?"
        cls = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else op.split(".")[0]
        n, ex = num(r[ix["# Samples"]]), num(r[ix["Instructions Executed"]])
        by[cls]["samples"] += n; by[cls]["exec"] += ex
        tot["samples"] += n; tot["exec"] += ex
        for s in stalls:
            v = num(r[ix[s]]); by[cls][s] += v; tot[s] += v
    print("total samples", tot["samples"], "warp instructions", tot["exec"])
    print({k[6:]: round(v / tot["samples"], 3) for k, v in tot.items() if k.startswith("stall") and v > 0.01 * tot["samples"]})
    for cls, c in sorted(by.items(), key=lambda kv: -kv[1]["samples"])[:16]:
        top = sorted(((s, c[s]) for s in stalls), key=lambda x: -x[1])[:4]
        print("%-10s exec=%.3f samples=%.3f rel_cost=%.2f " % (cls, c["exec"] / tot["exec"], c["samples"] / tot["samples"],
              c["samples"] / max(c["exec"], 1) * tot["exec"] / tot["samples"]), [(s[6:], round(v / max(c["samples"], 1), 2)) for s, v in top])


if __name__ == "__main__":
    main(sys.argv[1])
