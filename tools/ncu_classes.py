#!/usr/bin/env python
"""Stall samples of an ncu report grouped by opcode: where the warps' time goes.  python tools/ncu_classes.py x.ncu-rep"""
import collections
import csv
import subprocess
import sys

src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError):
        return 0.0


cls, ex = collections.Counter(), collections.Counter()
stall_by = collections.defaultdict(collections.Counter)
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    s = r[ix["Source"]].strip()
    op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
    cls[op] += f(r, "# Samples")
    ex[op] += f(r, "Instructions Executed")
    for k in keys:
        stall_by[op][k] += f(r, k)
tot = sum(cls.values())
print("total samples %d, warp instructions %d" % (tot, sum(ex.values())))
for op, c in cls.most_common(16):
    top = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in stall_by[op].most_common(4))
    print("%-8s samples %6d (%4.1f%%) exec %9d | %s" % (op, c, 100 * c / tot, ex[op], top))
