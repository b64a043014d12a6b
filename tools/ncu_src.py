#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: top stalled SASS instructions
and an opcode histogram of the warp-stall samples.  usage: ncu_src.py file.csv [kernel-index]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    # the file holds one block per profiled launch: "Kernel Name" row, header row, data rows
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    s = starts[which]
    e = starts[which + 1] if which + 1 < len(starts) else len(rows)
    print(rows[s][1][:100])
    hdr = rows[s + 1]
    data = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda x: int(x) if x.isdigit() else 0
    tot = sum(num(r[isamp]) for r in data)
    print("total samples", tot, "instructions", len(data))
    for r in sorted(data, key=lambda r: -num(r[isamp]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
        st = sorted(((h, num(r[i])) for i, h in stall_cols if num(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print(r[isamp].rjust(7), r[iex].rjust(10), r[isrc][:72].ljust(72), st)
    hist, exh = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[isrc].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        hist[op] += num(r[isamp]); exh[op] += num(r[iex])
    print()
    for op, c in hist.most_common(25):
        print(op.ljust(12), str(c).rjust(8), f"{c / max(tot, 1) * 100:5.1f}%", str(exh[op]).rjust(12))
    agg = collections.Counter()
    for r in data:
        for i, h in stall_cols:
            agg[h] += num(r[i])
    print()
    print({k: v for k, v in agg.most_common(8)})


if __name__ == "__main__":
    main()
