#!/usr/bin/env python
"""Dynamic SASS opcode mix of one kernel from `ncu --page source --csv` output.

usage: ncu -i rep.ncu-rep --page source --csv --kernel-id :::2 > src.csv
       python profiles/ncu_source_mix.py src.csv [top]
"""
import collections
import csv
import re
import sys


def main() -> None:
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    i_s, i_e, i_n = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot, samp = collections.Counter(), collections.Counter()
    first = None
    for r in rows[hi + 1:]:
        if len(r) <= i_e or not r[i_e].isdigit():
            continue
        src = r[i_s].strip()
        ex, sm = int(r[i_e]), int(r[i_n] or 0)
        if first is None:
            first = ex
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
        op = m.group(2) if m else src.split()[0]
        tot[op] += ex
        samp[op] += sm
    total = sum(tot.values())
    print(f"warps launched (executions of the first instruction): {first}")
    print(f"warp-instructions per warp: {total / first:.1f}; stall samples: {sum(samp.values())}")
    for op, c in tot.most_common(top):
        print(f"  {op:10s} {c / first:8.2f} per warp   {100.0 * c / total:5.1f} %   samples {samp[op]}")


if __name__ == "__main__":
    main()
