#!/usr/bin/env python3
"""Per-source-line instruction counts of one kernel from an ncu report captured with --import-source on:
    ncu -i REPORT.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_source_profile.py src.csv [N]
Prints the N hottest lines (warp instructions executed and stall samples)."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur, hdr, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) == 2:
            continue
        if r and r[0] == "Line No":
            hdr = r
            ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) < 10:
            continue
        if r[2] == "-" and r[0].isdigit():
            n = int(r[ii]) if r[ii].isdigit() else 0
            s = int(r[isamp]) if r[isamp].isdigit() else 0
            if n or s:
                agg[(cur, int(r[0]))] = (n, s, r[1][:100])
    total = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {total}")
    for (f, l), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f[:16]:16s} {l:5d} {n / total * 100:5.1f}% inst {s / ts * 100:5.1f}% smp  {src}")


if __name__ == "__main__":
    main()
