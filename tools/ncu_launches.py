#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum,... --csv` launch list: one line per launch (second half of the
file = the timed step), or per-kernel totals with --sum."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault((int(r[iid]), r[ik]), {})[r[im]] = float(r[iv].replace(",", ""))
    items = list(per.items())
    if "--sum" in sys.argv:
        tot = collections.OrderedDict()
        for (i, k), m in items:
            name = k.split("(")[0][-40:]
            t = tot.setdefault(name, [0, 0.0])
            t[0] += 1
            t[1] += m["gpu__time_duration.sum"] / 1e3
        for k, (n, us) in tot.items():
            print(f"{k:42s} x{n:3d} {us:10.1f} us")
        return
    for (i, k), m in items:
        name = k.split("(")[0][-34:]
        print(f"{i:4d} {name:36s} {m['gpu__time_duration.sum'] / 1e3:9.1f} us  R {m.get('dram__bytes_read.sum', 0) / 1e6:7.0f} MB  "
              f"W {m.get('dram__bytes_write.sum', 0) / 1e6:7.0f} MB  inst {m.get('smsp__inst_executed.sum', 0) / 1e6:7.1f} M")


if __name__ == "__main__":
    main()
