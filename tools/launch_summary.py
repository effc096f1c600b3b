#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, ni, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki][:64]
        v = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        if r[ni] == "gpu__time_duration.sum":
            a["n"] += 1
            a["us"] += v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
        else:
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(r[ui], 1e-6)
            a["rd" if "read" in r[ni] else "wr"] += v * scale
    tot = sum(v["us"] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = max(v["n"], 1)
        print(f"{v['us']:10.1f} us {v['n']:4d}x {100 * v['us'] / tot:5.1f}%  avg {v['us'] / n:8.1f} us  dram rd/wr {v['rd'] / n:7.1f}/{v['wr'] / n:7.1f} MB  {k}")
    print(f"{tot:10.1f} us total")


if __name__ == "__main__":
    main(sys.argv[1])
