#!/usr/bin/env python
"""Per-launch table (DRAM bytes, duration) from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log.
usage: python profiles/launch_table.py gpurun_out/launches.csv [max_rows]"""
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, mi, mn, gi, bi, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "Grid Size", "Block Size", "ID"))
cur = {}
for r in rows[1:]:
    cur.setdefault((r[ii], r[ki].split("(")[0][-40:], r[gi], r[bi]), {})[r[mn]] = float(r[mi].replace(",", ""))
for k, v in list(cur.items())[: int(sys.argv[2]) if len(sys.argv) > 2 else 1000]:
    rd, wr, t = v.get("dram__bytes_read.sum", 0), v.get("dram__bytes_write.sum", 0), v["gpu__time_duration.sum"]
    print("%-42s %-16s %-14s rd %7.0f MB  wr %7.0f MB  %8.3f ms  %6.0f GB/s" % (k[1], k[2], k[3], rd / 1e6, wr / 1e6, t / 1e6, (rd + wr) / t))
