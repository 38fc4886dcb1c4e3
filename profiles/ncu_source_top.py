#!/usr/bin/env python
"""Top stall-sample instructions of one kernel from `ncu -i rep --page source --csv --kernel-name regex:NAME > file.csv`.
usage: python profiles/ncu_source_top.py file.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
blocks = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        blocks.append([])
    elif hdr and len(r) == len(hdr) and r[0].startswith("0x"):
        blocks[-1].append(r)
for data in blocks[:1]:
    si, src, ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    sh_ex, sh = hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("L1 Wavefronts Shared")
    tot = sum(int(r[si]) for r in data)
    print("total samples", tot, "instructions", len(data), "executed warp-instr", sum(int(r[ex]) for r in data))
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stall_cols}
    print("stall totals:", "  ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -int(r[si]))[:n_top]:
        st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print("%6d %5.1f%% ex=%9s shw=%9s/%9s  %-58s %s" % (int(r[si]), 100 * int(r[si]) / tot, r[ex], r[sh], r[sh_ex], r[src].strip()[:58], st))
    print("shared wavefronts total", sum(int(r[sh]) for r in data), "excessive", sum(int(r[sh_ex]) for r in data))
