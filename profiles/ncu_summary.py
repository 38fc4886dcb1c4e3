#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel and the SASS opcode mix per thread (from the source page).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [warps_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "launch__waves_per_multiprocessor",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for k in KEYS:
            if k in hdr:
                print("   %-75s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        # warp-state breakdown: stalled warps per issued instruction, by reason (largest first)
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("   stalls per issue: " + "  ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:9]))
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    kern, data = None, {}
    for r in src:
        if r and r[0] == "Kernel Name":
            kern = r[1][:80]
            data[kern] = []
        elif r and r[0].startswith("0x"):
            data[kern].append(r)
    for k, rs in data.items():
        tot, n, thr = collections.Counter(), 0, 0
        for r in rs:
            toks = r[1].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            op = op.split(".")[0]
            tot[op] += int(r[5])
            n += int(r[5])
            thr += int(r[6])
        warps = thr / 32.0 / max(1, n) * 1.0
        # instructions per thread = thread-instructions / threads; threads unknown here -> use the STG count as sites
        sites = sum(int(r[6]) for r in rs if r[1].split()[0 if not r[1].split()[0].startswith("@") else 1].startswith("STG")) / 3.0
        print("== opcode mix per site (thread instructions / sites), %s" % k)
        print("   sites %.0f  total instr/site %.1f" % (sites, thr / max(1.0, sites)))
        per = collections.Counter()
        for r in rs:
            toks = r[1].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            per[op] += int(r[6])
        for op, c in per.most_common(28):
            print("   %-10s %8.1f" % (op, c / max(1.0, sites)))


if __name__ == "__main__":
    main()
