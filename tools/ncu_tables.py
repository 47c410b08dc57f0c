#!/usr/bin/env python
"""Markdown tables from an ncu report: `python tools/ncu_tables.py gpurun_out/r1_v9.ncu-rep` (runs `ncu -i ... --page raw --csv`
and `--page source --csv` per kernel; needs ncu on PATH).  Used to write profiles/*_ncu_summary.md."""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.split("(")[0]
        print("**%s**\n\n| metric | value |\n|---|---|" % short)
        for m in METRICS:
            i = hdr.index(m)
            print("| %s | %s %s |" % (m, r[i], units[i]))
        key = short.replace("void ", "").split("<")[0]
        src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--kernel-name", "regex:" + key))))
        h = src[1]
        cols = [(c, i) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = collections.Counter()
        ops = collections.Counter()
        isrc, iex = h.index("Source"), h.index("Instructions Executed")
        for s in src[2:]:
            for c, i in cols:
                try:
                    tot[c] += int(s[i])
                except (ValueError, IndexError):
                    pass
            try:
                t = s[isrc].split()
                op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                ops[op] += int(s[iex])
            except (ValueError, IndexError):
                pass
        n = float(sum(tot.values())) or 1.
        print("| stall reasons (pc sampling) | %s |" % ", ".join("%s %.1f%%" % (c[6:], 100 * v / n) for c, v in tot.most_common(7)))
        n = float(sum(ops.values())) or 1.
        print("| executed instruction mix | %s |" % ", ".join("%s %.1f%%" % (c, 100 * v / n) for c, v in ops.most_common(8)))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
