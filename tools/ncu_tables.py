#!/usr/bin/env python
"""Markdown tables from an ncu report: `python tools/ncu_tables.py gpurun_out/r1_v9.ncu-rep` (runs `ncu -i ... --page raw --csv`
and `--page source --csv` per kernel; needs ncu on PATH).  Used to write profiles/*_ncu_summary.md.

`python tools/ncu_tables.py --counts units.json report.ncu-rep > profiles/kernel_counts.json`: executed FP64 operations per
work unit of the three stage kernels (thread-level opcode counts of the source page: flop = 2 DFMA + DMUL + DADD), with the
units of the profiled launch from tools/profile_tile.py --units; bench.py turns them into achieved FP64 flop/s."""
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


def counts(units_json, rep):
    import json
    units = json.load(open(units_json))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr = rows[0]
    out = {"source": "%s (ncu --set full, config 5, one %d-ray tile launch per kernel: %d (ray, point) pairs); opcode counts "
                     "from the source page, thread level: flop = 2 DFMA + DMUL + DADD per unit" % (
                         rep, units["rays"], units["points"]), "units": units}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        key = name.split("<")[0].split("::")[-1]
        nunits = units["point_freqs"] if key == "k_transfer" else units["points"]
        src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--kernel-name", "regex:" + key))))
        h = src[1]
        isrc, ith = h.index("Source"), h.index("Thread Instructions Executed")
        ops = collections.Counter()
        for s_ in src[2:]:
            try:
                t = s_[isrc].split()
                op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                ops[op] += int(s_[ith])
            except (ValueError, IndexError):
                pass
        scale = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}

        def g(m):  # (bytes come back in GB, times in ms, whatever unit ncu chose for the column)
            i = hdr.index(m)
            return float(r[i]) * scale.get(rows[1][i], 1.0)
        fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
        out[key] = {"kernel": name, "unit": "(point, frequency)" if key == "k_transfer" else "(ray, point)",
                    "units_per_launch": nunits,
                    "fp64_flop_per_unit": round((2 * ops["DFMA"] + ops["DMUL"] + ops["DADD"]) / nunits, 1),
                    "fp64_instr_per_unit": round(fp64 / nunits, 1),
                    "instr_per_unit": round(sum(ops.values()) / nunits, 1),
                    "fp64_pipe_busy_pct": round(g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 1),
                    "issue_active_pct": round(g("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                    "registers": int(g("launch__registers_per_thread")),
                    "dram_read_gb": round(g("dram__bytes_read.sum"), 3), "dram_write_gb": round(g("dram__bytes_write.sum"), 3),
                    "dram_bytes_per_launch": round((g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) * 1e9),
                    "launch_ms_under_ncu": round(g("gpu__time_duration.sum"), 2)}
    out["k_geo_points"]["alg_write_gb"] = round(42 * units["points"] / 1e9, 2)
    out["k_fluid_ortho"]["alg_write_gb"] = round(128 * units["points"] / 1e9, 2)
    out["k_transfer"]["alg_read_gb"] = round(128 * units["points"] / 1e9, 2)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "--counts":
        counts(sys.argv[2], sys.argv[3])
    else:
        main(sys.argv[1])
