"""Summarise ncu output for profiles/: `python scripts/ncu_summary.py TAG [launches.csv] [report.ncu-rep]`.

Writes profiles/TAG_launches.md (per-kernel launch count, total/avg device time, share of the step) from the
`--metrics gpu__time_duration.sum` CSV and profiles/TAG_kernels.md (one block per captured launch of the
`--set full` report: duration, DRAM bytes, pipe utilisation, occupancy, registers, instruction count)."""
import collections
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe (LSU wavefronts) % of peak"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 sectors, global loads"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps per issue)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
]


def launches(tag, path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif row["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open("profiles/%s_launches.md" % tag, "w") as out:
        out.write("# %s: every launch of the bench command under `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n" % tag)
        out.write("(cold-cache, serialised: compare SHARES, not absolutes)\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            out.write("| `%s` | %d | %.1f | %.2f | %.3f |\n" % (k[:110], a[0], a[1] / 1e3, a[1] / a[0] / 1e3, a[1] / tot))


def kernels(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open("profiles/%s_kernels.md" % tag, "w") as out:
        out.write("# %s: `ncu --set full --clock-control none --import-source on` captures (raw page)\n" % tag)
        for r in rows[2:]:
            out.write("\n## `%s`  (launch id %s)\n\n| metric | value |\n|---|---|\n" % (r[hdr.index("Kernel Name")][:120], r[0]))
            for m, label in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    out.write("| %s (`%s`) | %s %s |\n" % (label, m, r[i], units[i]))


if __name__ == "__main__":
    tag = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] != "-":
        launches(tag, sys.argv[2])
    if len(sys.argv) > 3:
        kernels(tag, sys.argv[3])
