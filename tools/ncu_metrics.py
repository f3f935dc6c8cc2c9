"""Key metrics of every kernel in an ncu report: python tools/ncu_metrics.py report.ncu-rep  (runs `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if any(h == w or h.endswith("." + w) or w in h for w in WANT)]
    for r in rows[2:]:
        print("== %s  grid %s block %s" % (r[hdr.index("Kernel Name")][:90], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for i in cols:
            print("   %-100s %s %s" % (hdr[i], r[i], units[i]))


if __name__ == "__main__":
    main()
