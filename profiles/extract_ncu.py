"""How the `*_ncu_full_*.csv` summaries in this directory were made from the `.ncu-rep` captures (scratch files, not committed):

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c <n> -f -o gpurun_out/prof python bench.py ...
    python profiles/extract_ncu.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] > profiles/<name>.csv

One row per captured launch with the metrics the roofline discussion uses (duration, DRAM bytes, utilisation, issue activity,
FP64 pipe, instruction count, cache hit rates, the main stall reasons)."""
import csv
import subprocess
import sys

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,"
           "launch__registers_per_thread,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,"
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,"
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,"
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,"
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,"
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio")


def main(paths):
    out = csv.writer(sys.stdout)
    first = True
    for path in paths:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--metrics", METRICS], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if not rows:
            continue
        hdr = rows[0]
        keep = [i for i, h in enumerate(hdr) if h == "Kernel Name" or h in METRICS.split(",")]
        for r in rows[(0 if first else 2):]:             # header + units once
            out.writerow([r[i] for i in keep])
        first = False


if __name__ == "__main__":
    main(sys.argv[1:])
