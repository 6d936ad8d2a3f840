#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the few metrics quoted in profiles/ and DESIGN.md."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "smsp__inst_executed.sum", "launch__occupancy_limit_registers",
        "sm__inst_executed_pipe_xu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"]


def main(path):
    if path.endswith(".csv"):  # already exported with `ncu -i x.ncu-rep --page raw --csv`
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel:", [d[idx["Kernel Name"]][:60] for d in data])
    for w in WANT:
        if w in idx:
            print(f"{w} [{units[idx[w]]}]:", [d[idx[w]] for d in data])


if __name__ == "__main__":
    main(sys.argv[1])
