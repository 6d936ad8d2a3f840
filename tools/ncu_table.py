#!/usr/bin/env python
"""`ncu -i x.ncu-rep --page raw --csv` export -> one line per launch (the table format of profiles/rNN_ncu_full_eval_*.txt).
usage: ncu_table.py raw.csv "header comment" > out.txt"""
import csv
import sys

COLS = [("us", "gpu__time_duration.sum"), ("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("MHz", "smsp__cycles_elapsed.avg.per_second"),
        ("grid", "launch__grid_size"), ("blk", "launch__block_size"), ("regs", "launch__registers_per_thread"),
        ("dsmem", "launch__shared_mem_per_block_dynamic")]


def main(path, comment):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(n, m) for n, m in COLS if m in idx]
    print(f"# {comment}")
    print("# cold-cache, serialised per-kernel replays: absolute times are not bench values; units: " +
          ", ".join(f"{n}[{units[idx[m]]}]" for n, m in cols))
    print("idx | kernel | " + " | ".join(n for n, _ in cols))
    for i, d in enumerate(data):
        name = d[idx["Kernel Name"]].split("(")[0][:40]
        vals = []
        for n, m in cols:
            v = d[idx[m]].replace(",", "")
            try:
                f = float(v)
                vals.append(f"{f:.1f}" if (f != int(f) or n in ("us", "rd", "wr")) else str(int(f)))
            except ValueError:
                vals.append(v)
        print(f"{i:3d} | {name} | " + " | ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
