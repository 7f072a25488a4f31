"""`ncu -i <report>.ncu-rep --page raw --csv` of an `ncu --set full` capture -> the few columns the design discussion uses
(duration, DRAM bytes and throughput, registers, achieved occupancy, L1/L2 hit rates, issue utilisation, occupancy limiters).
Usage: ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > /tmp/raw.csv; python tools/summarize_ncu_full.py /tmp/raw.csv profiles/r2_ncu_full.csv"""
import csv
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def main(src, dst):
    rows = list(csv.reader(open(src, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    head, units, body = rows[hi], rows[hi + 1], rows[hi + 2:]
    idx = [head.index(c) for c in COLS if c in head]
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in body:
            if len(r) == len(head):
                w.writerow([r[i] for i in idx])
    print(len(body), "launches ->", dst)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
