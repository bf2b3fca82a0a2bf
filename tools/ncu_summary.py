"""ncu report → the small CSV kept under profiles/ (launch,kernel,metric,value,unit).  usage: ncu_summary.py <rep> <out.csv>"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["launch", "kernel", "metric", "value", "unit"])
        for n, row in enumerate(rows[2:]):
            for i, h in enumerate(hdr):
                stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(row[i] or 0) > 0.2
                if h in WANT or stall:
                    w.writerow([n, row[ki], h, row[i], units[i]])


if __name__ == "__main__":
    main()
