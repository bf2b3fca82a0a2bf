"""Summarise an ncu report: key metrics + the hottest SASS lines by stall samples.  usage: ncu_hot.py <rep> [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units, row = r[0], r[1], r[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for i, h in enumerate(hdr):
    if h in want or 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(row[i] or 0) > 0.3:
        print(f"{h:95s} {row[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = None; data = []; started = False
for x in rows:
    if x and x[0] == "Kernel Name":
        if started: break
        started = True; continue
    if x and x[0] == "Address": h = x; continue
    if h: data.append(dict(zip(h, x)))
tot = sum(int(d["# Samples"] or 0) for d in data)
print("total samples", tot, "instructions", len(data))
for d in sorted(data, key=lambda d: -int(d["# Samples"] or 0))[:N]:
    print(f'{int(d["# Samples"]):8d} {100*int(d["# Samples"])/tot:5.1f}%  exec={d["Instructions Executed"]:>10s}  {d["Source"][:110]}')
