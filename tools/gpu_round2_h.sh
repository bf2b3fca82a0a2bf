#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-e2e --no-configs"
$B --workload c2_state --burst 15 --steps 30 --warmup 15 > $O/plain_c2b.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_step_burst --launch-skip 2 --launch-count 1 -f -o $O/c2_burst $B --workload c2_state --burst 15 --steps 30 --warmup 15 > $O/ncu_c2b.log 2>&1
python tools/ncu_summary.py $O/c2_burst.ncu-rep $O/c2_burst.ncu_summary.csv; python tools/ncu_hot.py $O/c2_burst.ncu-rep 45 > $O/c2_burst.hot.txt 2>&1
python tools/ncu_opcodes.py $O/c2_burst.ncu-rep $O/c2_burst.sass_exec.txt | head -30
$B --workload c2 --steps 10 --warmup 5 > $O/plain_c2.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_step_obs_rt --launch-skip 8 --launch-count 1 -f -o $O/c2 $B --workload c2 --steps 10 --warmup 5 > $O/ncu_c2.log 2>&1
python tools/ncu_summary.py $O/c2.ncu-rep $O/c2.ncu_summary.csv; python tools/ncu_hot.py $O/c2.ncu-rep 45 > $O/c2.hot.txt 2>&1
rm -f $O/*.ncu-rep
