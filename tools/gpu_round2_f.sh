#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-e2e --no-configs"
$B --workload c5 --steps 10 --warmup 5 > $O/plain_c5.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_env_step_staged' --launch-skip 58 --launch-count 1 -f -o $O/c5 $B --workload c5 --steps 10 --warmup 5 > $O/ncu_c5.log 2>&1
python tools/ncu_opcodes.py $O/c5.ncu-rep $O/c5.sass_exec.txt
python tools/ncu_summary.py $O/c5.ncu-rep $O/c5.ncu_summary.csv
ncu -i $O/c5.ncu-rep --page details --csv > $O/c5.details.csv 2>/dev/null
rm -f $O/*.ncu-rep
