#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-e2e --no-configs"
cap() {  # name workload kernel-regex skip extra-args
  $B --workload $2 --steps 6 --warmup 3 $5 > $O/plain_$1.json 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name "regex:$3" --launch-skip $4 --launch-count 1 -f -o $O/$1 $B --workload $2 --steps 6 --warmup 3 $5 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/$1.ncu-rep $O/$1.ncu_summary.csv; python tools/ncu_hot.py $O/$1.ncu-rep 40 > $O/$1.hot.txt 2>&1
  rm -f $O/$1.ncu-rep
}
cap f9 c4_f9 k_env_step_obs_rt 6 ""
cap c4 c4_shard k_env_step_obs_rt 6 ""
head -24 $O/f9.hot.txt
head -24 $O/c4.hot.txt
