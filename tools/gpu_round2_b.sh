#!/bin/bash
# round-2 second GPU pass: whole parity suite (no -x), then ncu --set full of the kernels that changed
O=gpurun_out/r2b; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest.log
tail -5 $O/pytest.log
B="python bench.py --no-cpu-baseline --no-e2e"
$B --workload c5 --steps 10 --warmup 5 > $O/plain_c5.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_stepI --launch-skip 58 --launch-count 1 -f -o $O/c5 $B --workload c5 --steps 10 --warmup 5 > $O/ncu_c5.log 2>&1
$B --workload c5 --burst 15 --steps 15 --warmup 15 > $O/plain_c5b.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_step_burst --launch-skip 1 --launch-count 1 -f -o $O/c5_burst $B --workload c5 --burst 15 --steps 15 --warmup 15 > $O/ncu_c5b.log 2>&1
$B --workload c2_state --burst 15 --steps 30 --warmup 15 > $O/plain_c2b.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_step_burst --launch-skip 2 --launch-count 1 -f -o $O/c2_burst $B --workload c2_state --burst 15 --steps 30 --warmup 15 > $O/ncu_c2b.log 2>&1
$B --workload c4_shard --steps 5 --warmup 3 > $O/plain_c4.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_step_obs_rt --launch-skip 4 --launch-count 1 -f -o $O/rt $B --workload c4_shard --steps 5 --warmup 3 > $O/ncu_rt.log 2>&1
$B --workload c4_state --steps 10 --warmup 5 > $O/plain_c4s.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_stepI --launch-skip 58 --launch-count 1 -f -o $O/c4s $B --workload c4_state --steps 10 --warmup 5 > $O/ncu_c4s.log 2>&1
for i in 1 2 3; do $B --workload c4_shard --steps 30 --warmup 6 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4_shard', d['ms_per_step'], d['roofline']['frac'])"; done
ls -la $O
