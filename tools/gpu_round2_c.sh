#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/pytest.log
tail -8 $O/pytest.log
B="python bench.py --no-cpu-baseline --no-e2e"
for wl in c4_shard c5 c2 c2_state c4_state c3; do $B --workload $wl --steps 30 --warmup 6 > $O/bench_$wl.json 2> $O/bench_$wl.err; done
$B --workload c2_state --burst 15 --steps 150 --warmup 30 > $O/bench_c2_state_burst15.json 2>&1
$B --workload c5 --burst 15 --steps 30 --warmup 15 > $O/bench_c5_burst15.json 2>&1
$B --workload c4_state --burst 15 --steps 30 --warmup 15 > $O/bench_c4_state_burst15.json 2>&1
$B --workload c2_state --graph 15 --steps 150 --warmup 30 > $O/bench_c2_state_graph15.json 2>&1
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('bench_')[1], 'ms/step %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
$B --workload c5 --steps 10 --warmup 5 > $O/plain_c5.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name 'regex:^k_env_step$' --launch-skip 58 --launch-count 1 -f -o $O/c5 $B --workload c5 --steps 10 --warmup 5 > $O/ncu_c5.log 2>&1
$B --workload c5 --burst 15 --steps 15 --warmup 15 > $O/plain_c5b.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_step_burst --launch-skip 1 --launch-count 1 -f -o $O/c5_burst $B --workload c5 --burst 15 --steps 15 --warmup 15 > $O/ncu_c5b.log 2>&1
$B --workload c2_state --burst 15 --steps 30 --warmup 15 > $O/plain_c2b.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_step_burst --launch-skip 2 --launch-count 1 -f -o $O/c2_burst $B --workload c2_state --burst 15 --steps 30 --warmup 15 > $O/ncu_c2b.log 2>&1
ls $O
