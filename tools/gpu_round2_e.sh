#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
python -m pytest tests/test_env_gpu.py -m gpu -q -x -k "staged" 2>&1 | tail -8 > $O/pytest.log
tail -4 $O/pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-configs"
for t in "ctas=0" "ctas=6" "ctas=7"; do
  for i in 1 2; do $B --workload c5 --steps 30 --warmup 6 --tune $t | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 $t', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"; done
done
for t in 0 7; do
$B --workload c5 --steps 10 --warmup 5 --tune ctas=$t > $O/plain_c5.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_env_step_staged' --launch-skip 58 --launch-count 1 -f -o $O/c5_$t $B --workload c5 --steps 10 --warmup 5 --tune ctas=$t > $O/ncu_c5.log 2>&1
python tools/ncu_summary.py $O/c5_$t.ncu-rep $O/c5_$t.ncu_summary.csv; python tools/ncu_opcodes.py $O/c5_$t.ncu-rep $O/c5_$t.sass_exec.txt | head -12
done
rm -f $O/*.ncu-rep
