#!/bin/bash
# round-2 pass d: parity suite, the full default bench line (with configs), ncu of the c5 state-only kernel (summaries only)
O=gpurun_out/r2d; mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/pytest.log
tail -4 $O/pytest.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
tail -3 $O/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d/bench_default.json').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k, 'ERROR', v['error']); continue
    if 'ms_per_step' in v: print(k, 'ms %.4f'%v['ms_per_step'], 'frac %.3f'%v['roofline']['frac'], 'launches', v['gpu_launches'], 'e2e', (v.get('e2e') or {}).get('ms_per_step'), 'flushed', (v.get('l2_flushed') or {}).get('ms_per_step'))
    else: print(k, json.dumps(v)[:600])
PY
B="python bench.py --no-cpu-baseline --no-e2e --no-configs"
$B --workload c5 --steps 10 --warmup 5 > $O/plain_c5.json 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name 'regex:^k_env_step$' --launch-skip 58 --launch-count 1 -f -o $O/c5 $B --workload c5 --steps 10 --warmup 5 > $O/ncu_c5.log 2>&1
python tools/ncu_summary.py $O/c5.ncu-rep $O/c5.ncu_summary.csv; python tools/ncu_hot.py $O/c5.ncu-rep 60 > $O/c5.hot.txt 2>&1
rm -f $O/*.ncu-rep
ls -la $O
