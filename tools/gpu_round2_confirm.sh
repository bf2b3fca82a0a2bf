#!/bin/bash
# confirmation pass after the last code change: parity suite, smoke, the default bench line, the reference arm
O=gpurun_out/r2confirm; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest.log
tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -4 $O/smoke.log
python bench.py --steps 20 --warmup 5 > $O/bench_default_1gpu.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
PMRL_BENCH_REF_SECONDS=4 python bench.py --impl reference --steps 2 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2confirm/bench_default_1gpu.json').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], '%.3e'%d['e2e']['value'], d['clocks'])
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k, 'ERROR', v['error']); continue
    if 'ms_per_step' in v: print(k, 'ms %.4f'%v['ms_per_step'], 'frac %.3f'%v['roofline']['frac'], 'e2e', (v.get('e2e') or {}).get('ms_per_step'))
PY
