#!/bin/bash
# round-2 first GPU pass: parity suite, smoke, and the workloads whose kernels changed
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a/smoke.log 2>&1
for wl in c4_shard c5 c2 c2_state c4_state; do
  python bench.py --workload $wl --steps 30 --warmup 6 --no-cpu-baseline > gpurun_out/r2a/bench_$wl.json 2> gpurun_out/r2a/bench_$wl.err
done
python bench.py --workload c2_state --burst 15 --steps 150 --warmup 30 --no-cpu-baseline --no-e2e > gpurun_out/r2a/bench_c2_state_burst15.json 2> gpurun_out/r2a/bench_c2_state_burst15.err
python bench.py --workload c5 --burst 15 --steps 30 --warmup 15 --no-cpu-baseline --no-e2e > gpurun_out/r2a/bench_c5_burst15.json 2> gpurun_out/r2a/bench_c5_burst15.err
python bench.py --workload c2_state --graph 15 --steps 150 --warmup 30 --no-cpu-baseline --no-e2e > gpurun_out/r2a/bench_c2_state_graph15.json 2> gpurun_out/r2a/bench_c2_state_graph15.err
tail -3 gpurun_out/r2a/pytest.log; cat gpurun_out/r2a/smoke.log | tail -5
for f in gpurun_out/r2a/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('bench_')[1], 'ms/step %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e', d.get('e2e',{}).get('value'), 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
