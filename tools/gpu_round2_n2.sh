#!/bin/bash
O=gpurun_out/r2n2; mkdir -p $O
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2.json 2> $O/bench_n2.err
tail -5 $O/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n2/bench_n2.json').read().strip().splitlines()[-1])
print('headline', d['n_gpus'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d.get('stats_allreduce'))
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k, 'ERROR', v['error']); continue
    if 'ms_per_step' in v: print(k, v['envs_per_gpu'], 'ms %.4f'%v['ms_per_step'], 'frac %.3f'%v['roofline']['frac'], 'launches', v['gpu_launches'], 'e2e', (v.get('e2e') or {}).get('ms_per_step'))
    else: print(k, json.dumps(v)[:300])
PY
python -m pytest tests -m gpu -q -x -k "burst or step_io or staged" 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-e2e --no-configs --workload c2_state --burst 15 --steps 150 --warmup 30 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 burst', d['ms_per_step'])"
