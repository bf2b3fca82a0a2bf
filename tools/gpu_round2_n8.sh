#!/bin/bash
N=${1:-8}
O=gpurun_out/r2n$N; mkdir -p $O
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -4 $O/bench_n$N.err
python - $N <<'PY'
import json, sys
N=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2n{N}/bench_n{N}.json').read().strip().splitlines()[-1])
print('headline', d['n_gpus'], d['ms_per_step'], '%.3e'%d['value'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], '%.3e'%d['e2e']['value'], d.get('stats_allreduce'))
for k,v in d.get('configs',{}).items():
    if 'error' in v: print(k, 'ERROR', v['error']); continue
    if 'ms_per_step' in v: print(k, v['envs_per_gpu'], 'ms %.4f'%v['ms_per_step'], 'val %.3e'%v['value'], 'frac %.3f'%v['roofline']['frac'], 'e2e', (v.get('e2e') or {}).get('ms_per_step'))
PY
nvidia-smi topo -m 2>/dev/null | head -14 > $O/topo.txt
