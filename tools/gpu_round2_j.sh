#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
timeout 600 python -m pytest tests/test_env_gpu.py -m gpu -q -x -k "step_host" 2>&1 | tail -6 > $O/pytest.log
tail -4 $O/pytest.log
B="timeout 300 python bench.py --no-cpu-baseline --no-configs"
for wl in c4_shard c3 c5 c4_state c2; do
 for t in 0 1; do
  $B --workload $wl --steps 20 --warmup 5 --tune hoststream=$t | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl hoststream=$t', 'dev', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
 done
done
