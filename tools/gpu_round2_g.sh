#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
python -m pytest tests/test_env_gpu.py tests/test_features_buffers_gpu.py -m gpu -q -x -k "table_driven or indicators" 2>&1 | tail -5 > $O/pytest.log
tail -3 $O/pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-configs"
for wl in c4_f9 c4_shard; do
  for i in 1 2 3; do $B --workload $wl --steps 30 --warmup 6 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"; done
done
