#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_features_buffers_gpu.py tests/test_step_io_gpu.py -m gpu -q -x -k "table_driven or indicators or full_size or thousand_step or fused or graph or two_envs or midsize or random_shape" 2>&1 | tail -5 > $O/pytest.log
tail -3 $O/pytest.log
B="timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-configs"
for wl in c4_shard c4_f9 c3 c2; do
  for i in 1 2 3; do $B --workload $wl --steps 30 --warmup 6 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"; done
done
