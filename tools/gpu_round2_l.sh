#!/bin/bash
O=gpurun_out/r2final; mkdir -p $O
timeout 300 python -m pytest tests/test_env_gpu.py -m gpu -q -x -k "step_host" 2>&1 | tail -2
python tools/e2e_ab.py
python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > $O/plain_launchlist.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c4_shard.csv python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > $O/ncu_launchlist.log 2>&1
python tools/launch_summary.py $O/launches_bench_c4_shard.csv "bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline (c4_shard; pre-roll + warm-up + timed + e2e steps)" > $O/launches_bench_c4_shard.summary.txt 2>&1
head -5 $O/launches_bench_c4_shard.summary.txt | cut -c1-160
