set -x
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 3 2>/dev/null | tail -1 > $O/bench_reference_arm.json
timeout -s KILL 600 python bench.py 2>/dev/null | tail -1 > $O/bench_c4_shard_1gpu.json
for w in c2 c2_state c3 c5 c4_state c5_obs; do timeout -s KILL 300 python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_${w}_1gpu.json; done
timeout -s KILL 300 python tools/e2e_breakdown.py 2>/dev/null > $O/e2e_breakdown.jsonl
timeout -s KILL 300 python tools/bench_feature_path.py 2>/dev/null > $O/feature_path.jsonl
timeout -s KILL 300 python tools/bench_compat.py 2>/dev/null > $O/compat.jsonl
# launch list of the default bench command (after it ran plain above)
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
# full captures: RT kernel (c4 shard), state-only wide (c5), obs tile kernel (c5_obs)
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_step_obs_rt --launch-skip 4 --launch-count 1 -f -o $O/rt python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_rt.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_env_step --launch-skip 58 --launch-count 1 -f -o $O/c5 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --workload c5 > $O/ncu_c5.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_obs_build_rows --launch-skip 4 --launch-count 1 -f -o $O/obsrows python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --workload c5_obs > $O/ncu_obsrows.log 2>&1
ls -la $O
