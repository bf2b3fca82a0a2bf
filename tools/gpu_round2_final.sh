#!/bin/bash
# round-2 measurement batch: parity suite, smoke, the default bench line (N = 1), the reference arm, the ncu launch list of the
# same command, and one ncu --set full capture per dominant kernel (summaries only; the .ncu-rep files stay on the box)
O=gpurun_out/r2final; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest.log
tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -4 $O/smoke.log
python bench.py --steps 20 --warmup 5 > $O/bench_default_1gpu.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
PMRL_BENCH_REF_SECONDS=4 python bench.py --impl reference --steps 2 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference.err
python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > $O/plain_launchlist.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c4_shard.csv python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > $O/ncu_launchlist.log 2>&1
python tools/launch_summary.py $O/launches_bench_c4_shard.csv "bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline (c4_shard; pre-roll + warm-up + timed + e2e steps)" > $O/launches_bench_c4_shard.summary.txt 2>&1
B="python bench.py --no-cpu-baseline --no-e2e --no-configs"
cap() {  # name workload kernel-regex skip extra-args
  $B --workload $2 --steps 6 --warmup 3 $5 > $O/plain_$1.json 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name "regex:$3" --launch-skip $4 --launch-count 1 -f -o $O/$1 $B --workload $2 --steps 6 --warmup 3 $5 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/$1.ncu-rep $O/$1.ncu_summary.csv; python tools/ncu_hot.py $O/$1.ncu-rep 40 > $O/$1.hot.txt 2>&1
  python tools/ncu_opcodes.py $O/$1.ncu-rep > $O/$1.opcodes.txt 2>&1
  rm -f $O/$1.ncu-rep
}
cap a_k_env_step_obs_rt_c4_shard c4_shard k_env_step_obs_rt 6 ""
cap b_k_env_step_obs_rt_c4_f9 c4_f9 k_env_step_obs_rt 6 ""
cap c_k_env_step_staged_c5 c5 k_env_step_staged 56 ""
cap d_k_env_step_obs_rt_c2 c2 k_env_step_obs_rt 6 ""
cap e_k_env_step_burst_c2_state c2_state k_env_step_burst 1 "--burst 15 --steps 30 --warmup 15"
cap f_k_env_step_obs_rt_c3 c3 k_env_step_obs_rt 6 ""
cap g_k_env_step_state_c4_state c4_state "^k_env_step$" 56 ""
ls $O
