#!/usr/bin/env python
"""bench.py — throughput of the portfolio-env hot path on B200 (driver contract in DESIGN.md §Measurement).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU step (op-for-op port) on the host cores

A "step" is one lockstep transition of every env on the rank: fused step kernel + observation
materialisation (Mode O).  Default workload = the per-GPU shard of BASELINE config 4
(1,048,576 envs x 100 assets x window 50 over 8 GPUs → 131,072 envs per GPU, weak scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (envs per GPU, assets, window, features, commission, obs materialised, description)
    "c4_shard": (131072, 100, 50, 5, 0.0, True, "BASELINE config 4 shard: 131,072 envs/GPU x 100 assets x window 50, obs materialised"),
    "c4": (1048576, 100, 50, 5, 0.0, True, "BASELINE config 4 whole: 1,048,576 envs x 100 assets x window 50 divided over the ranks (strong scaling), obs materialised"),
    "c2": (4096, 50, 50, 5, 0.0, True, "BASELINE config 2: 4,096 envs x 50 assets x window 50, obs materialised"),
    "c2_state": (4096, 50, 50, 5, 0.0, False, "BASELINE config 2, state-only step (no obs)"),
    "c3": (65536, 100, 50, 5, 0.0, True, "BASELINE config 3: 65,536 envs x 100 assets, obs materialised"),
    "c5": (262144, 500, 50, 5, 0.0025, False, "BASELINE config 5: 262,144 envs x 500 assets, commission 0.0025, state-only"),
    "c5_obs": (32768, 500, 50, 5, 0.0025, True, "wide universe with obs: 32,768 envs x 500 assets x window 50, commission 0.0025, obs materialised"),
    "c4_state": (131072, 100, 50, 5, 0.0, False, "config 4 shard, state-only step (no obs)"),
}
EPISODE_LEN = 1000
TABLE_ROWS = 4096


def algorithmic_bytes_per_asset_step(A, W, F, commission, obs):
    """Mode S = 4 (action) + 4 (w' ring write) + 4·[c>0] (w_last) + 25/A;
    Mode O adds 4·W·F (obs write) + 4·W (ring read for the weight channel, subsumes w_last)."""
    b = 4.0 + 4.0 + 25.0 / A
    if obs:
        return b + 4.0 * W * F + 4.0 * W
    return b + (4.0 if commission > 0 else 0.0)


def ncu_traffic_bytes(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture of this workload (profiles/traffic.json), or None if no capture exists for it."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline(A, W, seconds=12.0):
    """The reference's CPU step (oracle/ref_port.py, op-for-op torch port) on every host core."""
    from oracle.ref_port import time_port, time_port_all_cores
    procs = len(os.sched_getaffinity(0))
    rate1 = time_port(A, W, 500, warmup=50)                      # calibrate (≈0.1 s)
    steps = max(500, int(rate1 * seconds))
    total, rates = time_port_all_cores(A, W, steps, procs)
    # context only: the batched numpy restatement (oracle/env_oracle.py, one thread, state-only, no obs) — what a
    # vectorised CPU rewrite of the reference would reach per core
    import numpy as np
    from oracle.env_oracle import OracleEnv
    Eb = 4096
    env = OracleEnv(Eb, A, W, 5)
    rs = np.random.RandomState(0)
    acts = rs.standard_normal((4, Eb, A)).astype(np.float32)
    ys = (1 + 0.01 * rs.standard_normal((4, Eb, A))).astype(np.float32)
    env.step(acts[0], ys[0])
    t0 = time.perf_counter()
    nb = 12
    for i in range(nb):
        env.step(acts[i % 4], ys[i % 4])
    batched = nb * Eb * A / (time.perf_counter() - t0)
    try:                                                      # plain-C restatement with OpenMP over all cores (state-only)
        from oracle.c_oracle import time_all_cores
        c_port = time_all_cores(A, W)
    except Exception:
        c_port = None
    return {"value": total * A, "unit": "asset-steps/s", "cores": procs, "kind": "port",
            "env_steps_per_s": total, "batched_numpy_1core_asset_steps_per_s": batched,
            "c_port_openmp_all_cores_asset_steps_per_s": c_port,
            "sample": f"{procs} single-thread processes x {steps} steps of one env ({A} assets, window {W}), "
                      f"oracle/ref_port.py (op-for-op torch-CPU port of env/sim/trading_env.py:44-105)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    E, A, W, F, c, obs, desc = WORKLOADS[args.workload]
    t0 = time.perf_counter()
    # each "step" of this arm is a bounded sample: all cores stepping single envs for a fixed wall time
    per = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    per = float(os.environ.get("PMRL_BENCH_REF_SECONDS", per))          # tests shrink the sample
    vals = []
    for i in range(args.warmup + args.steps):
        b = cpu_baseline(A, W, seconds=per)
        if i >= args.warmup:
            vals.append(b)
    v = statistics.mean(x["value"] for x in vals)
    cores = vals[-1]["cores"]
    line = {"impl": "reference", "metric": "asset_steps_per_s", "value": v, "unit": "asset-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (time.perf_counter() - t0) / max(1, args.steps + args.warmup),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "envs_per_gpu": E, "assets": A, "window": W},
            "cpu_baseline": {"value": v, "unit": "asset-steps/s", "cores": cores, "kind": "port",
                             "sample": vals[-1]["sample"] + f"; {per:.0f} s per step"},
            "e2e": {"value": v, "unit": "asset-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4_shard", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=0,
                    help="env slices of the host-buffer step: >0 geometric x2.5, <0 equal, 0 library default (5 geometric)")
    ap.add_argument("--stats-every", type=int, default=1,
                    help="multi-GPU: all-reduce the stats vector over NCCL every K steps (SURVEY 8e: K = 1); 0 = only at the end")
    ap.add_argument("--graph", type=int, nargs="?", const=1, default=0, metavar="K",
                    help="replay the step from a CUDA graph (launch-bound small batches); K > 1 captures K-step bursts in one graph")
    ap.add_argument("--burst", type=int, default=0, metavar="K",
                    help="state-only workloads: advance K steps per launch through pmrl_env_step_burst (imagination bursts)")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VAL",
                    help="kernel launch-shape override: rows|group|ctas|fused|fast = int (pmrl_set_tuning)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import pmrl_b200
    from pmrl_b200 import dist as pdist, synth
    from pmrl_b200.env import BatchedTradingEnv

    rank, world, local_rank = pdist.init_from_env()
    from pmrl_b200 import _lib
    tune_keys = {"group": _lib.TUNE_GROUP_ENVS, "ctas": _lib.TUNE_CTAS_PER_SM, "fused": _lib.TUNE_FUSED,
                 "fast": _lib.TUNE_FAST_FILL, "rt": _lib.TUNE_RING_TMA}
    for kv in args.tune:
        k, v = kv.split("=")
        _lib.set_tuning(tune_keys[k], int(v))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    E, A, W, F, commission, obs, desc = WORKLOADS[args.workload]
    if args.envs:
        E = args.envs
    strong = args.workload == "c4" and not args.envs
    if strong:                                                # config 4 whole: the 1,048,576 envs are divided over the ranks
        E = E // world
    first_env = rank * E                                      # weak scaling: rank r owns global envs [r*E, (r+1)*E)
    tbl = synth.gbm_ohlc(TABLE_ROWS, A)
    t0 = synth.episode_offsets(E, TABLE_ROWS, W, EPISODE_LEN, first_env=first_env)
    cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, num_features=F, commission=commission,
                              episode_len=EPISODE_LEN)
    env = BatchedTradingEnv(cfg, prices=tbl, t0=t0, device=dev, collect_stats=True)

    n_pool = 4
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    pool = [torch.randn(E, A, generator=gen, device=dev) for _ in range(n_pool)]   # "random actions": raw scores
    env.reset(obs=obs)
    # steady state of a 1,000-step episode: the weight ring is full after W-1 steps (95 % of all steps), and only
    # then does the obs weight channel read the whole ring.  Pre-roll W state-only steps (untimed) to get there.
    preroll = W
    for i in range(preroll):
        env.step(pool[i % n_pool], obs=False)

    # multi-GPU: the 10-double stats vector is all-reduced over NCCL EVERY step (SURVEY.md §8(e), K = 1), issued
    # asynchronously behind a snapshot so the next step's kernel is not ordered after the collective
    reducer = pdist.AsyncStatsReducer(dev) if (world > 1 and args.stats_every > 0) else None
    stats_in_sync = None

    graph_launches_per_step = 0
    burst = max(1, args.graph)
    if args.graph:
        if args.steps % burst or args.warmup % burst:
            args.steps = max(burst, args.steps // burst * burst)
            args.warmup = max(burst, -(-args.warmup // burst) * burst)
        c0 = _lib.load().pmrl_launch_count()
        static_actions, replay = env.graphed_step(obs=obs, steps=burst)
        graph_launches_per_step = int(_lib.load().pmrl_launch_count() - c0) // (burst + 1)   # one warm-up step + the captured burst
        burst_pool = [torch.stack([pool[(j + k) % n_pool] for k in range(burst)]) if burst > 1 else pool[j] for j in range(n_pool)]

        def one_step(i):                                      # called once per env step; a burst is replayed on its first step
            if i % burst == 0:
                static_actions.copy_(burst_pool[(i // burst) % n_pool])   # the policy would write its actions here
                replay()
    elif args.burst:
        burst = args.burst
        if obs:
            raise SystemExit("--burst is a state-only mode")
        args.steps = max(burst, args.steps // burst * burst)
        args.warmup = max(burst, -(-args.warmup // burst) * burst)
        burst_pool = [torch.stack([pool[(j + k) % n_pool] for k in range(burst)]) for j in range(n_pool)]
        b_rew = torch.empty(burst, E, dtype=torch.float32, device=dev)
        b_done = torch.empty(burst, E, dtype=torch.uint8, device=dev)

        def one_step(i):
            if i % burst == 0:
                env.step_burst(burst_pool[(i // burst) % n_pool], b_rew, b_done)
    else:
        def one_step(i):
            env.step(pool[i % n_pool], obs=obs)
            if reducer is not None and i % args.stats_every == 0:
                reducer.launch(env._stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                       # samples clocks / throttle reasons over warm-up + timed region
        time.sleep(0.25)                                      # nvidia-smi needs a moment before its first sample
    for i in range(args.warmup):
        one_step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = _lib.load().pmrl_launch_count()               # kernels launched by libpmrl_b200, counted by the library itself
    ev0.record()
    for i in range(args.steps):
        one_step(args.warmup + i)
    ev1.record()
    gpu_launches = int(_lib.load().pmrl_launch_count() - launches0)
    if args.graph:
        gpu_launches = graph_launches_per_step * args.steps      # replayed from the graph: counted at capture time
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    if reducer is not None:                                   # the last step's reduced vector must equal a fresh reduce
        last = reducer.result()
        fresh = pdist.all_reduce_stats(env._stats.clone())
        stats_in_sync = bool(torch.allclose(last, fresh, rtol=1e-12, atol=0.0)) if args.stats_every == 1 else None
    stats = env.stats(all_reduce=world > 1)                   # NCCL all-reduce of the 10-double stats vector

    # ---- end to end through the public API with host buffers (H2D actions, D2H reward/done every step) ----
    e2e = None
    if not args.no_e2e:
        h_act = [p.cpu().pin_memory() for p in pool[:2]]
        h_rew = torch.empty(E, dtype=torch.float32).pin_memory()
        h_done = torch.empty(E, dtype=torch.uint8).pin_memory()

        def e2e_step(i):
            # public host-buffer API (C-ABI pmrl_env_step_host): pinned actions in, reward/done out, the caller blocks until
            # they are on the host.  Internally the batch is issued as env slices so the copies overlap the kernels.
            env.step_host(h_act[i % 2], h_rew, h_done, obs=obs, chunks=args.e2e_chunks)

        for i in range(3):
            e2e_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        e1.record()
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": world * E * A * args.steps / (float(ems.item()) * 1e-3), "unit": "asset-steps/s",
               "h2d_bytes_per_step": E * A * 4, "d2h_bytes_per_step": E * 5, "chunks": args.e2e_chunks,
               "transfer": ("zero-copy: the step kernel reads the pinned host actions over PCIe; reward/done copied D2H"
                            if args.e2e_chunks == 0 else "sliced H2D/D2H copies overlapped with per-slice kernels"),
               "api": "BatchedTradingEnv.step_host -> C-ABI pmrl_env_step_host",
               "ms_per_step": float(ems.item()) / args.steps}

    if rank == 0:
        sec = ms * 1e-3
        asset_steps = world * E * A * args.steps
        value = asset_steps / sec
        bpa = algorithmic_bytes_per_asset_step(A, W, F, commission, obs)
        peak, peak_src = measured_peak_gbs()
        per_gpu_gbs = (value / world) * bpa / 1e9
        line = {
            "metric": "asset_steps_per_s", "value": value, "unit": "asset-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "env_steps_per_s": value / A,
            "config": {"workload": args.workload, "description": desc, "envs_per_gpu": E, "envs_total": world * E,
                       "assets": A, "window": W, "features": F, "commission": commission, "obs_materialised": obs,
                       "episode_len": EPISODE_LEN, "table_rows": TABLE_ROWS, "preroll_steps": preroll, "tune": args.tune, "cuda_graph": bool(args.graph), "graph_burst_steps": burst if args.graph else None, "burst_kernel_steps": args.burst or None, "actions": "raw N(0,1) scores (softmax branch)",
                       "parallelism": f"env-shard x{world}, no data-path collective" + (", NCCL stats all-reduce every step (async)" if world > 1 else ""),
                       "l2_policy": "working set per step (obs write + ring) exceeds L2 (126 MB)" if E * A * W * 4 > 126e6
                                    else "working set smaller than L2: L2-resident by construction"},
            "roofline": {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s", "frac": per_gpu_gbs / peak,
                         "traffic": ncu_traffic_bytes(args.workload), "algorithmic_bytes_per_launch": bpa * E * A,
                         "bytes_per_asset_step": bpa, "peak_source": peak_src,
                         "kernel": ("k_env_step_obs_rt" if gpu_launches == args.steps else "k_env_step + k_obs_build") if obs else "k_env_step"},
            "clocks": clocks,
            "gpu_launches": gpu_launches,
            "stats": {k: stats[k] for k in ("n_envs", "mean_reward", "mean_value", "n_done")},
        }
        if reducer is not None:
            line["stats_allreduce"] = {"every_steps": args.stats_every, "launches": reducer.launches, "last_equals_fresh_reduce": stats_in_sync}
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(A, W)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
