"""A/B of the host-buffer step on ONE box: zero-copy reads vs copy-engine streaming (PMRL_TUNE_HOST_STREAM), alternating,
several repetitions, per-step wall times (host clock around step_host: the call blocks until reward/done are on the host)."""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmrl_b200
from pmrl_b200 import _lib, synth
from pmrl_b200.env import BatchedTradingEnv

def main():
    for name, E, A, W, obs in (("c4_shard", 131072, 100, 50, True), ("c3", 65536, 100, 50, True), ("c2", 4096, 50, 50, True)):
        tbl = synth.gbm_ohlc(4096, A)
        cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=1000)
        env = BatchedTradingEnv(cfg, prices=tbl, t0=synth.episode_offsets(E, 4096, W, 1000), collect_stats=True)
        env.reset()
        acts = [torch.randn(E, A).pin_memory() for _ in range(2)]
        d_acts = [a.cuda() for a in acts]
        h_r = torch.empty(E).pin_memory(); h_d = torch.empty(E, dtype=torch.uint8).pin_memory()
        for i in range(W):
            env.step(d_acts[i % 2], obs=False)
        torch.cuda.synchronize()
        # device-only reference on this box
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(5): env.step(d_acts[i % 2], obs=obs)
        ev0.record()
        for i in range(20): env.step(d_acts[i % 2], obs=obs)
        ev1.record(); torch.cuda.synchronize()
        dev_ms = ev0.elapsed_time(ev1) / 20
        res = {"workload": name, "device_ms": dev_ms}
        for rep in range(3):
            for mode in (0, 2):
                for mirror in (0, 1):
                    _lib.set_tuning(_lib.TUNE_HOST_STREAM, mode)
                    _lib.set_tuning(_lib.TUNE_HOST_MIRROR, mirror)
                    for i in range(4): env.step_host(acts[i % 2], h_r, h_d, obs=obs)
                    ts = []
                    for i in range(20):
                        t0 = time.perf_counter(); env.step_host(acts[i % 2], h_r, h_d, obs=obs); ts.append((time.perf_counter() - t0) * 1e3)
                    res.setdefault(("stream" if mode else "zerocopy") + ("+mirror" if mirror else "+d2h"), []).append(round(statistics.median(ts), 4))
        _lib.set_tuning(_lib.TUNE_HOST_STREAM, 0); _lib.set_tuning(_lib.TUNE_HOST_MIRROR, 1)
        print(json.dumps(res), flush=True)
        del env

if __name__ == "__main__":
    main()
