"""Where the end-to-end step's time goes (config-4 shard): kernel only, + per-step host sync, + D2H, + H2D, full host step."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmrl_b200
from pmrl_b200 import synth
from pmrl_b200.env import BatchedTradingEnv


def main():
    E, A, W = (int(sys.argv[1]) if len(sys.argv) > 1 else 131072), (int(sys.argv[2]) if len(sys.argv) > 2 else 100), 50
    tbl = synth.gbm_ohlc(4096, A)
    cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=1000)
    env = BatchedTradingEnv(cfg, prices=tbl, t0=synth.episode_offsets(E, 4096, W, 1000), device="cuda", collect_stats=True)
    env.reset()
    acts = [torch.randn(E, A, device="cuda") for _ in range(2)]
    h_act = [a.cpu().pin_memory() for a in acts]
    h_r = torch.empty(E).pin_memory(); h_d = torch.empty(E, dtype=torch.uint8).pin_memory()
    stage = torch.empty(E, A, device="cuda")
    for i in range(W):
        env.step(acts[i % 2], obs=False)
    n = 30

    def wall(fn, label):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for i in range(n):
            fn(i)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t) / n * 1e3
        print(json.dumps({"case": label, "ms_per_step": round(ms, 4)}), flush=True)

    wall(lambda i: env.step(acts[i % 2]), "kernel only, no per-step sync")

    def k_sync(i):
        env.step(acts[i % 2]); torch.cuda.synchronize()
    wall(k_sync, "kernel + host sync every step")

    def k_d2h(i):
        env.step(acts[i % 2]); h_r.copy_(env.reward, non_blocking=True); h_d.copy_(env.done, non_blocking=True); torch.cuda.synchronize()
    wall(k_d2h, "kernel + D2H reward/done + sync")

    def h2d_only(i):
        stage.copy_(h_act[i % 2], non_blocking=True); torch.cuda.synchronize()
    wall(h2d_only, "H2D of the actions alone (52 MB pinned)")

    def serial(i):
        stage.copy_(h_act[i % 2], non_blocking=True); env.step(stage)
        h_r.copy_(env.reward, non_blocking=True); h_d.copy_(env.done, non_blocking=True); torch.cuda.synchronize()
    wall(serial, "H2D + kernel + D2H, serial")
    from pmrl_b200 import _lib
    from pmrl_b200.env import OBS_FULL
    buf = env._obs_buffer(None)

    def zero_copy(i, out_host):
        # the kernel reads the pinned host actions over PCIe itself (UVA: the pinned pointer is device-accessible)
        rc = env.lib.pmrl_env_step(env._p_cfg, env._p_tbl, env._p_st, h_act[i % 2].data_ptr(), None,
                                   h_r.data_ptr() if out_host else env.reward.data_ptr(),
                                   h_d.data_ptr() if out_host else env.done.data_ptr(), buf.data_ptr(), OBS_FULL,
                                   _lib.ptr(env._stats), _lib.current_stream())
        _lib.check(rc, "step")
        if not out_host:
            h_r.copy_(env.reward, non_blocking=True); h_d.copy_(env.done, non_blocking=True)
        torch.cuda.synchronize()
    wall(lambda i: zero_copy(i, False), "zero-copy actions (kernel reads pinned host memory) + D2H copies")
    wall(lambda i: zero_copy(i, True), "zero-copy actions + kernel writes reward/done to pinned host memory")
    wall(lambda i: zero_copy(i, False), "(repeat) zero-copy actions + D2H copies")
    wall(lambda i: zero_copy(i, True), "(repeat) zero-copy actions + host-mapped reward/done")
    wall(lambda i: env.step_host(h_act[i % 2], h_r, h_d, chunks=0), "pmrl_env_step_host slices=0 (zero-copy default)")
    for c in (-1, 3, 6, 0):
        wall(lambda i: env.step_host(h_act[i % 2], h_r, h_d, chunks=c), f"pmrl_env_step_host slices={c}")


if __name__ == "__main__":
    main()
