"""Tiny run of every env kernel variant for compute-sanitizer (memcheck / racecheck / synccheck): reset, fused RT kernel at
F = 5 / 9 / 13, register-staged fused kernel, two-kernel paths (division-free and generic tile fill), state-only step (register
loads and TMA-staged rows, both staging depths), K-step burst, step_io sinks, streamed host step, feature path, buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmrl_b200
from pmrl_b200 import _lib, synth, features
from pmrl_b200.buffers import DeviceRolloutBuffer
from pmrl_b200.env import BatchedTradingEnv

DEFAULTS = {_lib.TUNE_FUSED: 1, _lib.TUNE_FAST_FILL: 1, _lib.TUNE_RING_TMA: 1, _lib.TUNE_STAGED: 1, _lib.TUNE_HOST_STREAM: 0,
            _lib.TUNE_GROUP_ENVS: 0, _lib.TUNE_CTAS_PER_SM: 0}


def run(A, W, E, tune, F=5, commission=0.0025, host=False, burst=False, sinks=False):
    for k, v in tune.items():
        _lib.set_tuning(k, v)
    T, L = 96, 12
    tbl = synth.gbm_ohlc(T, A)
    feats = None if F == 5 else torch.cat([tbl, torch.rand(T, A, F - 5)], dim=-1)
    cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, num_features=F, episode_len=L, commission=commission)
    env = BatchedTradingEnv(cfg, prices=tbl, features=feats, t0=synth.episode_offsets(E, T, W, L), collect_stats=True)
    env.reset()
    buf = DeviceRolloutBuffer(F, L + W + 4, E, A, W, mode="index", feat_am=env.feat_am, y_tm=env.y_tm) if sinks else None
    for s in range(L + 3):
        a = torch.randn(E, A, device="cuda")
        if host:
            env.step_host(a.cpu().pin_memory(), obs=(s % 3 != 2))
        elif sinks:
            env.step_io(a, **buf.sinks(W + s)); buf.advance()
        else:
            env.step(a, obs=(s % 3 != 2))
    if burst:
        env.step_burst(torch.randn(5, E, A, device="cuda"))
    torch.cuda.synchronize()
    for k, v in DEFAULTS.items():
        _lib.set_tuning(k, v)


run(100, 50, 19, {})                                          # fused RT kernel, F = 5
run(100, 50, 19, {}, F=9)                                     # ... F = 9 (16-row tiles)
run(52, 20, 21, {}, F=13)                                     # ... F = 13 (8-row tiles, narrow envs: two per warp)
run(100, 50, 19, {_lib.TUNE_RING_TMA: 0})                     # register-staged fused kernel
run(33, 9, 21, {_lib.TUNE_FAST_FILL: 0}, F=7)                 # state-only step + generic tile kernel (incremental indices)
run(33, 9, 21, {_lib.TUNE_FUSED: 0})                          # two kernels, division-free tile fill
run(500, 8, 40, {_lib.TUNE_STAGED: 2})                        # TMA-staged wide-env step, double-buffered rows
run(500, 8, 40, {_lib.TUNE_STAGED: 2, _lib.TUNE_CTAS_PER_SM: 3})   # ... single-stage rows, previous weights re-read from shared memory
run(260, 8, 30, {_lib.TUNE_STAGED: 2}, commission=0.0)
run(500, 8, 40, {_lib.TUNE_STAGED: 0}, burst=True)            # register-load step + burst kernel (wide)
run(50, 12, 64, {}, burst=True)                               # burst kernel (narrow, register prefetch)
run(100, 50, 19, {}, sinks=True)                              # step_io sinks into an index-mode rollout buffer
run(100, 16, 300, {_lib.TUNE_HOST_STREAM: 2}, host=True)      # streamed host step through the fused kernel
run(500, 8, 300, {_lib.TUNE_HOST_STREAM: 2, _lib.TUNE_STAGED: 2}, host=True)   # ... through the staged kernel + tile kernel
out, widths, mw = features.ffd_transform(torch.rand(6, 700).cuda() + 1, [0.4] * 6, 1e-3)
features.scale_series(out)
t = features.build_env_tables(synth.gbm_ohlc(400, 6), d=0.5, thres=1e-3, indicators=[("ema", {"timeperiod": 10}), ("bbands", {"timeperiod": 5})])
torch.cuda.synchronize()
print("sanitize run ok")
