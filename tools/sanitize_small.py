"""Tiny run of every env kernel variant for compute-sanitizer (memcheck): reset, fused RT / fast / TM / TMA-pipeline /
generic / two-kernel paths, state-only step, feature path, buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmrl_b200
from pmrl_b200 import _lib, synth, features
from pmrl_b200.env import BatchedTradingEnv

def run(A, W, E, tune):
    for k, v in tune.items():
        _lib.set_tuning(k, v)
    T, L = 96, 12
    tbl = synth.gbm_ohlc(T, A)
    cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=L, commission=0.0025)
    env = BatchedTradingEnv(cfg, prices=tbl, t0=synth.episode_offsets(E, T, W, L), collect_stats=True)
    env.reset()
    for s in range(L + 3):
        env.step(torch.randn(E, A, device="cuda"), obs=(s % 3 != 2))
    torch.cuda.synchronize()
    for k in tune:
        _lib.set_tuning(k, 1 if k in (_lib.TUNE_FUSED, _lib.TUNE_FAST_FILL, _lib.TUNE_RING_TMA) else 0)

run(100, 50, 19, {})                                          # RT
run(100, 50, 19, {_lib.TUNE_RING_TMA: 0})                     # register-ring fast kernel
run(100, 50, 19, {_lib.TUNE_TENSORMAP: 1})                    # tensor-map TM + steppers
run(100, 50, 19, {_lib.TUNE_TMA_PIPELINE: 1})                 # first TMA pipeline
run(33, 9, 21, {_lib.TUNE_FAST_FILL: 0})                      # generic fused
run(33, 9, 21, {_lib.TUNE_FUSED: 0})                          # two kernels
out, widths, mw = features.ffd_transform(torch.rand(6, 700).cuda() + 1, [0.4] * 6, 1e-3)
features.scale_series(out)
torch.cuda.synchronize()
print("sanitize run ok")
