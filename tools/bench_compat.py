"""Latency of the drop-in E = 1 `TradingEnv` (BASELINE config 1 shape: 11 assets, window 50) driven exactly like
train/on_policy.py:59-67 drives the reference env — one step at a time, features / prices supplied by the caller."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmrl_b200
from pmrl_b200.compat import TradingEnv


def run(device, steps=2000):
    A, W, F = 11, 50, 5
    env = TradingEnv(pmrl_b200.EnvConfig(num_assets=A, window_size=W, num_features=F))
    g = torch.Generator().manual_seed(0)
    feat = torch.rand(A, W, F, generator=g).to(device)
    acts = torch.randn(64, 1, A, 1, generator=g).to(device)
    ys = (1 + 0.01 * torch.randn(64, A, generator=g)).to(device)
    env.reset(feat)
    for s in range(50):
        env.step(acts[s % 64], feat, ys[s % 64])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps):
        if s % 500 == 499:
            env.reset(feat)
        r, _ = env.step(acts[s % 64], feat, ys[s % 64])
    float(r)                                                    # the loop logs the reward every step in the reference
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return steps / dt


if __name__ == "__main__":
    for dev in ("cpu", "cuda"):
        print(json.dumps({"bench": "compat.TradingEnv E=1, 11 assets, window 50", "caller_tensors": dev,
                          "env_steps_per_s": run(dev), "note": "reference CPU env: ~7,000 steps/s per core (BASELINE.md §2)"}))
