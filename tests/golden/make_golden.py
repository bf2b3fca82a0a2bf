"""Generate the committed golden fixtures by executing the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py

Produces tests/golden/env_*.npz (TradingEnv rollouts), rollout_buffer.npz (RolloutBuffer add /
sample_random), pg_reward.npz (PG._reward fwd + autograd gradient).  The FFD fixture comes from
make_golden_ffd.py (needs import stubs).  Inputs are seeded numpy RandomState streams, stored in the
fixture (small cases) or reproduced from the seed with a sha256 check (large cases).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import live_reference as live  # noqa: E402

f32 = np.float32


def make_inputs(seed: int, kind: str, S: int, A: int):
    """Seeded action / price-relative streams (also imported by the tests to regenerate large cases)."""
    rs = np.random.RandomState(seed)
    z = rs.standard_normal((S, A))
    sig = 0.01 * (1.0 + (np.arange(A) % 5) / 5.0)
    y = (1.0 + 2e-4 + sig[None, :] * z).astype(f32)
    y[:, 0] = 1.0                                           # asset 0 is cash
    raw = rs.standard_normal((S, A)).astype(f32)
    if kind == "raw":                                       # policy-like raw scores → softmax branch
        act = raw
    elif kind == "simplex":                                 # already a simplex → passes through
        ex = np.exp(raw.astype(np.float64))
        act = (ex / ex.sum(1, keepdims=True)).astype(f32)
    elif kind == "nonneg":                                  # non-negative, not a simplex → quirk Q1 pass-through
        ex = np.exp(raw.astype(np.float64))
        scale = 1.0 + 0.1 * (rs.random_sample((S, 1)) - 0.5)
        act = (ex / ex.sum(1, keepdims=True) * scale).astype(f32)
    elif kind == "near1neg":                                # sums to ≈1 but has negatives → isclose short-circuits
        ex = np.exp(raw.astype(np.float64))
        base = ex / ex.sum(1, keepdims=True)
        d = rs.standard_normal((S, A)) * (0.5 / A)
        d -= d.mean(1, keepdims=True)
        act = (base + d).astype(f32)
    else:
        raise ValueError(kind)
    return act, y


def input_hash(act, y) -> str:
    return hashlib.sha256(act.tobytes() + y.tobytes()).hexdigest()


def run_env_case(A, W, S, seed, kind, commission=0.0, F=5, store_inputs=True, weights_every=1, reward_variants=False):
    te = live.load_env_module(A, W, commission, patch_maximum=commission > 0)
    import env.reward as rw
    torch.set_num_threads(1)
    act, y = make_inputs(seed, kind, S, A)
    env = te.TradingEnv()
    feats = torch.zeros(A, W, F)
    obs0 = env.reset(feats.clone())
    snaps = sorted(set(s for s in [0, 1, 2, W - 2, W - 1, W, W + 1, W + 3, 2 * W, S - 1] if 0 <= s < S))
    values = np.zeros(S, f32); rewards = np.zeros(S, f32)
    idx = np.zeros(S, np.int32); full = np.zeros(S, np.uint8)
    weights = []
    obs_w = []
    rv = {"returns": np.zeros(S, f32), "log_returns": np.zeros(S, f32), "sharpe_ratio": np.zeros(S, np.float64)}
    for s in range(S):
        a = torch.from_numpy(act[s]).reshape(1, A, 1)
        r, obs = env.step(a, feats.clone(), torch.from_numpy(y[s]))
        values[s] = float(env.value); rewards[s] = float(r)
        idx[s] = env.weights.idx; full[s] = env.weights.is_full
        if s % weights_every == 0 or s == S - 1:
            weights.append(env.weights.get_last().numpy().copy())
        if s in snaps:
            obs_w.append(obs[:, :, -1].numpy().copy())
        if reward_variants:
            with np.errstate(all="ignore"):
                rv["returns"][s] = float(env.reward.returns())
                rv["log_returns"][s] = float(env.reward.log_returns())
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    rv["sharpe_ratio"][s] = float(env.reward.sharpe_ratio())
    out = dict(A=A, W=W, F=F, S=S, seed=seed, kind=kind, commission=commission,
               initial_cash=25000.0, reward_scale=1.0, risk_free=float(rw.RISK_FREE_RATE),
               values=values, rewards=rewards, idx=idx, is_full=full,
               weights=np.stack(weights), weights_every=weights_every,
               obs_w_steps=np.array(snaps, np.int32), obs_w=np.stack(obs_w),
               reset_obs_w=obs0[:, :, -1].numpy().copy(),
               input_sha256=input_hash(act, y))
    if store_inputs:
        out["actions"] = act; out["y"] = y
    if reward_variants:
        out.update({f"rv_{k}": v for k, v in rv.items()})
    return out


def main():
    os.makedirs(HERE, exist_ok=True)
    cases = [
        # (A, W, S, seed, kind, commission, store_inputs, weights_every, reward_variants)
        (11, 50, 1000, 123, "raw", 0.0, True, 1, True),
        (11, 50, 1000, 7, "simplex", 0.0, True, 10, False),
        (11, 50, 400, 8, "nonneg", 0.0, True, 10, False),
        (11, 50, 400, 9, "near1neg", 0.0, True, 10, False),
        (11, 8, 64, 10, "raw", 0.0, True, 1, False),
        (50, 50, 1000, 11, "raw", 0.0, True, 10, False),
        (100, 50, 1000, 12, "raw", 0.0, False, 50, False),
        (500, 50, 1000, 13, "raw", 0.0, False, 100, False),
        (11, 50, 300, 21, "raw", 0.0025, True, 10, True),
        (100, 50, 300, 22, "simplex", 0.0025, False, 50, False),
        (500, 50, 300, 23, "raw", 0.0025, False, 100, False),
    ]
    for (A, W, S, seed, kind, c, store, every, rv) in cases:
        out = run_env_case(A, W, S, seed, kind, c, store_inputs=store, weights_every=every, reward_variants=rv)
        name = f"env_A{A}_W{W}_{kind}_c{str(c).replace('.', 'p')}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, "V_end", out["values"][-1], "bytes", os.path.getsize(os.path.join(HERE, name)))


if __name__ == "__main__":
    main()
