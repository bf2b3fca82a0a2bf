"""Single-environment, op-for-op torch-CPU port of the reference step — TEST / BASELINE INFRASTRUCTURE.

The reference env is pure Python over tiny ATen CPU ops (≈20 dispatches per step); its cost is
dispatch overhead, not arithmetic.  The reference tree cannot travel to the GPU box, so bench.py's `cpu_baseline`
/ `--impl reference` leg times THIS port there (`kind: "port"`): it issues the same torch operations in the same
order on the same shapes, including the per-step host copies the reference makes for its `info` lists
(env/sim/trading_env.py:80,85,90,100).  tests/test_ref_port.py checks it bit-for-bit against the live reference
where that is available, and against the golden fixtures everywhere.

Follows env/sim/trading_env.py:54-105 and env/sim/weight_buffer.py:13-44.  Parity: pinned (c == 0).
"""
from __future__ import annotations

import torch


class RingPort:
    """weight_buffer.py:6-50."""

    def __init__(self, window: int, assets: int):
        self.W, self.A = window, assets
        self.clear()

    def clear(self):
        self.rows = torch.zeros((self.W, self.A))
        self.rows[0, 0] = 1
        self.pos = 1
        self.wrapped = False

    def push(self, w):
        if w.shape != (self.A,):
            raise ValueError(f"Action must have shape ({self.A},), got {w.shape}")
        self.rows[self.pos] = w.detach()
        self.pos = (self.pos + 1) % self.W
        if self.pos == 0:
            self.wrapped = True

    def newest(self):
        return self.rows[(self.pos - 1) % self.W]

    def as_channel(self):
        if self.wrapped:
            m = self.rows
        else:
            m = torch.concat((torch.zeros((self.W - self.pos, self.A)), self.rows[:self.pos]), axis=0)
        return m.T


class RefPortEnv:
    def __init__(self, assets: int, window: int, initial_cash=25000, commission=0.0, reward_scale=1):
        self.A, self.W = assets, window
        self.cash0, self.c, self.scale = initial_cash, commission, reward_scale
        self.ring = RingPort(window, assets)
        self.value = initial_cash
        self._new_log()

    def _new_log(self):
        self.info = {"values": [self.cash0], "actions": [self.ring.newest().flatten()], "rewards": [0], "returns": [0]}

    def reset(self, features):
        self.value = self.cash0
        self.ring.clear()
        features[:, :, -1] = self.ring.as_channel()
        self._new_log()
        return features

    def step(self, action, features, prices):
        w = action.flatten()
        y = prices.flatten()
        if not torch.isclose(torch.sum(w), torch.tensor(1.0), atol=1e-6) and torch.min(action) < 0:
            e = torch.exp(w)
            w = e / torch.sum(e)
        prev = self.ring.newest()
        c = self.c
        if c > 0:
            mu_old = 1
            mu = 1 - 2 * c + c ** 2
            while abs(mu - mu_old) > 1e-10:
                mu_old = mu
                top = 1 - (c * prev[0]) - (2 * c - c ** 2) * torch.sum(torch.clamp_min(prev[1:] - mu * w[1:], 0))
                mu = top / (1 - c * w[0])
            self.value = mu * self.value
        holdings = self.value * (w * y)
        total = torch.sum(holdings)
        self.info["values"].append(total)
        w = holdings / total
        self.ring.push(w)
        self.info["actions"].append(w.detach().flatten().cpu().numpy())
        gross = total / self.value
        self.value = total.detach()
        self.info["returns"].append(gross.detach().cpu().numpy())
        r = torch.log(gross) * self.scale
        self.info["rewards"].append(r.detach().cpu().numpy())
        features[:, :, -1] = self.ring.as_channel()
        return r, features


def time_env(kind: str, assets: int, window: int, steps: int, warmup: int = 50, features: int = 5, seed: int = 0,
             commission: float = 0.0, policy: bool = False):
    """Steps/s of ONE reference env on one thread: pre-materialised features / price relatives, raw
    N(0,1) scores as actions (the softmax branch, like the GPU arm's "random actions").

    kind="live": the reference's own `env.sim.trading_env.TradingEnv` (tree found by oracle/live_reference.locate();
                 for commission > 0 the one broken call at trading_env.py:72 is patched to relu, see live_reference).
    kind="port": RefPortEnv above (the same torch ops in the same order) — what runs where the tree is absent.
    policy=True: BASELINE config 1 — the reference's PG agent (LSRE-CANN policy) produces every action from the
                 observation, as in train/on_policy.py:59-67 (live only)."""
    import time
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(seed)
    agent = None
    if kind == "live":
        from oracle import live_reference as live
        te = live.load_env_module(assets, window, commission, patch_maximum=commission > 0)
        env = te.TradingEnv()
        if policy:
            agent = live.load_pg_agent(assets, window, features, seed)
    else:
        if policy:
            raise RuntimeError("the PG policy is reference code (agent/pg, net/lsre_cann) and needs the live tree")
        env = RefPortEnv(assets, window, commission=commission)
    feat = torch.rand(assets, window, features, generator=g)
    n = 64
    acts = torch.randn(n, 1, assets, 1, generator=g)
    ys = 1.0 + 0.01 * torch.randn(n, assets, generator=g)
    s_obs = env.reset(feat)

    def one(s, s_obs):
        a = agent.act(s_obs) if agent is not None else acts[s % n]
        return env.step(a, feat, ys[s % n])[1]

    for s in range(warmup):
        s_obs = one(s, s_obs)
    t0 = time.perf_counter()
    for s in range(steps):
        if s % 1000 == 999:
            s_obs = env.reset(feat)              # keeps the info lists bounded like an episode boundary
        s_obs = one(s, s_obs)
    dt = time.perf_counter() - t0
    return steps / dt


def time_port(assets: int, window: int, steps: int, warmup: int = 50, features: int = 5, seed: int = 0):
    return time_env("port", assets, window, steps, warmup, features, seed)


def _worker(args):
    return time_env(*args)


def time_all_cores(kind: str, assets: int, window: int, steps: int, procs: int, commission: float = 0.0, policy: bool = False):
    """Sum of steps/s over `procs` independent single-thread processes."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        rates = pool.map(_worker, [(kind, assets, window, steps, 50, 5, i, commission, policy) for i in range(procs)])
    return sum(rates), rates


def time_port_all_cores(assets: int, window: int, steps: int, procs: int):
    return time_all_cores("port", assets, window, steps, procs)
