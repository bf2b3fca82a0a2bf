"""CPU oracle for the environment hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A batched numpy/fp32 restatement of the reference transition, used only by tests/, by
`__graft_entry__.smoke()` and by bench.py's cpu_baseline leg as the *checker*.  Nothing under
pm-rl_b200/ may import it.

Parity status: PINNED for commission == 0 against the live reference (tests/golden/env_*.npz were
produced by importing /root/reference/env with tests/golden/make_golden.py; tests/test_oracle_golden.py
replays them).  For commission > 0 the reference raises TypeError (`torch.maximum(x, )`,
env/sim/trading_env.py:72); the fixtures for c > 0 come from the reference executed with that one call
monkeypatched to relu (the upstream PGPortfolio form it cites at :66) — parity for c > 0 is otherwise
UNPINNED.

Every function cites the reference lines it follows (paths relative to the pm-rl tree).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

REWARD_STEP_LOG, REWARD_RETURNS, REWARD_LOG_RETURNS, REWARD_SHARPE = 0, 1, 2, 3


def isclose_one(s: np.ndarray) -> np.ndarray:
    """torch.isclose(s, 1.0, atol=1e-6) with the hidden default rtol=1e-5, in fp32 (trading_env.py:58)."""
    s = s.astype(f32)
    allowed = f32(1e-6) + np.abs(f32(1e-5) * f32(1.0))
    with np.errstate(invalid="ignore"):
        err = np.abs(s - f32(1.0))
        return (s == f32(1.0)) | (np.isfinite(err) & (err <= allowed))


def normalise_actions(act: np.ndarray, strict: bool = True) -> np.ndarray:
    """trading_env.py:54-60 — un-stabilised softmax iff (!isclose(sum,1) AND min<0) (quirks Q1-Q3).
    strict=False: OR-condition with a max-subtracted softmax (agent/pg/pg.py:52-53)."""
    act = np.ascontiguousarray(act, dtype=f32)
    s = act.sum(axis=1, dtype=f32)
    with np.errstate(invalid="ignore"):
        mn = act.min(axis=1)                      # np.min propagates NaN like torch.min
        neg = mn < 0
    nc = ~isclose_one(s)
    do = (nc & neg) if strict else (nc | neg)
    w = act.copy()
    if do.any():
        sub = act[do]
        if not strict:
            sub = sub - sub.max(axis=1, keepdims=True)
        with np.errstate(over="ignore", invalid="ignore"):
            ex = np.exp(sub).astype(f32)
            w[do] = ex / ex.sum(axis=1, dtype=f32, keepdims=True)
    return w


def commission_mu(w_last: np.ndarray, w: np.ndarray, c: float, max_iter: int = 16) -> np.ndarray:
    """Transaction remainder factor (trading_env.py:67-74 with the upstream relu in place of the broken
    one-argument torch.maximum): mu <- (1 - c*w_last[0] - (2c-c^2)*sum_{i>=1} relu(w_last[i]-mu*w[i])) / (1 - c*w[0]),
    start mu=1-2c+c^2, stop when |mu-mu_last| <= 1e-10 (fp32 → an exact fixed point) or after max_iter."""
    E = w.shape[0]
    cf = f32(c)
    c2 = f32(2.0 * c - c * c)
    mu = np.full(E, f32(1.0 - 2.0 * c + c * c), dtype=f32)
    mu_last = np.ones(E, dtype=f32)
    denom = f32(1.0) - cf * w[:, 0]
    cw = cf * w_last[:, 0]
    active = np.abs(mu - mu_last) > f32(1e-10)
    it = 0
    while active.any() and it < max_iter:
        mu_last = np.where(active, mu, mu_last)
        part = np.maximum(w_last[:, 1:] - mu[:, None] * w[:, 1:], f32(0)).sum(axis=1, dtype=f32)
        numer = (f32(1.0) - cw) - c2 * part
        mu = np.where(active, (numer / denom).astype(f32), mu)
        active = active & (np.abs(mu - mu_last) > f32(1e-10))
        it += 1
    return mu.astype(f32)


class OracleEnv:
    """Batched restatement of TradingEnv + ActionBuffer + Reward over E independent envs.

    Tables are given in the reference's natural time-major layout: close [T, A], feat [T, A, F-1].
    Window/time alignment (data/instrument.py:79,351-356; train/on_policy.py:59-66): env e at local
    step k sees rows [t0+k, t0+k+W) and the price relative of row t0+k+W-1.
    """

    def __init__(self, E, A, W, F, close=None, feat=None, t0=None, episode_len=0,
                 initial_cash=25000.0, commission=0.0, reward_mode=REWARD_STEP_LOG, reward_scale=1.0,
                 risk_free=0.04, strict_reference=True, mu_max_iter=16):
        self.E, self.A, self.W, self.F = E, A, W, F
        self.close = None if close is None else np.ascontiguousarray(close, dtype=f32)
        self.feat = None if feat is None else np.ascontiguousarray(feat, dtype=f32)
        self.t0 = np.zeros(E, np.int32) if t0 is None else np.asarray(t0, np.int32).copy()
        self.episode_len = int(episode_len)
        self.initial_cash = f32(initial_cash)
        self.commission = float(commission)
        self.reward_mode = int(reward_mode)
        self.reward_scale = f32(reward_scale)
        self.risk_free = float(risk_free)
        self.strict = bool(strict_reference)
        self.mu_max_iter = int(mu_max_iter)
        self.value = np.empty(E, f32)
        self.hist = np.empty((E, W, A), f32)
        self.idx = np.empty(E, np.int32)
        self.is_full = np.zeros(E, np.uint8)
        self.t = np.zeros(E, np.int32)
        self.sharpe = np.zeros((E, 3), np.float64)       # running (n, mean, M2) of gross returns
        self.ep_return = np.zeros(E, f32)
        self.reset()

    # -- TradingEnv.reset (trading_env.py:28-29) + ActionBuffer.reset (weight_buffer.py:46-50) --
    def reset(self, mask=None):
        m = np.ones(self.E, bool) if mask is None else np.asarray(mask).astype(bool)
        self.value[m] = self.initial_cash
        self.hist[m] = 0
        self.hist[m, 0, 0] = 1
        self.idx[m] = 1
        self.is_full[m] = 0
        self.t[m] = 0
        self.sharpe[m] = 0
        self.ep_return[m] = 0

    # -- ActionBuffer.get_all (weight_buffer.py:32-44), batched → [E, A, W] --
    def weight_channel(self):
        E, A, W = self.E, self.A, self.W
        g = np.zeros((E, A, W), f32)
        for e in range(E):
            if self.is_full[e]:
                g[e] = self.hist[e].T                         # ring order once full (quirk Q8)
            else:
                i = int(self.idx[e])
                g[e, :, W - i:] = self.hist[e, :i].T          # zero front padding, chronological
        return g

    def window_rows(self):
        return self.t0.astype(np.int64) + self.t.astype(np.int64)

    # -- features[:, :, -1] = weights.get_all() (trading_env.py:32,103) on the gathered window --
    def obs(self):
        E, A, W, F = self.E, self.A, self.W, self.F
        out = np.empty((E, A, W, F), f32)
        r0 = self.window_rows()
        for e in range(E):
            out[e, :, :, : F - 1] = self.feat[r0[e]: r0[e] + W].transpose(1, 0, 2)
        out[..., F - 1] = self.weight_channel()
        return out

    def step(self, actions, y=None):
        """TradingEnv.step (trading_env.py:54-100).  Returns (reward[E] f32, done[E] u8)."""
        E, A, W = self.E, self.A, self.W
        act = np.ascontiguousarray(np.asarray(actions, dtype=f32).reshape(E, A))
        auto = (self.t >= self.episode_len) if self.episode_len > 0 else np.zeros(E, bool)
        k_new = self.t + 1
        if y is None:
            rows = self.t0.astype(np.int64) + k_new + W - 1
            rows = np.where(auto, 1, rows)                                   # dummy row for envs being reset
            y = (self.close[rows] / self.close[rows - 1]).astype(f32)        # instrument.py:79
        else:
            y = np.ascontiguousarray(np.asarray(y, dtype=f32).reshape(E, A))

        w = normalise_actions(act, self.strict)                              # :58-60
        ar = np.arange(E)
        last = (self.idx - 1) % W
        w_last = self.hist[ar, last]                                         # :63, weight_buffer.py:30
        V_prev = self.value.copy()
        V = V_prev
        if self.commission > 0:                                              # :67-75
            mu = commission_mu(w_last, w, self.commission, self.mu_max_iter)
            V = (mu * V_prev).astype(f32)
        with np.errstate(all="ignore"):
            port = (V[:, None] * (w * y)).astype(f32)                        # :78
            Vn = port.sum(axis=1, dtype=f32)                                 # :79
            wn = (port / Vn[:, None]).astype(f32)                            # :83
            ret = (Vn / V).astype(f32)                                       # :88
            if self.reward_mode == REWARD_STEP_LOG:
                r = (np.log(ret) * self.reward_scale).astype(f32)            # :99
            elif self.reward_mode == REWARD_RETURNS:
                r = ((Vn / V_prev) * self.reward_scale).astype(f32)          # reward.py:20-21
            elif self.reward_mode == REWARD_LOG_RETURNS:
                r = (np.log((Vn / V_prev).astype(f32)) * self.reward_scale).astype(f32)   # reward.py:23-24
            else:                                                            # reward.py:26-31 (running)
                g = Vn.astype(np.float64) / V_prev.astype(np.float64)
                n = self.sharpe[:, 0] + 1.0
                d1 = g - self.sharpe[:, 1]
                mean = self.sharpe[:, 1] + d1 / n
                m2 = self.sharpe[:, 2] + d1 * (g - mean)
                sd = np.sqrt(m2 / (n - 1.0))                                 # ddof=1 → NaN at n == 1 (Q11)
                r = (((mean - self.risk_free) / sd) * float(self.reward_scale)).astype(f32)
                live = ~auto
                self.sharpe[live, 0] = n[live]
                self.sharpe[live, 1] = mean[live]
                self.sharpe[live, 2] = m2[live]

        live = ~auto
        le = ar[live]
        self.hist[le, self.idx[live]] = wn[live]                             # weight_buffer.py:21
        idx_new = (self.idx + 1) % W                                         # :22
        self.is_full[live & (idx_new == 0)] = 1                              # :25-26
        self.idx[live] = idx_new[live]
        self.value[live] = Vn[live]
        self.t[live] = k_new[live]
        self.ep_return[live] = (self.ep_return[live] + r[live]).astype(f32)
        done = np.zeros(E, np.uint8)
        if self.episode_len > 0:
            done[live & (k_new == self.episode_len)] = 1
        reward = np.where(live, r, f32(0)).astype(f32)
        if auto.any():                                                       # on_policy.py:60-61
            self.reset(auto)
        self.last_w = np.where(live[:, None], wn, self.hist[ar, 0])          # for tests
        return reward, done
