"""CPU oracle of the buffer layouts — TEST INFRASTRUCTURE (see oracle/env_oracle.py header).

  RolloutOracle  restates replay/rollout_buffer.py:7-142 for ONE env (exactly the reference's shapes);
                 PINNED by tests/golden/rollout_buffer.npz (live reference RolloutBuffer, importable here).
  replay_gather  restates the per-sample body of ReplayBuffer.sample, replay/buffer.py:58-77 (identical in
                 replay/traj_buffer.py:68-87).  The reference module itself is not importable (it imports the
                 non-existent `loader.data_loader`, buffer.py:4) → parity UNPINNED beyond this restatement.
"""
from __future__ import annotations

import numpy as np


class RolloutOracle:
    def __init__(self, num_features, train_len, train_prices, A, W, initial_cash=25000.0, batch_size=64):
        self.F, self.A, self.W, self.bs = num_features, A, W, batch_size
        self.step_offset = W - 1                                                  # :10
        self.epoch_len = train_len - self.step_offset                             # :11
        self.prices = np.asarray(train_prices, np.float64).T[W - 1:][..., None]   # :12  [L, A, 1]
        self.cash = initial_cash
        self.reset()

    def reset(self):                                                              # :29-41
        n = self.epoch_len + 1
        self.s = np.zeros((n, self.A, self.W, self.F)); self.a = np.zeros((n, self.A, 1))
        self.v = np.zeros((n, 1, 1)); self.r = np.zeros((n, 1, 1))
        self.a[0, 0] = 1; self.v[0] = self.cash
        self.step = 1

    def add(self, s, a, v, r):                                                    # :43-57
        if self.step > self.step_offset:
            k = self.step - self.step_offset
            self.s[k] = np.asarray(s); self.a[k] = np.asarray(a).reshape(self.A, 1)
            self.v[k] = np.asarray(v); self.r[k] = np.asarray(r)
        self.step += 1

    def batch(self, idx):                                                         # :125-140
        idx = np.asarray(idx)
        f = np.float32
        return (self.s[idx].astype(f), self.a[idx].astype(f), self.r[idx].astype(f), self.v[idx - 1].astype(f),
                self.a[idx - 1].astype(f), self.prices[idx].astype(f))

    def sample_random(self, rng=np.random):                                       # :103-142
        bs = self.epoch_len if self.bs == -1 else self.bs
        nb = (self.epoch_len - 1) // bs
        idxs = rng.choice(np.arange(1, self.epoch_len), (nb, bs), replace=False)
        return [self.batch(i) for i in idxs], idxs


def replay_gather(feat, bi, ba, br, epoch, env, start, W):
    """buffer.py:58-77 for one sample.  feat [T, A, F-1] (time-major dataset rows), bi [P, L, E], ba [P, L, E, A],
    br [P, L, E].  Returns s [A, W, F], a [A, 1], r [1, 1], s_ [A, W, F]."""
    end = start + W
    a_hist = ba[epoch, start:end + 1, env].T                                      # [A, W+1]   (:59,62)
    r = br[epoch, end - 1, env].reshape(1, 1)                                     # :60,63
    i = int(bi[epoch, end - 1, env])                                              # :65
    A, Fm1 = feat.shape[1], feat.shape[2]
    s = np.empty((A, W, Fm1 + 1), np.float32); s2 = np.empty_like(s)
    s[..., :Fm1] = feat[i:i + W].transpose(1, 0, 2)                               # dataset[i][0]     (:66)
    s2[..., :Fm1] = feat[i + 1:i + 1 + W].transpose(1, 0, 2)                      # dataset[i+1][0]   (:67)
    s[..., -1] = a_hist[:, :-1]                                                   # :69
    s2[..., -1] = a_hist[:, 1:]                                                   # :70
    return s, a_hist[:, -1:].astype(np.float32), r.astype(np.float32), s2         # :72
