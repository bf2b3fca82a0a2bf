"""CPU oracle of the buffer layouts — TEST INFRASTRUCTURE (see oracle/env_oracle.py header).

  RolloutOracle  restates replay/rollout_buffer.py:7-142 for ONE env (exactly the reference's shapes);
                 PINNED by tests/golden/rollout_buffer.npz (live reference RolloutBuffer, importable here).
  replay_gather  restates the per-sample body of ReplayBuffer.sample, replay/buffer.py:58-77 (identical in
                 replay/traj_buffer.py:68-87);  ReplayOracle restates `add` (buffer.py:23-38 / traj_buffer.py:26-43) and the
                 two samplers' draw sequences (buffer.py:45-49, traj_buffer.py:52-60).
                 PINNED by tests/golden/replay_buffer.npz: both reference classes executed UNMODIFIED by
                 tests/golden/make_golden_replay.py behind a stub for their one missing import (`loader.data_loader`,
                 buffer.py:4) and the constants no config defines (PERCENT_LATEST).
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


class ReplayOracle:
    """replay/buffer.py:6-38 and replay/traj_buffer.py:6-43 for ONE env: the (i, a, r) index-replay rows."""

    def __init__(self, train_len, A, W, buffer_size, batch_size, percent_latest=0.5):
        self.A, self.W, self.bs = A, W, batch_size
        self.step_offset = 2 * (W - 1)                                            # :11
        self.epoch_len = train_len - self.step_offset                             # :12
        self.max_epoch = buffer_size // self.epoch_len                            # :13
        self.n_last = int(percent_latest * batch_size)                            # buffer.py:14
        self.i = np.zeros((self.max_epoch, self.epoch_len, 1), np.float32)        # the reference stores i as f32 (Q14)
        self.a = np.zeros((self.max_epoch, self.epoch_len, A), np.float32)
        self.r = np.zeros((self.max_epoch, self.epoch_len, 1, 1), np.float32)
        self.curr_epoch, self.newest_epoch, self.full = 0, 0, False

    def add(self, e, i, a, r):
        if i < self.W - 1:                                                        # :31
            return
        ep = int(e % self.max_epoch)
        self.curr_epoch = ep                                                      # traj_buffer.py:35
        self.newest_epoch = max(self.newest_epoch, ep)                            # buffer.py:38
        step = i - self.step_offset                                               # may be negative: wraps like torch indexing
        self.i[ep, step] = i; self.a[ep, step] = np.asarray(a).reshape(self.A); self.r[ep, step] = np.asarray(r).reshape(1)
        if not self.full and ep == self.max_epoch - 1:                            # traj_buffer.py:42
            self.full = True

    def draws(self, sampler):
        """The (epochs, starts) a `sample()` call draws from torch's global RNG, in the reference's call order."""
        import torch
        L, W, B = self.epoch_len, self.W, self.bs
        if sampler == "traj":                                                     # traj_buffer.py:52-60
            n = self.max_epoch if self.full else self.curr_epoch
            ep = torch.cat([torch.tensor([self.curr_epoch]), torch.randperm(n)[:B - 1]])
            st = torch.randint(0, L - W - 1, (1,)).repeat(B)
        else:                                                                     # buffer.py:45-49
            ep = torch.cat((torch.tensor([self.newest_epoch] * self.n_last, dtype=torch.long),
                            torch.randint(0, self.newest_epoch + 1, (B - self.n_last,))))
            st = torch.randint(0, L - W - 1, (B,))
        return ep.numpy(), st.numpy()

    def sample(self, table, sampler):
        ep, st = self.draws(sampler)
        bi = self.i[..., 0].astype(np.int64)[..., None]; br = self.r[..., 0, 0][..., None]
        out = [replay_gather(table, bi, self.a[:, :, None, :], br, int(e), 0, int(s0), self.W) for e, s0 in zip(ep, st)]
        return tuple(np.stack([o[j] for o in out]) for j in range(4)), ep, st
