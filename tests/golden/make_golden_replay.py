"""Golden vectors from the reference's OWN off-policy buffers, replay/buffer.py and replay/traj_buffer.py.

    python tests/golden/make_golden_replay.py  →  tests/golden/replay_buffer.npz

Neither module imports as shipped: both do `from loader.data_loader import StockDataLoader` (buffer.py:4 — the package
is `data/`, there is no `loader/`) and buffer.py:1 imports `PERCENT_LATEST`, which no config defines.  Exactly like
make_golden_ffd.py stubs tensordict/statsmodels, this script provides
  * a module `loader.data_loader` whose `StockDataLoader` serves a toy windowed dataset through the three accessors the
    buffers call (`get_train_data().dataset`, `get_num_features()`, `get_train_len()`), and
  * the missing constants on `config.base` (BUFFER_SIZE / BATCH_SIZE / INCLUDE_LAST come from config/dsac.py in the
    reference; PERCENT_LATEST is chosen here),
and then runs the reference classes UNMODIFIED: `add` over three epochs of loader steps exactly as
train/off_policy.py:76-89 calls it, then `sample()` under a seeded torch RNG.  The random draws are recovered by
replaying the same torch.randint / torch.randperm calls from the same seed.

The dataset is a plain tensor [T', A, W, F] of windows (`dataset[i][0]` with the buffer's 1-element index tensor `i`
yields window i, buffer.py:65-67); window i = table rows [i, i+W) — data/instrument.py:351-353.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import live_reference as live  # noqa: E402

A, W, F, TRAIN_LEN, BATCH, EPOCHS_KEPT = 7, 4, 5, 30, 4, 6


def install_stubs(table):
    """table [T, A, F-1] → the windowed dataset and the stub loader module."""
    T = table.shape[0]
    nwin = T - W + 1
    windows = torch.zeros(nwin, A, W, F)
    for i in range(nwin):
        windows[i, :, :, :F - 1] = table[i:i + W].permute(1, 0, 2)

    class _Train:
        dataset = windows

    class StockDataLoader:
        def get_train_data(self):
            return _Train

        def get_num_features(self):
            return F

        def get_train_len(self):
            return TRAIN_LEN

    pkg = types.ModuleType("loader"); pkg.__path__ = []
    mod = types.ModuleType("loader.data_loader"); mod.StockDataLoader = StockDataLoader
    sys.modules["loader"], sys.modules["loader.data_loader"] = pkg, mod
    live._ensure_path()
    import config.base as cb
    epoch_len = TRAIN_LEN - 2 * (W - 1)
    cb.NUM_ASSETS, cb.WINDOW_SIZE, cb.BATCH_SIZE = A, W, BATCH
    cb.BUFFER_SIZE, cb.PERCENT_LATEST, cb.INCLUDE_LAST = EPOCHS_KEPT * epoch_len, 0.5, True
    return StockDataLoader(), windows


def main():
    rs = np.random.RandomState(5)
    T = TRAIN_LEN + W + 1
    table = torch.tensor(rs.standard_normal((T, A, F - 1)), dtype=torch.float32)
    data, windows = install_stubs(table)
    import replay.buffer as rb
    import replay.traj_buffer as tb
    importlib.reload(rb); importlib.reload(tb)

    n_epochs = EPOCHS_KEPT + 2                                   # wraps around max_epochs
    acts = torch.tensor(rs.standard_normal((n_epochs, TRAIN_LEN, A, 1)), dtype=torch.float32)
    rews = torch.tensor(0.01 * rs.standard_normal((n_epochs, TRAIN_LEN)), dtype=torch.float32)
    out = dict(A=A, W=W, F=F, train_len=TRAIN_LEN, batch=BATCH, epochs_kept=EPOCHS_KEPT, n_epochs=n_epochs, percent_latest=0.5,
               table=table.numpy(), acts=acts.numpy(), rews=rews.numpy())
    for tag, cls in (("buf", rb.ReplayBuffer), ("traj", tb.ReplayBuffer)):
        buf = cls(data)
        for e in range(n_epochs):
            for step in range(1, TRAIN_LEN):                     # off_policy.py:76-89: add() for every step >= 1
                buf.add(e, step, acts[e, step], rews[e, step].reshape(1))
        out[f"{tag}_i"] = buf.buffer["i"].numpy(); out[f"{tag}_a"] = buf.buffer["a"].numpy(); out[f"{tag}_r"] = buf.buffer["r"].numpy()
        L = buf.epoch_len
        for k in range(3):
            torch.manual_seed(100 + k)
            s, a, r, s2 = buf.sample()
            torch.manual_seed(100 + k)                           # replay the draws
            if tag == "buf":                                     # buffer.py:45-49
                n_last = int(buf.num_epochs_last)
                ep = torch.cat((torch.tensor([buf.epoch] * n_last, dtype=torch.long),
                                torch.randint(0, buf.epoch + 1, (BATCH - n_last,))))
                st = torch.randint(0, L - W - 1, (BATCH,))
                out["buf_epoch"] = buf.epoch
            else:                                                # traj_buffer.py:52-60
                n = buf.max_epoch if buf.full else buf.curr_epoch
                ep = torch.cat([torch.tensor([buf.curr_epoch]), torch.randperm(n)[:BATCH - 1]])
                st = torch.randint(0, L - W - 1, (1,)).repeat(BATCH)
                out["traj_curr_epoch"], out["traj_full"] = buf.curr_epoch, buf.full
            out[f"{tag}{k}_epochs"], out[f"{tag}{k}_starts"] = ep.numpy(), st.numpy()
            out[f"{tag}{k}_s"], out[f"{tag}{k}_a"], out[f"{tag}{k}_r"], out[f"{tag}{k}_s2"] = s.numpy(), a.numpy(), r.numpy(), s2.numpy()
    np.savez_compressed(os.path.join(HERE, "replay_buffer.npz"), **out)
    print("replay golden:", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 1})


if __name__ == "__main__":
    main()
