"""Golden vectors from the live reference RolloutBuffer (replay/rollout_buffer.py) and PG._reward (agent/pg/pg.py).

    python tests/golden/make_golden_buffers.py  →  tests/golden/rollout_buffer.npz, tests/golden/pg_reward.npz
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import live_reference as live  # noqa: E402


def rollout():
    A, W, F, L, BS = 6, 5, 4, 40, 8
    rb = live.load_rollout_buffer_module(A, W, BS)
    rs = np.random.RandomState(3)
    prices = torch.tensor(1 + 0.01 * rs.standard_normal((A, L)), dtype=torch.float32)       # [A, L] like train_prices
    buf = rb.RolloutBuffer(F, L, prices)
    buf.reset()
    S = torch.tensor(rs.standard_normal((L, A, W, F)), dtype=torch.float32)
    Act = torch.tensor(rs.standard_normal((L, 1, A, 1)), dtype=torch.float32)
    V = torch.tensor(25000 + rs.standard_normal(L), dtype=torch.float32)
    R = torch.tensor(0.01 * rs.standard_normal(L), dtype=torch.float32)
    for step in range(1, L):                              # on_policy.py:59-67: add() is called for steps >= 1
        buf.add(S[step], Act[step].reshape(A, 1), V[step], R[step])
    np.random.seed(11)
    batches = list(buf.sample_random())
    np.random.seed(11)
    idxs = np.random.choice(np.arange(1, buf.epoch_len), ((buf.epoch_len - 1) // BS, BS), replace=False)
    seq = list(buf.sample())
    out = dict(A=A, W=W, F=F, L=L, BS=BS, prices=prices.numpy(), S=S.numpy(), Act=Act.numpy(), V=V.numpy(), R=R.numpy(),
               idxs=idxs, n_seq=len(seq))
    for b, t in enumerate(batches):
        for j, name in enumerate(["s", "a", "r", "pv", "pa", "p"]):
            out[f"rand{b}_{name}"] = t[j].numpy()
    for b, t in enumerate(seq):
        for j, name in enumerate(["s", "a", "r", "pv", "pa", "p"]):
            out[f"seq{b}_{name}"] = t[j].numpy()
    np.savez_compressed(os.path.join(HERE, "rollout_buffer.npz"), **out)
    print("rollout batches", len(batches), "seq", len(seq), "epoch_len", buf.epoch_len)


def pg_reward():
    """PG._reward forward value and autograd gradient w.r.t. the raw action (agent/pg/pg.py:40-82)."""
    live._ensure_path()
    import config.base as cb
    A, W = 12, 8
    cb.NUM_ASSETS, cb.WINDOW_SIZE, cb.COMISSION = A, W, 0.0
    import net.lsre_cann as net
    import agent.pg.pg as pg
    importlib.reload(net); importlib.reload(pg)
    out = {}
    rs = np.random.RandomState(4)
    for mode in ("log_returns", "returns", "sharpe_ratio"):
        pg.REWARD = mode
        agent = pg.PG.__new__(pg.PG)                     # _reward needs no network
        B = 16
        a = torch.tensor(rs.standard_normal((B, A, 1)), dtype=torch.float32, requires_grad=True)
        pv = torch.tensor(25000 + 100 * rs.standard_normal((B, 1, 1)), dtype=torch.float32)
        pa = torch.softmax(torch.tensor(rs.standard_normal((B, A, 1)), dtype=torch.float32), dim=1)
        p = torch.tensor(1 + 0.01 * rs.standard_normal((B, A, 1)), dtype=torch.float32)
        r = agent._reward(a, pv, pa, p)
        r.backward()
        out.update({f"{mode}_a": a.detach().numpy(), f"{mode}_pv": pv.numpy(), f"{mode}_pa": pa.numpy(), f"{mode}_p": p.numpy(),
                    f"{mode}_r": r.detach().numpy(), f"{mode}_grad": a.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "pg_reward.npz"), **out)
    print("pg_reward", {k: v.shape for k, v in out.items() if k.endswith("_r") or k.endswith("_grad")})


if __name__ == "__main__":
    rollout()
    pg_reward()
