"""Golden rollouts of the reference's on-policy loop with its OWN agent in the loop.

    python tests/golden/make_golden_pg_rollout.py  →  tests/golden/pg_rollout_A11_W{50,16}_softmax.npz, pg_rollout_A11_W16_init.npz

Runs the body of `Train._rollout` (train/on_policy.py:56-67) literally — live `agent.pg.pg.PG(F).act`
(net/lsre_cann.py LSRE-CANN policy, eval mode, seeded init), live `env.sim.trading_env.TradingEnv`, live
`replay.rollout_buffer.RolloutBuffer` — over a seeded toy loader that yields `(datetime, prices [A], data [A, W, F])`
like the reference's loader (data/instrument_pool.py:512-533; window i = table rows [i, i+W), instrument.py:351-353).
Only the logger call is dropped.

The fixture stores what a run WITHOUT the reference needs to replay the loop: the feature table, the price relatives, and
per step the action the policy produced, the weight channel of the observation it was shown, env.value and the reward;
plus the rollout buffer's a / v / r rows.  (The buffer's s rows are asserted here to be exactly the observations shown to
the agent — s[k] = obs of step k + W - 1 — so the test rebuilds them from the table and the stored weight channels.)
The GPU test drives `compat.TradingEnv` through the same loop with an `act` that asserts it is shown the same
observation (≤ 1e-5) and returns the recorded action: config 1 of BASELINE.json (PG agent, 1 env, 11 assets, window 50).
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


def toy_loader(A, W, F, L, seed):
    rs = np.random.RandomState(seed)
    T = L + W - 1
    table = torch.tensor(rs.random_sample((T, A, F - 1)), dtype=torch.float32)         # min-max scaled features in [0, 1]
    sig = 0.01 * (1.0 + (np.arange(A) % 5) / 5.0)
    y = torch.tensor(1.0 + 2e-4 + sig[None, :] * rs.standard_normal((L, A)), dtype=torch.float32)
    y[:, 0] = 1.0                                                                      # cash
    items = []
    for i in range(L):
        data = torch.zeros(A, W, F)
        data[:, :, :F - 1] = table[i:i + W].permute(1, 0, 2)
        items.append((i, y[i].clone(), data))
    return table, y, items


def run_case(A, W, F, L, seed, batch=8, head=None, tag=""):
    """head=None: the policy exactly as PG(F) initialises it — its outputs are all ≈ +0.2, so the env's AND-rule (quirk
    Q1, trading_env.py:58) passes them through un-normalised and the value grows ≈ 2.3x per step.
    head=(gain, bias): a 'checkpoint' that rescales only the parameters of the last linear layer (cann.out) so that the
    scores have mixed signs and take the softmax branch — parameters, not code."""
    live._ensure_path()
    te = live.load_env_module(A, W)                       # patches config.base and reloads the env modules
    rbm = live.load_rollout_buffer_module(A, W, batch)
    import net.lsre_cann as net
    import agent.pg.pg as pg
    importlib.reload(net); importlib.reload(pg)
    table, y, train_dl = toy_loader(A, W, F, L, seed)
    torch.manual_seed(seed)
    torch.set_num_threads(1)
    agent = pg.PG(F)
    if head is not None:
        with torch.no_grad():
            agent.policy.cann.out.weight.mul_(head[0]); agent.policy.cann.out.bias.mul_(head[0]).add_(head[1])
    env = te.TradingEnv()
    buffer = rbm.RolloutBuffer(F, L, y.T.contiguous())    # train_prices [A, L] (rollout_buffer.py:12)

    acts = np.zeros((L, A), np.float32); obs_w = np.zeros((L, A, W), np.float32)
    vals = np.zeros(L, np.float32); rews = np.zeros(L, np.float32)
    seen = []
    # ---- train/on_policy.py:56-67, verbatim apart from `self.` and the logger ----
    agent.training_mode(False)
    buffer.reset()
    for step, (datetime, prices, data) in enumerate(train_dl):
        if step == 0:
            s = env.reset(data)
        else:
            a = agent.act(s)
            obs_w[step] = s[:, :, -1].numpy(); seen.append(s.clone())                  # (recording only)
            r, s_ = env.step(a, data, prices)
            buffer.add(s, a, env.value, r)
            s = s_
            acts[step] = a.flatten().numpy(); vals[step] = float(env.value); rews[step] = float(r)
    # ------------------------------------------------------------------------------
    off = W - 1
    for k in range(1, buffer.epoch_len):                   # slot k holds the obs shown at step k + W - 1 (the last slot stays empty)
        assert np.array_equal(np.asarray(buffer.s[k], np.float32), seen[k + off - 1].numpy())
    np.random.seed(seed)
    batches = list(buffer.sample_random())
    np.random.seed(seed)
    idxs = np.random.choice(np.arange(1, buffer.epoch_len), ((buffer.epoch_len - 1) // batch, batch), replace=False)
    out = dict(A=A, W=W, F=F, L=L, seed=seed, batch=batch, table=table.numpy(), y=y.numpy(), actions=acts, obs_w=obs_w,
               values=vals, rewards=rews, reset_obs_w=np.asarray(seen[0][:, :, -1]),
               buf_a=np.asarray(buffer.a, np.float32), buf_v=np.asarray(buffer.v, np.float32),
               buf_r=np.asarray(buffer.r, np.float32), idxs=idxs,
               final_value=float(env.value), ring_idx=int(env.weights.idx), ring_full=bool(env.weights.is_full))
    for j, name in enumerate(["s", "a", "r", "pv", "pa", "p"]):      # first random minibatch (rollout_buffer.py:125-140)
        out[f"rand0_{name}"] = batches[0][j].numpy()
    path = os.path.join(HERE, f"pg_rollout_A{A}_W{W}{tag}.npz")
    np.savez_compressed(path, **out)
    print(path, "final value", float(env.value), "softmax-branch steps",
          int(sum(abs(acts[s].sum() - 1) > 1e-6 and acts[s].min() < 0 for s in range(1, L))), "of", L - 1)


if __name__ == "__main__":
    run_case(11, 50, 5, 140, seed=21, head=(40.0, -8.3), tag="_softmax")   # BASELINE config 1 shape, softmax branch
    run_case(11, 16, 5, 90, seed=22, head=(40.0, -8.3), tag="_softmax")    # ring wraps five times
    run_case(11, 16, 5, 60, seed=23, tag="_init")                          # untouched init: Q1 pass-through
