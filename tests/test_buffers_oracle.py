"""oracle/buffers_oracle.py against the live-reference RolloutBuffer golden vectors (CPU)."""
import os

import numpy as np

from oracle.buffers_oracle import ReplayOracle, RolloutOracle
from tests import util


def load(name):
    z = np.load(os.path.join(util.GOLDEN, name))
    return {k: z[k] for k in z.files}


def fill(g, buf):
    L, A = int(g["L"]), int(g["A"])
    for step in range(1, L):
        buf.add(g["S"][step], g["Act"][step].reshape(A, 1), g["V"][step], g["R"][step])


def test_rollout_oracle_matches_reference_batches():
    g = load("rollout_buffer.npz")
    A, W, F, L, BS = (int(g[k]) for k in ("A", "W", "F", "L", "BS"))
    buf = RolloutOracle(F, L, g["prices"], A, W, batch_size=BS)
    fill(g, buf)
    for b, idx in enumerate(g["idxs"]):
        got = buf.batch(idx)
        for j, name in enumerate(["s", "a", "r", "pv", "pa", "p"]):
            np.testing.assert_array_equal(got[j], g[f"rand{b}_{name}"], err_msg=f"batch {b} {name}")
    # sequential sampler (rollout_buffer.py:59-101): contiguous slots starting at 1
    for b in range(int(g["n_seq"])):
        got = buf.batch(np.arange(1 + b * BS, 1 + (b + 1) * BS))
        for j, name in enumerate(["s", "a", "r", "pv", "pa", "p"]):
            np.testing.assert_array_equal(got[j], g[f"seq{b}_{name}"], err_msg=f"seq {b} {name}")


def fill_replay(g, buf):
    for e in range(int(g["n_epochs"])):
        for step in range(1, int(g["train_len"])):                      # off_policy.py:76-89
            buf.add(e, step, g["acts"][e, step], g["rews"][e, step])


def test_replay_oracle_matches_reference_buffers_and_samples():
    """tests/golden/replay_buffer.npz = replay/buffer.py and replay/traj_buffer.py executed unmodified (stubbed loader)."""
    import torch
    g = load("replay_buffer.npz")
    A, W, TL, B = (int(g[k]) for k in ("A", "W", "train_len", "batch"))
    for tag, sampler in (("buf", "buffer"), ("traj", "traj")):
        buf = ReplayOracle(TL, A, W, int(g["epochs_kept"]) * (TL - 2 * (W - 1)), B, float(g["percent_latest"]))
        fill_replay(g, buf)
        np.testing.assert_array_equal(buf.i, g[f"{tag}_i"]); np.testing.assert_array_equal(buf.a, g[f"{tag}_a"])
        np.testing.assert_array_equal(buf.r, g[f"{tag}_r"])
        for k in range(3):
            torch.manual_seed(100 + k)
            (s, a, r, s2), ep, st = buf.sample(g["table"], sampler)
            np.testing.assert_array_equal(ep, g[f"{tag}{k}_epochs"]); np.testing.assert_array_equal(st, g[f"{tag}{k}_starts"])
            np.testing.assert_array_equal(s, g[f"{tag}{k}_s"]); np.testing.assert_array_equal(s2, g[f"{tag}{k}_s2"])
            np.testing.assert_array_equal(a, g[f"{tag}{k}_a"]); np.testing.assert_array_equal(r, g[f"{tag}{k}_r"])
