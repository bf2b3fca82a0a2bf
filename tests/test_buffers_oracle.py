"""oracle/buffers_oracle.py against the live-reference RolloutBuffer golden vectors (CPU)."""
import os

import numpy as np

from oracle.buffers_oracle import RolloutOracle
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
