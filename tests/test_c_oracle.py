"""oracle/pmrl_oracle.c (plain-C restatement, CPU baseline) against the live-reference golden fixtures and the numpy oracle."""
import numpy as np
import pytest

from oracle.c_oracle import COracleEnv
from oracle.env_oracle import OracleEnv
from tests import util


@pytest.mark.parametrize("name", [n for n in util.env_fixture_names() if "A500" not in n])
def test_c_oracle_replays_reference_rollout(name):
    d = util.load_env_fixture(name)
    A, W, S = d["A"], d["W"], d["S"]
    env = COracleEnv(1, A, W, commission=d["commission"])
    vals = np.zeros(S, np.float32); rews = np.zeros(S, np.float32); idx = np.zeros(S, np.int32); full = np.zeros(S, np.uint8)
    for s in range(S):
        r, _ = env.step(d["actions"][s][None], d["y"][s][None])
        vals[s], rews[s], idx[s], full[s] = env.value[0], r[0], env.idx[0], env.is_full[0]
    np.testing.assert_array_equal(idx, d["idx"]); np.testing.assert_array_equal(full, d["is_full"])
    util.assert_values_close(vals, d["values"], name)
    util.assert_rewards_close(rews, d["rewards"], name)


@pytest.mark.parametrize("c", [0.0, 0.0025])
def test_c_oracle_matches_numpy_oracle_batched(c):
    E, A, W, L = 37, 23, 7, 15
    rs = np.random.RandomState(4)
    ce = COracleEnv(E, A, W, episode_len=L, commission=c)
    ne = OracleEnv(E, A, W, 5, episode_len=L, commission=c)
    for s in range(2 * L + 3):
        act = rs.standard_normal((E, A)).astype(np.float32)
        if s % 3 == 1:
            act = np.abs(act) / np.abs(act).sum(1, keepdims=True)
        y = (1 + 0.01 * rs.standard_normal((E, A))).astype(np.float32); y[:, 0] = 1
        r1, d1 = ce.step(act, y)
        r2, d2 = ne.step(act, y)
        np.testing.assert_array_equal(d1, d2)
        util.assert_rewards_close(r1, r2, f"step {s}")
    np.testing.assert_array_equal(ce.idx, ne.idx); np.testing.assert_array_equal(ce.t, ne.t)
    util.assert_values_close(ce.value, ne.value)
    np.testing.assert_allclose(ce.hist, ne.hist, rtol=1e-5, atol=1e-6)
