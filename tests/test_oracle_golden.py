"""The restated oracle replays every golden fixture produced by the live reference (CPU, no GPU)."""
import numpy as np
import pytest

from oracle.env_oracle import (OracleEnv, REWARD_LOG_RETURNS, REWARD_RETURNS, REWARD_SHARPE, REWARD_STEP_LOG)
from tests import util


def replay(d, reward_mode=REWARD_STEP_LOG):
    A, W, F, S = d["A"], d["W"], d["F"], d["S"]
    env = OracleEnv(1, A, W, F, commission=d["commission"], reward_mode=reward_mode,
                    initial_cash=float(d["initial_cash"]), reward_scale=float(d["reward_scale"]),
                    risk_free=float(d["risk_free"]))
    out = dict(values=np.zeros(S, np.float32), rewards=np.zeros(S, np.float32), idx=np.zeros(S, np.int32),
               is_full=np.zeros(S, np.uint8), weights=[], obs_w=[], reset_obs_w=env.weight_channel()[0].copy())
    wsteps = set(util.weight_steps(d)); snaps = set(int(s) for s in d["obs_w_steps"])
    for s in range(S):
        r, done = env.step(d["actions"][s][None], d["y"][s][None])
        out["values"][s] = env.value[0]; out["rewards"][s] = r[0]
        out["idx"][s] = env.idx[0]; out["is_full"][s] = env.is_full[0]
        if s in wsteps:
            out["weights"].append(env.hist[0, (env.idx[0] - 1) % W].copy())
        if s in snaps:
            out["obs_w"].append(env.weight_channel()[0].copy())
    out["weights"] = np.stack(out["weights"]); out["obs_w"] = np.stack(out["obs_w"])
    return out


@pytest.mark.parametrize("name", util.env_fixture_names())
def test_oracle_replays_reference_rollout(name):
    d = util.load_env_fixture(name)
    o = replay(d)
    # integer state: bit-exact
    np.testing.assert_array_equal(o["idx"], d["idx"])
    np.testing.assert_array_equal(o["is_full"], d["is_full"])
    # float state within the north-star tolerance
    util.assert_values_close(o["values"], d["values"], name)
    util.assert_rewards_close(o["rewards"], d["rewards"], name)
    np.testing.assert_allclose(o["weights"], d["weights"], rtol=util.RTOL_WEIGHT, atol=util.ATOL_WEIGHT)
    np.testing.assert_allclose(o["obs_w"], d["obs_w"], rtol=util.RTOL_WEIGHT, atol=util.ATOL_WEIGHT)
    np.testing.assert_array_equal(o["reset_obs_w"], d["reset_obs_w"])
    # structure of the weight channel (zero padding / ring order) is exact
    np.testing.assert_array_equal(o["obs_w"] == 0, d["obs_w"] == 0)


@pytest.mark.parametrize("name", [n for n in util.env_fixture_names() if "A11_W50_raw" in n])
@pytest.mark.parametrize("mode,key", [(REWARD_RETURNS, "rv_returns"), (REWARD_LOG_RETURNS, "rv_log_returns"),
                                      (REWARD_SHARPE, "rv_sharpe_ratio")])
def test_oracle_reward_variants(name, mode, key):
    """env/reward.py:15-31 evaluated by the live reference after every step."""
    d = util.load_env_fixture(name)
    o = replay(d, reward_mode=mode)
    want = d[key]
    if mode == REWARD_SHARPE:
        assert np.isnan(o["rewards"][0]) and np.isnan(want[0])        # ddof=1 at n == 1 (quirk Q11)
        # the Sharpe ratio divides by a tiny std: compare with a relative bound on the well-conditioned tail
        np.testing.assert_allclose(o["rewards"][5:], want[5:], rtol=2e-4)
    else:
        util.assert_rewards_close(o["rewards"], want, name)
