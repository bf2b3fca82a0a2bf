"""The E = 1 drop-in `TradingEnv` (reference call surface) replaying a live-reference rollout on the GPU."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["env_A11_W8_raw_c0p0.npz", "env_A11_W50_raw_c0p0.npz"])
@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_compat_env_drives_like_the_reference(name, device):
    import pmrl_b200
    from pmrl_b200.compat import TradingEnv
    d = util.load_env_fixture(name)
    A, W, F = d["A"], d["W"], d["F"]
    S = min(d["S"], 120)
    env = TradingEnv(pmrl_b200.EnvConfig(num_assets=A, window_size=W, num_features=F))
    feats = torch.zeros(A, W, F, device=device)
    s = env.reset(feats.clone())                                   # train/on_policy.py:61
    np.testing.assert_array_equal(s[:, :, -1].cpu().numpy(), d["reset_obs_w"])
    assert s.device.type == device
    snaps = {int(k): i for i, k in enumerate(d["obs_w_steps"])}
    for k in range(S):
        a = torch.from_numpy(d["actions"][k]).reshape(1, A, 1).to(device)
        f = feats.clone()
        r, s_ = env.step(a, f, torch.from_numpy(d["y"][k]).to(device))          # on_policy.py:64
        assert s_.data_ptr() == f.data_ptr()                       # the caller's tensor is mutated in place (:103)
        if k in snaps:
            np.testing.assert_allclose(s_[:, :, -1].cpu().numpy(), d["obs_w"][snaps[k]], rtol=1e-5, atol=1e-6)
    util.assert_values_close([float(env.value)], [d["values"][S - 1]])
    info = env.info                                                # util/eval.py:15-50 reads these keys
    assert list(info) == ["values", "actions", "rewards", "returns"]
    assert len(info["values"]) == S + 1 and info["values"][0] == 25000.0 and info["rewards"][0] == 0
    util.assert_values_close(info["values"][1:], d["values"][:S])
    util.assert_rewards_close(info["rewards"][1:], d["rewards"][:S])
    np.testing.assert_allclose(info["returns"][1:], np.exp(d["rewards"][:S].astype(np.float64)), rtol=1e-5)
    assert env.weights.idx == d["idx"][S - 1] and env.weights.is_full == bool(d["is_full"][S - 1])
    np.testing.assert_allclose(env.weights.get_last().cpu().numpy(), info["actions"][-1], rtol=0, atol=0)
    assert env.weights.get_all().shape == (A, W)
    if "rv_returns" in d:
        np.testing.assert_allclose(float(env.reward.returns()), d["rv_returns"][S - 1], rtol=1e-5)
        np.testing.assert_allclose(float(env.reward.log_returns()), d["rv_log_returns"][S - 1], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(float(env.reward.sharpe_ratio()), d["rv_sharpe_ratio"][S - 1], rtol=5e-3)
    with pytest.raises(ValueError):
        env.step(torch.zeros(A + 1), feats.clone(), torch.ones(A))                # weight_buffer.py:18-19


def test_compat_three_tuple_variant():
    import pmrl_b200
    from pmrl_b200.compat import TradingEnv
    env = TradingEnv(pmrl_b200.EnvConfig(num_assets=4, window_size=3), three_tuple=True)
    f = torch.zeros(4, 3, 5)
    env.reset(f)
    obs, r, done = env.step(torch.randn(4), f, torch.ones(4))      # agent/dreamer/dreamer.py:190
    assert obs is f and r.shape == () and int(done) == 0 and env.init_cash == 25000.0
