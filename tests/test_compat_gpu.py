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


# ---------------------------------------------------------------- the reference's own loop with its own agent
class _RecordedPG:
    """Stands in for the live `agent.pg.pg.PG` on a box without the reference tree: `act` asserts it is shown the
    observation the live policy was shown at this step (feature window bit-exact, weight channel ≤ 1e-5) and returns the
    action the live policy produced from it (tests/golden/make_golden_pg_rollout.py)."""

    def __init__(self, g, device):
        self.g, self.device, self.step = g, device, 1
        self.table = torch.from_numpy(g["table"])

    def training_mode(self, mode):
        pass

    def act(self, s):
        g, k = self.g, self.step
        A, W, F = int(g["A"]), int(g["W"]), int(g["F"])
        assert s.device.type == self.device and tuple(s.shape) == (A, W, F)
        got = s.cpu().numpy()
        np.testing.assert_array_equal(got[:, :, :F - 1], self.table[k - 1:k - 1 + W].permute(1, 0, 2).numpy())
        np.testing.assert_allclose(got[:, :, -1], g["obs_w"][k], rtol=1e-5, atol=1e-6, err_msg=f"obs weight channel, step {k}")
        self.step += 1
        return torch.from_numpy(g["actions"][k]).reshape(1, A, 1).to(self.device)


@pytest.mark.parametrize("name", ["pg_rollout_A11_W50_softmax.npz", "pg_rollout_A11_W16_softmax.npz", "pg_rollout_A11_W16_init.npz"])
@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_on_policy_rollout_body_with_the_reference_agent(name, device):
    """BASELINE config 1: `Train._rollout` (train/on_policy.py:56-67) driven unchanged — the loop body below is the
    reference's, with compat.TradingEnv swapped in for env.sim.TradingEnv.  Golden = the same loop run on the live
    reference with the live PG policy and the live RolloutBuffer."""
    import os
    import pmrl_b200
    from pmrl_b200.buffers import DeviceRolloutBuffer
    from pmrl_b200.compat import TradingEnv
    from oracle.buffers_oracle import RolloutOracle
    z = np.load(os.path.join(util.GOLDEN, name)); g = {k: z[k] for k in z.files}
    A, W, F, L, BS = (int(g[k]) for k in ("A", "W", "F", "L", "batch"))
    table, y = torch.from_numpy(g["table"]), torch.from_numpy(g["y"])
    train_dl = []
    for i in range(L):                                              # the toy loader of the generator
        data = torch.zeros(A, W, F)
        data[:, :, :F - 1] = table[i:i + W].permute(1, 0, 2)
        train_dl.append((i, y[i].clone().to(device), data.to(device)))

    class Self:
        pass
    self = Self()
    self.env = TradingEnv(pmrl_b200.EnvConfig(num_assets=A, window_size=W, num_features=F))
    self.agent = _RecordedPG(g, device)
    if device == "cuda":
        self.buffer = DeviceRolloutBuffer(F, L, 1, A, W, batch_size=BS)
        self.buffer.set_prices(torch.cat([y[W - 1:], torch.zeros(self.buffer.S - (L - W + 1), A)]).reshape(self.buffer.S, 1, A))
    else:
        self.buffer = RolloutOracle(F, L, g["y"].T, A, W, batch_size=BS)
    values, rewards = np.zeros(L, np.float32), np.zeros(L, np.float32)
    # ---- train/on_policy.py:56-67 ----
    self.agent.training_mode(False)
    self.buffer.reset()
    for step, (datetime, prices, data) in enumerate(train_dl):
        if step == 0:
            self.s = self.env.reset(data)
        else:
            a = self.agent.act(self.s)
            r, s_ = self.env.step(a, data, prices)
            self.buffer.add(self.s, a, self.env.value, r)
            self.s = s_
            values[step], rewards[step] = float(self.env.value), float(r)
    # ----------------------------------
    util.assert_values_close(values[1:], g["values"][1:], name)
    util.assert_rewards_close(rewards[1:], g["rewards"][1:], name)
    assert self.env.weights.idx == int(g["ring_idx"]) and self.env.weights.is_full == bool(g["ring_full"])
    b = self.buffer
    if device == "cuda":
        ba, bv, br = b.a.cpu().numpy().reshape(-1, A, 1), b.v.cpu().numpy().reshape(-1, 1, 1), b.r.cpu().numpy().reshape(-1, 1, 1)
        batch = [t.cpu().numpy() for t in b.gather(g["idxs"][0], np.zeros(BS, np.int64))]
    else:
        ba, bv, br = b.a, b.v, b.r
        batch = b.batch(g["idxs"][0])
    np.testing.assert_array_equal(ba.astype(np.float32), g["buf_a"])                   # raw actions are stored as given
    util.assert_values_close(bv[:-1], g["buf_v"][:-1], name)
    util.assert_rewards_close(br, g["buf_r"], name)
    for j, field in enumerate(["s", "a", "r", "pv", "pa", "p"]):    # first minibatch of sample_random (rollout_buffer.py:125-140)
        np.testing.assert_allclose(batch[j], g[f"rand0_{field}"], rtol=1e-5, atol=1e-6, err_msg=field)
