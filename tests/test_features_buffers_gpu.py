"""GPU parity of the feature path, the buffers, the PG reward and the eval metrics (through the C-ABI)."""
import os

import numpy as np
import pytest
import torch

from oracle import ffd_oracle
from oracle.buffers_oracle import RolloutOracle, replay_gather
from tests import util

pytestmark = pytest.mark.gpu


def golden(name):
    z = np.load(os.path.join(util.GOLDEN, name))
    return {k: z[k] for k in z.files}


# ---------------------------------------------------------------- FFD / scaling / packing
def test_ffd_matches_reference_golden():
    from pmrl_b200 import features
    g = golden("ffd.npz")
    w, widths, _ = features.ffd_weights(g["d"], int(g["T"]), float(g["thres"]))
    np.testing.assert_array_equal(widths.cpu().numpy(), g["widths"])                       # integer: bit-exact
    ic = list(g["names"]).index("close")
    wd = int(g["widths"][ic])
    np.testing.assert_array_equal(w[ic, :wd].flip(0).cpu().numpy(), g["weights_close"])    # sequential product, double accumulator like ATen: bit-exact
    out, widths2, mw = features.ffd_transform(g["x"], g["d"], float(g["thres"]))
    assert mw == int(g["max_width"])
    # conv accumulation order differs from torch's CPU conv1d: fp32 tolerance scaled by the series magnitude (~100)
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=1e-5, atol=2e-4)
    out64 = ffd_oracle.ffd_transform_f64(g["x"], g["d"], float(g["thres"]))
    err_gpu = np.abs(out.cpu().numpy() - out64).max(); err_ref = np.abs(g["out"] - out64).max()
    assert err_gpu <= 4 * err_ref + 1e-4, (err_gpu, err_ref)                               # as accurate as the reference's fp32 conv


@pytest.mark.parametrize("d,T", [(0.1, 6000), (0.25, 3000), (0.4, 3000), (0.6, 1500), (0.9, 700), (0.0, 300)])
def test_ffd_vs_oracle_sweep(d, T):
    from pmrl_b200 import features
    rs = np.random.RandomState(int(d * 100) + T)
    x = (100 * np.exp(np.cumsum(0.01 * rs.standard_normal((3, T)), axis=1))).astype(np.float32)
    dd = np.array([d, d, 0.0])
    out, widths, mw = features.ffd_transform(x, dd, 1e-4)
    want, w_o, mw_o = ffd_oracle.ffd_transform(x, dd, 1e-4)
    assert mw == mw_o and list(widths) == list(w_o)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-5, atol=3e-4)
    np.testing.assert_array_equal(out[2].cpu().numpy(), x[2, mw:])                         # d <= 0: passed through untouched


@pytest.mark.parametrize("method", ["minmax", "standard"])
def test_scaling_matches_sklearn(method):
    from pmrl_b200 import features
    rs = np.random.RandomState(1)
    x = (50 + 10 * rs.standard_normal((7, 2111))).astype(np.float32)
    x[3] = 4.25                                                                            # constant series → scale treated as 1
    got = features.scale_series(x, method).cpu().numpy()
    want = ffd_oracle.scale_series(x, method)
    if method == "minmax":
        np.testing.assert_array_equal(got, want)                                           # same fp32 op sequence: bit-exact
    else:
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)


def test_build_env_tables_layout_and_env_consumes_it():
    import pmrl_b200
    from pmrl_b200 import features, synth
    from pmrl_b200.env import BatchedTradingEnv
    T, A, W, L = 1200, 6, 10, 20
    tbl = synth.gbm_ohlc(T, A)
    t = features.build_env_tables(tbl, d=0.6, thres=1e-4, scaler="minmax")
    mw = t["max_width"]
    series = tbl.permute(1, 2, 0).reshape(A * 4, T).numpy()
    want, _, mw_o = ffd_oracle.ffd_transform(series, np.full(A * 4, 0.6), 1e-4)
    assert mw == mw_o
    want = ffd_oracle.scale_series(want, "minmax").reshape(A, 4, T - mw).transpose(0, 2, 1)     # [A, T', C]
    np.testing.assert_allclose(t["feat_am"].cpu().numpy(), want, rtol=1e-4, atol=2e-5)
    np.testing.assert_array_equal(t["close_tm"].cpu().numpy(), tbl[mw:, :, 3].numpy())
    # the env gathers its windows from the packed table
    E = 5
    cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=L)
    env = BatchedTradingEnv.from_tables(cfg, t["close_tm"], t["feat_am"], t0=torch.arange(E, dtype=torch.int32) * 3)
    obs = env.reset()
    for e in range(E):
        assert torch.equal(obs[e, :, :, :4], t["feat_am"][:, 3 * e:3 * e + W, :])


def test_fixedfracdiff_class_surface():
    from pmrl_b200.features import FixedFracDiff
    g = golden("ffd.npz")
    names = [str(n) for n in g["names"]]
    data = {n: torch.from_numpy(g["x"][i]) for i, n in enumerate(names)}
    data["volume"] = torch.arange(int(g["T"]), dtype=torch.float32)
    ffd = FixedFracDiff(data, thres=float(g["thres"]), d_opt={n: float(d) for n, d in zip(names, g["d"])})
    out = ffd.fit_transform()
    assert ffd.get_max_width() == int(g["max_width"])
    for i, n in enumerate(names):
        np.testing.assert_allclose(out[n].numpy(), g["out"][i], rtol=1e-5, atol=2e-4)
    np.testing.assert_array_equal(out["volume"].numpy(), np.arange(int(g["T"]), dtype=np.float32)[int(g["max_width"]):])
    with pytest.raises(NotImplementedError):
        FixedFracDiff(data, d_opt=None).fit()


# ---------------------------------------------------------------- rollout buffer
def test_rollout_buffer_reproduces_reference_batches():
    from pmrl_b200.buffers import DeviceRolloutBuffer
    g = golden("rollout_buffer.npz")
    A, W, F, L, BS = (int(g[k]) for k in ("A", "W", "F", "L", "BS"))
    buf = DeviceRolloutBuffer(F, L, 1, A, W, batch_size=BS)
    buf.set_prices(torch.from_numpy(g["prices"]).t()[W - 1:].reshape(buf.S - 1 + 0, 1, A) if False else
                   torch.cat([torch.from_numpy(g["prices"]).t()[W - 1:], torch.zeros(buf.S - (L - W + 1), A)]).reshape(buf.S, 1, A))
    for step in range(1, L):
        buf.add(torch.from_numpy(g["S"][step]).cuda()[None], torch.from_numpy(g["Act"][step]).cuda().reshape(1, A),
                torch.tensor([g["V"][step]]).cuda(), torch.tensor([g["R"][step]]).cuda())
    np.random.seed(11)
    batches = list(buf.sample_random())
    assert len(batches) == len(g["idxs"])
    for b, t in enumerate(batches):
        for j, name in enumerate(["s", "a", "r", "pv", "pa", "p"]):
            np.testing.assert_array_equal(t[j].cpu().numpy(), g[f"rand{b}_{name}"], err_msg=f"batch {b} {name}")
    seq = list(buf.sample())
    assert len(seq) == int(g["n_seq"])
    for b, t in enumerate(seq):
        for j, name in enumerate(["s", "a", "r", "pv", "pa", "p"]):
            np.testing.assert_array_equal(t[j].cpu().numpy(), g[f"seq{b}_{name}"], err_msg=f"seq {b} {name}")


def test_rollout_buffer_batched_direct_obs_write():
    """The step kernel writes obs straight into the buffer slot; add() appends a / v / r (E > 1)."""
    import pmrl_b200
    from pmrl_b200 import synth
    from pmrl_b200.buffers import DeviceRolloutBuffer
    from pmrl_b200.env import BatchedTradingEnv
    E, A, W, F, L = 9, 7, 6, 5, 24
    tbl = synth.gbm_ohlc(128, A)
    env = BatchedTradingEnv(pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=L), prices=tbl)
    buf = DeviceRolloutBuffer(F, L, E, A, W, batch_size=8)
    ora = [RolloutOracle(F, L, np.ones((A, L)), A, W, batch_size=8) for _ in range(E)]
    s = env.reset().clone()
    g = torch.Generator().manual_seed(2)
    for step in range(1, L):
        a = torch.randn(E, A, generator=g).cuda()
        slot_view = buf.obs_slot()                      # where add() will put s (the obs BEFORE the step, on_policy.py:65)
        if slot_view is not None:
            slot_view.copy_(s)
        s_next, r, _ = env.step(a)
        buf.add(None if slot_view is not None else s, a, env.value, r)
        for e in range(E):
            ora[e].add(s[e].cpu().numpy(), a[e].cpu().numpy(), env.value[e].item(), r[e].item())
        s = s_next.clone()
    slots = np.array([1, 5, 7, 12, 18]); envs = np.array([0, 3, 8, 2, 5])
    got = buf.gather(slots, envs)
    for b in range(len(slots)):
        want = ora[envs[b]].batch(np.array([slots[b]]))
        for j in range(5):                              # s, a, r, _v, _a (prices are unit here)
            np.testing.assert_allclose(got[j][b].cpu().numpy(), want[j][0], rtol=1e-6, atol=0, err_msg=f"b{b} field {j}")


# ---------------------------------------------------------------- index replay
def test_replay_add_and_gather_vs_oracle():
    from pmrl_b200.buffers import DeviceReplayBuffer
    from pmrl_b200 import synth
    E, A, W, F, train_len = 6, 11, 5, 5, 60
    tbl = synth.gbm_ohlc(train_len + W + 2, A)
    feat_am = tbl.permute(1, 0, 2).contiguous().cuda()
    buf = DeviceReplayBuffer(feat_am, F, train_len, E, A, W, buffer_size=3 * (train_len - 2 * (W - 1)), batch_size=8)
    P, L = buf.max_epoch, buf.epoch_len
    bi = np.zeros((P, L, E), np.int32); ba = np.zeros((P, L, E, A), np.float32); br = np.zeros((P, L, E), np.float32)
    rs = np.random.RandomState(0)
    for epoch in range(4):                               # wraps around max_epoch = 3
        for i in range(train_len):
            a = rs.standard_normal((E, A)).astype(np.float32); r = rs.standard_normal(E).astype(np.float32)
            buf.add(epoch, i, torch.from_numpy(a).cuda(), torch.from_numpy(r).cuda())
            st = i - 2 * (W - 1)
            if i >= W - 1 and 0 <= st < L:
                bi[epoch % P, st] = i; ba[epoch % P, st] = a; br[epoch % P, st] = r
    np.testing.assert_array_equal(buf.bi.cpu().numpy(), bi)
    np.testing.assert_array_equal(buf.ba.cpu().numpy(), ba)
    assert len(buf) == P
    epochs = np.array([0, 1, 2, 2, 0]); envs = np.array([0, 5, 2, 3, 1]); starts = np.array([0, 3, L - W - 2, 10, 7])
    s, a, r, s2 = buf.gather(epochs, envs, starts)
    for b in range(len(epochs)):
        ws, wa, wr, ws2 = replay_gather(tbl.numpy(), bi, ba, br, epochs[b], envs[b], starts[b], W)
        np.testing.assert_array_equal(s[b].cpu().numpy(), ws); np.testing.assert_array_equal(s2[b].cpu().numpy(), ws2)
        np.testing.assert_array_equal(a[b].cpu().numpy(), wa); np.testing.assert_array_equal(r[b].cpu().numpy(), wr)
    with pytest.raises(IndexError):                      # 3 stored epochs cannot fill 8 rows (traj_buffer.py:56-57,67)
        buf.sample(torch.Generator().manual_seed(0))
    s, a, r, s2 = buf.sample(torch.Generator().manual_seed(0), sampler="buffer")
    assert s.shape == (8, A, W, F) and a.shape == (8, A, 1) and r.shape == (8, 1, 1) and s2.shape == (8, A, W, F)


@pytest.mark.parametrize("tag,sampler", [("buf", "buffer"), ("traj", "traj")])
def test_replay_buffer_matches_reference_golden(tag, sampler):
    """pmrl_replay_add / pmrl_replay_gather and the two samplers against replay/buffer.py and replay/traj_buffer.py run
    unmodified (tests/golden/make_golden_replay.py): stored rows and sampled (s, a, r, s') bit-exact at E = 1."""
    from pmrl_b200.buffers import DeviceReplayBuffer
    g = golden("replay_buffer.npz")
    A, W, F, TL, B = (int(g[k]) for k in ("A", "W", "F", "train_len", "batch"))
    L = TL - 2 * (W - 1)
    feat_am = torch.from_numpy(g["table"]).permute(1, 0, 2).contiguous().cuda()
    buf = DeviceReplayBuffer(feat_am, F, TL, 1, A, W, buffer_size=int(g["epochs_kept"]) * L, batch_size=B)
    for e in range(int(g["n_epochs"])):
        for step in range(1, TL):
            buf.add(e, step, torch.from_numpy(g["acts"][e, step]).cuda(), torch.from_numpy(g["rews"][e, step:step + 1]).cuda())
    np.testing.assert_array_equal(buf.bi.cpu().numpy()[..., 0], g[f"{tag}_i"][..., 0].astype(np.int32))
    np.testing.assert_array_equal(buf.ba.cpu().numpy()[:, :, 0], g[f"{tag}_a"])
    np.testing.assert_array_equal(buf.br.cpu().numpy()[..., 0], g[f"{tag}_r"][..., 0, 0])
    for k in range(3):
        torch.manual_seed(100 + k)
        s, a, r, s2 = buf.sample(sampler=sampler, percent_latest=float(g["percent_latest"]))
        np.testing.assert_array_equal(s.cpu().numpy(), g[f"{tag}{k}_s"]); np.testing.assert_array_equal(s2.cpu().numpy(), g[f"{tag}{k}_s2"])
        np.testing.assert_array_equal(a.cpu().numpy(), g[f"{tag}{k}_a"]); np.testing.assert_array_equal(r.cpu().numpy(), g[f"{tag}{k}_r"])


# ---------------------------------------------------------------- PG reward (fwd + grad) and eval metrics
@pytest.mark.parametrize("mode", ["log_returns", "returns", "sharpe_ratio"])
def test_pg_reward_matches_reference_autograd(mode):
    from pmrl_b200.pg_reward import pg_reward
    g = golden("pg_reward.npz")
    a = torch.from_numpy(g[f"{mode}_a"]).cuda(); pv = torch.from_numpy(g[f"{mode}_pv"]).cuda()
    pa = torch.from_numpy(g[f"{mode}_pa"]).cuda(); p = torch.from_numpy(g[f"{mode}_p"]).cuda()
    rew, grad = pg_reward(a, pv, pa, p, mode=mode, normalise=True)
    want_g = g[f"{mode}_grad"].reshape(grad.shape)
    if mode == "sharpe_ratio":
        # mean(ret) / std(ret) with std ≈ 3e-3 of values ≈ 1: fp32 rounding of the returns (6e-8) is amplified by 1 / std in the
        # ratio and by 1 / std^2 in its gradient — the bounds below are that amplification, not slack in the kernel
        np.testing.assert_allclose(rew.mean().item(), float(g[f"{mode}_r"]), rtol=2e-4)
        np.testing.assert_allclose(grad.cpu().numpy(), want_g, rtol=5e-3, atol=2e-3 * np.abs(want_g).max())
        return
    np.testing.assert_allclose(rew.mean().item(), float(g[f"{mode}_r"]), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(grad.cpu().numpy(), want_g, rtol=2e-4, atol=2e-8)


def test_pg_reward_autograd_function():
    from pmrl_b200.pg_reward import PGReward
    B, A = 32, 50
    a = torch.randn(B, A, 1, device="cuda", requires_grad=True)
    pv = torch.full((B, 1, 1), 25000.0, device="cuda"); pa = torch.softmax(torch.randn(B, A, 1, device="cuda"), 1)
    p = 1 + 0.01 * torch.randn(B, A, 1, device="cuda")
    r = PGReward.apply(a, pv, pa, p, "log_returns", True, 0.0, 1.0)
    (-r).backward()
    a2 = a.detach().clone().requires_grad_(True)
    w = torch.softmax(a2, dim=1)
    ref = torch.log((pv * (w * p)).sum(1, keepdim=True) / pv).mean()
    (-ref).backward()
    np.testing.assert_allclose(r.item(), ref.item(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(a.grad.cpu().numpy(), a2.grad.cpu().numpy(), rtol=2e-4, atol=2e-8)


def test_eval_metrics_vs_numpy():
    from pmrl_b200.metrics import eval_metrics
    E, N, A = 17, 333, 9
    rs = np.random.RandomState(2)
    vals = (25000 * np.exp(np.cumsum(0.01 * rs.standard_normal((E, N)), axis=1))).astype(np.float32)
    w = rs.random_sample((E, N, A)).astype(np.float32); w /= w.sum(-1, keepdims=True)
    got = eval_metrics(torch.from_numpy(vals).cuda(), torch.from_numpy(w).cuda(), rf=0.04, periods=252).cpu().numpy()
    v = vals.astype(np.float64)
    r = v[:, 1:] / v[:, :-1] - 1 - ((1.04) ** (1 / 252) - 1)
    sharpe = r.mean(1) / r.std(1, ddof=1) * np.sqrt(252)
    sortino = r.mean(1) / np.sqrt((np.minimum(r, 0) ** 2).sum(1) / r.shape[1]) * np.sqrt(252)
    mdd = (v / np.maximum.accumulate(v, axis=1)).min(1) - 1
    to = np.abs(np.diff(w.astype(np.float64), axis=1)).sum(-1).mean(1)        # util/eval.py:32-37
    np.testing.assert_allclose(got[:, 0], sharpe, rtol=1e-4); np.testing.assert_allclose(got[:, 1], sortino, rtol=1e-4)
    np.testing.assert_allclose(got[:, 2], mdd, rtol=1e-5, atol=1e-6); np.testing.assert_allclose(got[:, 3], to, rtol=1e-5)


# ---------------------------------------------------------------- collect loops (N1)
def test_collect_loops_drive_env_and_buffers():
    import pmrl_b200
    from pmrl_b200 import loops, synth
    from pmrl_b200.buffers import DeviceReplayBuffer, DeviceRolloutBuffer
    from pmrl_b200.env import BatchedTradingEnv
    from oracle.env_oracle import OracleEnv
    E, A, W, F, items = 6, 9, 5, 5, 30
    tbl = synth.gbm_ohlc(128, A)
    cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=items)
    env = BatchedTradingEnv(cfg, prices=tbl)
    torch.manual_seed(0)
    proj = torch.randn(W * F, device="cuda") * 2e-3                          # prices ~1e2 → raw scores of order 1
    act_fn = lambda s: (s.reshape(E, A, W * F) @ proj)                       # a deterministic "policy": [E, A] raw scores
    buf = DeviceRolloutBuffer(F, items, E, A, W, batch_size=4)
    loops.collect_on_policy(env, act_fn, buf, items)
    # replay the same loop against the oracle with the actions the policy produced from the oracle's own obs
    ora = OracleEnv(E, A, W, F, close=tbl[:, :, 3].numpy(), feat=tbl.numpy(), episode_len=items)
    s = ora.obs()
    for step in range(1, items):
        a = act_fn(torch.from_numpy(s).cuda()).cpu().numpy()
        slot = step - (W - 1)
        if slot > 0:
            np.testing.assert_allclose(buf.s[slot].cpu().numpy(), s, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(buf.a[slot].cpu().numpy(), a, rtol=1e-4, atol=1e-6)
        r, _ = ora.step(a)
        if slot > 0:
            np.testing.assert_allclose(buf.r[slot].cpu().numpy(), r, rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(buf.v[slot].cpu().numpy(), ora.value, rtol=1e-5)
        s = ora.obs()
    rb = DeviceReplayBuffer(env.feat_am, F, items, E, A, W, buffer_size=3 * items, batch_size=4)
    loops.collect_off_policy(env, act_fn, rb, 0, items)
    assert int(rb.bi[0, 0, 0]) == 2 * (W - 1) and int(rb.bi[0, -1, 0]) == items - 1        # t0 = 0: loader index = lockstep counter
    total, met = loops.evaluate(env, act_fn, items)
    assert total.shape == (E,) and met.shape == (E, 4) and torch.isfinite(met).all() and torch.isfinite(total).all()
    assert np.isfinite(ora.value).all()


def test_off_policy_collect_stores_each_envs_own_loader_index():
    """Per-env episode offsets: the replay row index is t0[e] + k, so a sampled (s, a, r, s') regenerates exactly the
    observations that env saw (feature channels) around the stored action — not rows of the lockstep counter."""
    import pmrl_b200
    from pmrl_b200 import loops, synth
    from pmrl_b200.buffers import DeviceReplayBuffer
    from pmrl_b200.env import BatchedTradingEnv
    E, A, W, F, items = 7, 12, 5, 5, 40
    tbl = synth.gbm_ohlc(256, A)
    t0 = torch.tensor([0, 3, 17, 101, 55, 9, 200], dtype=torch.int32)
    env = BatchedTradingEnv(pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=items), prices=tbl, t0=t0)
    seen, acts = [], []
    gen = torch.Generator(device="cuda").manual_seed(3)

    def act_fn(obs):
        seen.append(obs.clone())                                  # seen[k] = obs after step k (k = 0: reset)
        acts.append(torch.randn(E, A, generator=gen, device="cuda"))
        return acts[-1]

    rb = DeviceReplayBuffer(env.feat_am, F, items, E, A, W, buffer_size=2 * items, batch_size=4)
    last = loops.collect_off_policy(env, act_fn, rb, 0, items)
    seen.append(last.clone())
    off = 2 * (W - 1)
    steps = torch.arange(off, items, dtype=torch.int32)
    assert torch.equal(rb.bi[0].cpu(), steps[:, None] + t0[None, :])
    epochs = np.zeros(5, np.int64); envs = np.array([1, 3, 6, 2, 4]); starts = np.array([0, 4, 11, 20, rb.epoch_len - W - 2])
    s, a, r, s2 = rb.gather(epochs, envs, starts)
    for b in range(5):
        n = int(starts[b]) + W - 1 + off                          # loader step whose row is `end - 1`
        assert torch.equal(s[b, :, :, :F - 1], seen[n][envs[b], :, :, :F - 1]), f"s of sample {b}"
        assert torch.equal(s2[b, :, :, :F - 1], seen[n + 1][envs[b], :, :, :F - 1]), f"s' of sample {b}"
        assert torch.equal(a[b, :, 0], acts[n][envs[b]])         # a[:, -1] = action stored at `end` = taken at loader step n + 1
        # channel F-1 = the action history ending at that action (buffer.py:69-70)
        hist = torch.stack([acts[m - 1][envs[b]] for m in range(n - W + 1, n + 2)], dim=1)      # [A, W + 1]
        assert torch.equal(s[b, :, :, F - 1], hist[:, :-1]) and torch.equal(s2[b, :, :, F - 1], hist[:, 1:])


@pytest.mark.parametrize("A,W,E", [(12, 5, 7), (100, 50, 5), (50, 8, 9)])
def test_index_mode_rollout_buffer_regenerates_the_stored_observations(A, W, E):
    """mode='index' (loader index + raw action + un-wrapped weight history, 8A + 12 bytes per env-step) gathers exactly the
    minibatch the full-storage buffer (4·A·W·F bytes per env-step, the reference layout) holds — both filled by the step
    kernel itself through `collect_on_policy`."""
    import pmrl_b200
    from pmrl_b200 import loops, synth
    from pmrl_b200.buffers import DeviceRolloutBuffer
    from pmrl_b200.env import BatchedTradingEnv
    F, items = 5, W + 23
    tbl = synth.gbm_ohlc(items + W + 80, A)
    t0 = synth.episode_offsets(E, items + W + 80, W, items)
    cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=items)
    gens = [torch.Generator(device="cuda").manual_seed(8) for _ in range(2)]
    bufs = []
    for mode, gen in zip(("full", "index"), gens):
        env = BatchedTradingEnv(cfg, prices=tbl, t0=t0)
        buf = DeviceRolloutBuffer(F, items, E, A, W, batch_size=4, mode=mode, feat_am=env.feat_am, y_tm=env.y_tm)
        c0 = env.lib.pmrl_launch_count()
        loops.collect_on_policy(env, lambda s, gen=gen: torch.randn(E, A, generator=gen, device="cuda"), buf, items)
        # one fused launch per item (+ the reset's two): no copy kernels in the loop
        assert env.lib.pmrl_launch_count() - c0 == (items - 1) + 2
        if mode == "full":
            buf.fill_prices(env)
        bufs.append(buf)
    full, index = bufs
    assert index.nbytes() * 5 < full.nbytes()
    assert torch.equal(full.a, index.a) and torch.equal(full.v, index.v) and torch.equal(full.r, index.r)
    rs = np.random.RandomState(0)
    # the reference samples idx in [1, epoch_len) (rollout_buffer.py:123): the last slot of the epoch buffer is never filled
    slots = rs.randint(1, full.S - 1, 24); envs = rs.randint(0, E, 24)
    slots[:3] = (1, 2, full.S - 2)
    for got, want, name in zip(index.gather(slots, envs), full.gather(slots, envs), ("s", "a", "r", "pv", "pa", "p")):
        assert torch.equal(got, want), name
    with pytest.raises(IndexError):
        full.gather(np.array([0]), np.array([0]))                 # slot - 1 would be read (rollout_buffer.py:130-131)


# ---------------------------------------------------------------- indicator windows (N3)
def test_indicator_windows_vs_oracle():
    from oracle import indicators_oracle as io
    from pmrl_b200 import features, synth
    T, A = 600, 7
    tbl = synth.gbm_ohlc(T, A, seed=9)
    inds = [("ema", {"timeperiod": 30}), ("ema", {"timeperiod": 60}), ("bbands", {"timeperiod": 20}), ("macd", {}),
            ("atr", {"timeperiod": 14}), ("rsi", {"timeperiod": 30}), ("sma", {"timeperiod": 10})]
    names, out, lb = features.add_indicators(tbl, inds)
    assert lb == 59 and out.shape == (A, 11, T - lb)
    assert names[:2] == ["ema_30", "ema_60"] and names[2:5] == ["upperband_20", "middleband_20", "lowerband_20"]
    got = out.cpu().numpy()
    o, h, l, c = (tbl[:, :, i].numpy().T for i in range(4))
    for a in range(A):
        want = [io.ema(c[a], 30), io.ema(c[a], 60), *io.bbands(c[a], 20), *io.macd(c[a]), io.atr(h[a], l[a], c[a], 14),
                io.rsi(c[a], 30), io.sma(c[a], 10)]
        for j, w in enumerate(want):
            np.testing.assert_allclose(got[a, j], w[lb:], rtol=2e-6, atol=1e-6, err_msg=f"asset {a} output {names[j]}")
    assert np.isfinite(got).all()                              # nothing of the lookback survives the clip
    # analytic checks: cash (asset 0, constant price): EMA = price, bands collapse, RSI = 0 (no gains), ATR = 0
    np.testing.assert_allclose(got[0, 0], 1.0); np.testing.assert_allclose(got[0, 2], got[0, 4]); assert (got[0, 9] == 0).all()
    with pytest.raises(NotImplementedError):
        features.add_indicators(tbl, {"sar": {}})


def test_build_env_tables_with_indicators_feeds_the_fused_kernel():
    """N3 → a10 → a11 → table: indicator windows appended as features (config/base.py:30-44), FFD'ed, scaled and packed to an
    F = 9 table (OHLC + ema + bbands + weight slot) that the fused step+obs kernel consumes; every stage against its oracle."""
    import pmrl_b200
    from oracle import indicators_oracle as io
    from oracle.env_oracle import OracleEnv
    from pmrl_b200 import features, synth, _lib
    from pmrl_b200.env import BatchedTradingEnv
    T, A, W, L, E = 900, 36, 12, 30, 40
    tbl = synth.gbm_ohlc(T, A, seed=5)
    inds = [("ema", {"timeperiod": 30}), ("bbands", {"timeperiod": 20})]
    t = features.build_env_tables(tbl, d=0.6, thres=1e-4, scaler="minmax", indicators=inds)
    lb, mw = t["lookback"], t["max_width"]
    assert lb == 29 and t["names"] == ["open", "high", "low", "close", "ema_30", "upperband_20", "middleband_20", "lowerband_20"]
    assert t["feat_am"].shape == (A, T - lb - mw, 8) and t["rows"] == T - lb - mw
    c = tbl[:, :, 3].numpy().T
    series = np.empty((A, 8, T - lb), np.float32)
    series[:, :4] = tbl.permute(1, 2, 0).numpy()[:, :, lb:]
    for a in range(A):
        series[a, 4] = io.ema(c[a], 30)[lb:]
        series[a, 5], series[a, 6], series[a, 7] = (x[lb:] for x in io.bbands(c[a], 20))
    want, _, mw_o = ffd_oracle.ffd_transform(series.reshape(A * 8, T - lb), np.full(A * 8, 0.6), 1e-4)
    assert mw == mw_o
    want = ffd_oracle.scale_series(want, "minmax").reshape(A, 8, T - lb - mw).transpose(0, 2, 1)
    np.testing.assert_allclose(t["feat_am"].cpu().numpy(), want, rtol=1e-4, atol=5e-5)
    np.testing.assert_array_equal(t["close_tm"].cpu().numpy(), tbl[lb + mw:, :, 3].numpy())
    # the env steps on the packed F = 9 table through ONE fused launch per step
    rows = t["rows"]
    t0 = synth.episode_offsets(E, rows, W, L)
    cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, num_features=9, episode_len=L)
    env = BatchedTradingEnv.from_tables(cfg, t["close_tm"], t["feat_am"], t0=t0)
    feat_tm = t["feat_am"].permute(1, 0, 2).contiguous().cpu().numpy()
    ora = OracleEnv(E, A, W, 9, close=t["close_tm"].cpu().numpy(), feat=feat_tm, t0=t0.numpy(), episode_len=L)
    compare = lambda obs, msg: (np.testing.assert_array_equal(obs.cpu().numpy()[..., :8], ora.obs()[..., :8], err_msg=msg),
                                np.testing.assert_allclose(obs.cpu().numpy()[..., 8], ora.obs()[..., 8], rtol=1e-5, atol=1e-6, err_msg=msg))
    compare(env.reset(), "reset")
    g = torch.Generator().manual_seed(3)
    n0 = _lib.load().pmrl_launch_count()
    for s in range(L + 4):
        act = torch.randn(E, A, generator=g)
        obs, r, done = env.step(act.cuda())
        r_o, d_o = ora.step(act.numpy())
        np.testing.assert_array_equal(done.cpu().numpy(), d_o)
        np.testing.assert_allclose(r.cpu().numpy(), r_o, rtol=1e-5, atol=1e-6)
        compare(obs, f"step {s}")
    assert _lib.load().pmrl_launch_count() - n0 == L + 4              # fused: one kernel per step


def test_directional_movement_indicators_vs_oracle():
    """adx / dx of the reference's commented default set (config/base.py:38-39)."""
    from oracle import indicators_oracle as io
    from pmrl_b200 import features, synth
    T, A = 300, 6
    tbl = synth.gbm_ohlc(T, A, seed=21)
    names, out, lb = features.add_indicators(tbl, [("adx", {"timeperiod": 30}), ("dx", {"timeperiod": 30}), ("adx", {"timeperiod": 5})])
    assert lb == 59 and out.shape == (A, 3, T - lb) and names == ["adx_30", "dx_30", "adx_5"]
    got = out.cpu().numpy()
    o, h, l, c = (tbl[:, :, i].numpy().T for i in range(4))
    for a in range(A):
        want = [io.adx(h[a], l[a], c[a], 30), io.dx(h[a], l[a], c[a], 30), io.adx(h[a], l[a], c[a], 5)]
        for j, w in enumerate(want):
            np.testing.assert_allclose(got[a, j], w[lb:], rtol=2e-6, atol=1e-5, err_msg=f"asset {a} {names[j]}")
    assert np.isfinite(got).all() and (got >= 0).all() and (got <= 100.0 + 1e-4).all()
    assert (got[0] == 0).all()                                   # cash: no range, no movement → DX = ADX = 0


def test_volume_and_oscillator_indicators_vs_oracle():
    from oracle import indicators_oracle as io
    from pmrl_b200 import features, synth
    T, A = 400, 5
    tbl4 = synth.gbm_ohlc(T, A, seed=11)
    vol = (1e5 * (1 + torch.rand(T, A, generator=torch.Generator().manual_seed(1)))).unsqueeze(-1)
    tbl = torch.cat([tbl4, vol], dim=-1)                            # o, h, l, c, v
    names, out, lb = features.add_indicators(tbl, [("obv", {}), ("adosc", {}), ("cci", {"timeperiod": 14}), ("stoch", {})])
    assert lb == 13 and out.shape == (A, 5, T - lb) and names[3:] == ["slowk_", "slowd_"]
    got = out.cpu().numpy()
    o, h, l, c, v = (tbl[:, :, i].numpy().T for i in range(5))
    for a in range(1, A):
        want = [io.obv(c[a], v[a]), io.adosc(h[a], l[a], c[a], v[a]), io.cci(h[a], l[a], c[a], 14), *io.stoch(h[a], l[a], c[a])]
        for j, w in enumerate(want):
            np.testing.assert_allclose(got[a, j], w[lb:], rtol=3e-5, atol=1e-3 * max(1.0, float(np.nanmax(np.abs(w[lb:])))) * 1e-3,
                                       err_msg=f"asset {a} {names[j]}")
    with pytest.raises(Exception):
        features.add_indicators(tbl4, [("obv", {})])               # needs the volume channel
