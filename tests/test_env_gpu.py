"""GPU parity of the CUDA env path (through the C-ABI) against the golden fixtures and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle.env_oracle import OracleEnv
from tests import util

pytestmark = pytest.mark.gpu

REWARD_IDS = {"step_log": 0, "returns": 1, "log_returns": 2, "sharpe_ratio": 3}


def _mods():
    import pmrl_b200
    from pmrl_b200 import synth
    from pmrl_b200.env import BatchedTradingEnv
    return pmrl_b200, synth, BatchedTradingEnv


def make_pair(E, A, W, F=5, T=None, episode_len=60, commission=0.0, reward="step_log", seed=1234,
              strict=True, collect_stats=False, first_env=0):
    pmrl, synth, Env = _mods()
    T = T or (W + episode_len + 64)
    tbl4 = synth.gbm_ohlc(T, A, seed)                       # [T, A, 4] (CPU bits shared by both sides)
    if F - 1 == 4:
        feat = tbl4
    else:                                                   # other channel counts: deterministic mix of OHLC
        reps = [tbl4[..., i % 4] * (1.0 + 0.25 * (i // 4)) for i in range(F - 1)]
        feat = torch.stack(reps, dim=-1).contiguous()
    t0 = synth.episode_offsets(E, T, W, episode_len, first_env=first_env)
    cfg = pmrl.EnvConfig(num_envs=E, num_assets=A, window_size=W, num_features=F, commission=commission,
                         reward=reward, episode_len=episode_len, strict_reference=strict)
    gpu = Env(cfg, prices=tbl4, features=feat, t0=t0, collect_stats=collect_stats)
    ora = OracleEnv(E, A, W, F, close=tbl4[:, :, 3].numpy(), feat=feat.numpy(), t0=t0.numpy(),
                    episode_len=episode_len, commission=commission, reward_mode=REWARD_IDS[reward],
                    strict_reference=strict)
    return gpu, ora


def compare_state(gpu, ora, msg=""):
    np.testing.assert_array_equal(gpu.idx.cpu().numpy(), ora.idx, err_msg=msg)
    np.testing.assert_array_equal(gpu.is_full.cpu().numpy(), ora.is_full, err_msg=msg)
    np.testing.assert_array_equal(gpu.t.cpu().numpy(), ora.t, err_msg=msg)
    util.assert_values_close(gpu.value.cpu().numpy(), ora.value, msg)
    np.testing.assert_allclose(gpu.hist.cpu().numpy(), ora.hist, rtol=util.RTOL_WEIGHT, atol=util.ATOL_WEIGHT, err_msg=msg)


def compare_obs(obs_gpu, ora, msg=""):
    want = ora.obs()
    got = obs_gpu.cpu().numpy()
    F = want.shape[-1]
    np.testing.assert_array_equal(got[..., : F - 1], want[..., : F - 1], err_msg=msg + " feature window")   # pure gather: bit-exact
    np.testing.assert_array_equal(got[..., F - 1] == 0, want[..., F - 1] == 0, err_msg=msg + " weight-channel padding")
    np.testing.assert_allclose(got[..., F - 1], want[..., F - 1], rtol=util.RTOL_WEIGHT, atol=util.ATOL_WEIGHT, err_msg=msg)


@pytest.mark.parametrize("name", util.env_fixture_names())
def test_cuda_replays_reference_rollout(name):
    """The CUDA path on the exact inputs of the live-reference rollouts (E = 1, external price relatives)."""
    pmrl, synth, Env = _mods()
    d = util.load_env_fixture(name)
    A, W, F, S = d["A"], d["W"], d["F"], d["S"]
    cfg = pmrl.EnvConfig(num_envs=1, num_assets=A, window_size=W, num_features=F, commission=d["commission"],
                         episode_len=0)
    env = Env(cfg)
    act = torch.from_numpy(d["actions"]).cuda()
    y = torch.from_numpy(d["y"]).cuda()
    vals = torch.zeros(S, device="cuda"); rews = torch.zeros(S, device="cuda")
    idx = torch.zeros(S, dtype=torch.int32, device="cuda"); full = torch.zeros(S, dtype=torch.uint8, device="cuda")
    wsteps = util.weight_steps(d); snaps = [int(s) for s in d["obs_w_steps"]]
    wts, obs_w = [], []
    scratch = torch.zeros(1, A, W, F, device="cuda")
    reset_w = env.write_weight_channel(scratch.clone())[0, :, :, -1].cpu().numpy()
    for s in range(S):
        _, r, _ = env.step(act[s:s + 1], y=y[s:s + 1], obs=False)
        vals[s] = env.value[0]; rews[s] = r[0]; idx[s] = env.idx[0]; full[s] = env.is_full[0]
        if s in wsteps:
            wts.append(env.weights_last[0].clone())
        if s in snaps:
            obs_w.append(env.write_weight_channel(scratch.clone())[0, :, :, -1])
    np.testing.assert_array_equal(idx.cpu().numpy(), d["idx"])
    np.testing.assert_array_equal(full.cpu().numpy(), d["is_full"])
    util.assert_values_close(vals.cpu().numpy(), d["values"], name)
    util.assert_rewards_close(rews.cpu().numpy(), d["rewards"], name)
    np.testing.assert_allclose(torch.stack(wts).cpu().numpy(), d["weights"], rtol=util.RTOL_WEIGHT, atol=util.ATOL_WEIGHT)
    got_w = torch.stack(obs_w).cpu().numpy()
    np.testing.assert_array_equal(got_w == 0, d["obs_w"] == 0)
    np.testing.assert_allclose(got_w, d["obs_w"], rtol=util.RTOL_WEIGHT, atol=util.ATOL_WEIGHT)
    np.testing.assert_array_equal(reset_w, d["reset_obs_w"])


@pytest.fixture
def tuning():
    from pmrl_b200 import _lib
    yield _lib
    for k in (_lib.TUNE_GROUP_ENVS, _lib.TUNE_CTAS_PER_SM):
        _lib.set_tuning(k, 0)
    for k in (_lib.TUNE_FUSED, _lib.TUNE_RING_TMA, _lib.TUNE_FAST_FILL, _lib.TUNE_STAGED):
        _lib.set_tuning(k, 1)
    _lib.set_tuning(_lib.TUNE_HOST_STREAM, 0)


# (fused, group_envs, ring_tma, fast_fill): the fused step+obs kernel in several launch shapes — ring-through-TMA kernel (default, with
# forced group sizes incl. the 16-env narrow groups, and without the next-group L2 prefetch), register-staged kernel — and the
# two-kernel path with the division-free and with the generic obs tile kernel
VARIANTS = [(1, 0, 1, 1), (1, 3, 1, 1), (1, 5, 1, 1), (1, 7, 1, 1), (1, 12, 1, 1), (1, 16, 1, 1), (1, 0, 2, 1),
            (1, 0, 0, 1), (1, 3, 0, 1), (0, 0, 1, 1), (0, 0, 1, 0)]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("A,W,F,E", [(11, 50, 5, 48), (50, 50, 5, 48), (100, 50, 5, 45), (64, 24, 5, 37), (500, 50, 5, 21), (7, 5, 3, 50),
                                     (33, 9, 6, 48), (130, 16, 2, 40), (1, 4, 5, 9), (3, 300, 5, 5),
                                     # wider feature sets (OHLC + indicator outputs): fused kernel at F = 9 / 13 / 17, tile kernels else
                                     (100, 50, 9, 45), (50, 50, 9, 37), (64, 24, 13, 37), (36, 20, 17, 30), (100, 50, 7, 20),
                                     (40, 16, 12, 33), (34, 10, 3, 41),
                                     # every chunk count of the channel-padded fused path (PmrlTables.feat_am4)
                                     (64, 12, 2, 30), (36, 20, 6, 30), (100, 24, 8, 20), (52, 10, 10, 25), (44, 9, 14, 21), (32, 8, 16, 19)])
def test_table_driven_step_and_obs_vs_oracle(A, W, F, E, variant, tuning):
    """Batched, table-driven: window gather, y from the close plane, ring wrap, done and auto-reset."""
    fused, group, rt, fast = variant
    tuning.set_tuning(tuning.TUNE_FUSED, fused)
    tuning.set_tuning(tuning.TUNE_GROUP_ENVS, group)
    tuning.set_tuning(tuning.TUNE_RING_TMA, rt)
    tuning.set_tuning(tuning.TUNE_FAST_FILL, fast)
    L = W + 7
    gpu, ora = make_pair(E, A, W, F, episode_len=L)
    g = torch.Generator().manual_seed(99)
    obs = gpu.reset()
    compare_obs(obs, ora, "reset")
    for s in range(2 * L + 5):
        act = torch.randn(E, A, generator=g)
        if s % 7 == 3:
            act = torch.softmax(act, dim=1)                    # simplex rows → pass-through branch
        obs, r, done = gpu.step(act.cuda())
        r_o, d_o = ora.step(act.numpy())
        np.testing.assert_array_equal(done.cpu().numpy(), d_o, err_msg=f"done @ {s}")
        util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
        compare_state(gpu, ora, f"step {s}")
        if s < 4 or s % 5 == 0 or (W - 3 <= s % (L + 1) <= W + 2) or s % (L + 1) in (L - 1, L, 0):
            compare_obs(obs, ora, f"obs @ {s}")


@pytest.mark.parametrize("E", [300, 2048, 2500])
def test_midsize_batches_use_wave_sized_groups(E):
    """Batches of a few CTA rounds: the RT launcher sizes its env groups to fill whole rounds (2, 7, 8 envs here) and the
    last group is ragged."""
    A, W, L = 40, 8, 12
    gpu, ora = make_pair(E, A, W, 5, episode_len=L)
    g = torch.Generator().manual_seed(7)
    obs = gpu.reset()
    for s in range(L + 3):
        act = torch.randn(E, A, generator=g)
        obs, r, done = gpu.step(act.cuda())
        r_o, d_o = ora.step(act.numpy())
        np.testing.assert_array_equal(done.cpu().numpy(), d_o, err_msg=f"done @ {s}")
        util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
        if s in (0, 5, W - 1, W, L - 1, L, L + 2):
            compare_state(gpu, ora, f"step {s}")
            compare_obs(obs, ora, f"obs @ {s}")


@pytest.mark.parametrize("A,seed", [(100, 1), (100, 2), (11, 3), (500, 4)])
def test_thousand_step_rollout_within_north_star_tolerance(A, seed):
    """1,000 compounding steps: V, w' within 1e-5 relative, ints bit-exact (BASELINE.json north_star)."""
    E, W, L = 32, 50, 1000
    gpu, ora = make_pair(E, A, W, 5, T=W + L + 40, episode_len=L, seed=seed)
    g = torch.Generator().manual_seed(seed)
    for s in range(L):
        act = torch.randn(E, A, generator=g)
        _, r, done = gpu.step(act.cuda(), obs=False)
        r_o, d_o = ora.step(act.numpy())
        if s % 100 == 99 or s == L - 1:
            util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
            np.testing.assert_array_equal(done.cpu().numpy(), d_o)
            compare_state(gpu, ora, f"step {s}")
    assert done.cpu().numpy().all()


@pytest.mark.parametrize("A,seed", [(500, 11), (100, 12)])
def test_thousand_step_rollout_with_commission(A, seed):
    """Config 5's actual setting — commission 0.0025 — over 1,000 compounding steps (the mu fixed point runs every step):
    V, w' within 1e-5 relative of the oracle, ints bit-exact."""
    E, W, L = 16, 50, 1000
    gpu, ora = make_pair(E, A, W, 5, T=W + L + 40, episode_len=L, seed=seed, commission=0.0025)
    g = torch.Generator().manual_seed(seed)
    for s in range(L):
        act = torch.randn(E, A, generator=g)
        _, r, done = gpu.step(act.cuda(), obs=False)
        r_o, d_o = ora.step(act.numpy())
        if s % 100 == 99 or s == L - 1:
            util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
            np.testing.assert_array_equal(done.cpu().numpy(), d_o)
            compare_state(gpu, ora, f"step {s}")
    assert done.cpu().numpy().all()


def test_thousand_step_rollout_through_the_fused_obs_kernel():
    """The same 1,000-step bound with the observation materialised every step (default fused RT kernel path)."""
    E, A, W, L = 24, 100, 50, 1000
    gpu, ora = make_pair(E, A, W, 5, T=W + L + 40, episode_len=L, seed=5)
    g = torch.Generator().manual_seed(55)
    gpu.reset()
    for s in range(L):
        act = torch.randn(E, A, generator=g)
        obs, r, done = gpu.step(act.cuda())
        r_o, d_o = ora.step(act.numpy())
        if s in (0, 48, 49, 50, 499, 999):
            util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
            compare_state(gpu, ora, f"step {s}")
            compare_obs(obs, ora, f"obs @ {s}")
    assert done.cpu().numpy().all()


@pytest.mark.parametrize("reward", ["returns", "log_returns", "sharpe_ratio"])
@pytest.mark.parametrize("commission", [0.0, 0.0025])
def test_reward_variants_and_commission(reward, commission):
    E, A, W, L = 40, 37, 12, 30
    gpu, ora = make_pair(E, A, W, 5, episode_len=L, commission=commission, reward=reward)
    g = torch.Generator().manual_seed(5)
    for s in range(2 * L + 3):
        act = torch.randn(E, A, generator=g)
        _, r, done = gpu.step(act.cuda(), obs=False)
        r_o, d_o = ora.step(act.numpy())
        got, want = r.cpu().numpy(), r_o
        if reward == "sharpe_ratio":
            np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
            ok = ~np.isnan(want) & (ora.sharpe[:, 0] >= 4)
            np.testing.assert_allclose(got[ok], want[ok], rtol=5e-3, atol=1e-3)   # (mean-rf)/std with std ~ 1e-3: ill-conditioned in fp32 V
        else:
            util.assert_rewards_close(got, want, f"{reward} step {s}")
        compare_state(gpu, ora, f"step {s}")


def test_commission_large_asset_count():
    """C5 shape (A = 500, c = 0.0025) on a subsample of envs."""
    E, A, W, L = 24, 500, 50, 40
    gpu, ora = make_pair(E, A, W, 5, episode_len=L, commission=0.0025)
    g = torch.Generator().manual_seed(6)
    for s in range(L):
        act = torch.randn(E, A, generator=g)
        _, r, _ = gpu.step(act.cuda(), obs=False)
        r_o, _ = ora.step(act.numpy())
        util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
    compare_state(gpu, ora, "end")


def test_non_strict_normalisation():
    E, A, W, L = 16, 20, 8, 20
    gpu, ora = make_pair(E, A, W, 5, episode_len=L, strict=False)
    g = torch.Generator().manual_seed(7)
    for s in range(L):
        act = torch.rand(E, A, generator=g) * (2.0 if s % 2 else 1.0)     # non-negative, not a simplex
        _, r, _ = gpu.step(act.cuda(), obs=False)
        r_o, _ = ora.step(act.numpy())
        util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
    compare_state(gpu, ora, "end")


def test_masked_reset_and_observe():
    E, A, W, L = 20, 13, 6, 50
    gpu, ora = make_pair(E, A, W, 5, episode_len=L)
    g = torch.Generator().manual_seed(8)
    for s in range(9):
        act = torch.randn(E, A, generator=g)
        gpu.step(act.cuda(), obs=False); ora.step(act.numpy())
    mask = (torch.arange(E) % 3 == 0)
    before = gpu.observe().clone()
    obs = gpu.reset(mask=mask.cuda())
    ora.reset(mask.numpy())
    compare_state(gpu, ora, "after masked reset")
    want = ora.obs()
    got = obs.cpu().numpy()
    m = mask.numpy()
    np.testing.assert_allclose(got[m], want[m], rtol=util.RTOL_WEIGHT, atol=util.ATOL_WEIGHT)
    np.testing.assert_array_equal(got[~m], before.cpu().numpy()[~m])        # unmasked rows untouched
    compare_obs(gpu.observe(), ora, "observe")


def test_stats_vector_matches_oracle():
    E, A, W, L = 300, 21, 8, 11
    gpu, ora = make_pair(E, A, W, 5, episode_len=L, collect_stats=True)
    g = torch.Generator().manual_seed(9)
    tot = dict(n=0.0, sr=0.0, sr2=0.0, sv=0.0, slnv=0.0, nd=0.0, sep=0.0, sel=0.0, mx=-np.inf, mn=np.inf)
    for s in range(L + 3):
        act = torch.randn(E, A, generator=g)
        gpu.step(act.cuda(), obs=False)
        live = ora.t < L
        epr_before = ora.ep_return.copy()
        r_o, d_o = ora.step(act.numpy())
        r64 = r_o.astype(np.float64)[live]
        v64 = ora.value.astype(np.float64)[live]
        tot["n"] += live.sum(); tot["sr"] += r64.sum(); tot["sr2"] += (r64 ** 2).sum()
        tot["sv"] += v64.sum(); tot["slnv"] += np.log(v64).sum()
        dn = d_o.astype(bool)
        tot["nd"] += dn.sum(); tot["sep"] += ora.ep_return[dn].astype(np.float64).sum(); tot["sel"] += ora.t[dn].sum()
        if live.any():
            tot["mx"] = max(tot["mx"], v64.max()); tot["mn"] = min(tot["mn"], v64.min())
    st = gpu.stats()
    assert st["n_envs"] == tot["n"] and st["n_done"] == tot["nd"] and st["sum_ep_len"] == tot["sel"]
    np.testing.assert_allclose([st["sum_r"], st["sum_r2"], st["sum_v"], st["sum_lnv"], st["sum_ep_return"]],
                               [tot["sr"], tot["sr2"], tot["sv"], tot["slnv"], tot["sep"]], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose([st["max_v"], st["min_v"]], [tot["mx"], tot["mn"]], rtol=1e-5)


def test_full_size_properties_config2():
    """BASELINE config 2 (4,096 envs x 50 assets x window 50) through size-independent properties."""
    E, A, W, L = 4096, 50, 50, 64
    pmrl, synth, Env = _mods()
    T = 512
    tbl = synth.gbm_ohlc(T, A)
    t0 = synth.episode_offsets(E, T, W, L)
    cfg = pmrl.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=L)
    env = Env(cfg, prices=tbl, t0=t0)
    obs = env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    tblc = tbl.cuda()
    for s in range(L):
        v_prev = env.value.clone()
        act = torch.randn(E, A, generator=g, device="cuda")
        obs, r, done = env.step(act)
        w = env.weights_last
        assert torch.all(w >= 0) and torch.allclose(w.sum(1), torch.ones(E, device="cuda"), atol=2e-6)
        assert torch.all(env.value > 0)
        assert torch.allclose(r, torch.log(env.value / v_prev), atol=2e-6)          # reward = ln(V'/V) at c = 0
        assert torch.equal(env.t, torch.full((E,), s + 1, dtype=torch.int32, device="cuda"))
        assert torch.equal(env.idx, torch.full((E,), (s + 2) % W, dtype=torch.int32, device="cuda"))
        assert bool(done.all()) == (s == L - 1)
        # the obs feature channels are the table window [t0+k, t0+k+W) of every env
        e = (s * 37) % E
        r0 = int(t0[e]) + s + 1
        assert torch.equal(obs[e, :, :, :4], tblc[r0:r0 + W].permute(1, 0, 2))
        # newest weight column holds w'
        col = (W - 1) if s + 2 < W else (s + 1) % W
        assert torch.equal(obs[:, :, col, 4], w)


@pytest.mark.parametrize("E,A,commission,obs_on", [(131072, 100, 0.0, True), (65536, 100, 0.0, True), (262144, 500, 0.0025, False)],
                         ids=["config4_shard", "config3", "config5"])
def test_full_size_properties_large_configs(E, A, commission, obs_on):
    """BASELINE configs 3, 4 (per-GPU shard) and 5 at full size through size-independent properties: simplex weights,
    reward = ln of the value ratio, lockstep integer state, obs = table window + ring, and a sharding identity — a random
    sample of envs re-run as a small batch of their own reproduces the big batch bit for bit (envs are independent)."""
    W, L, steps = 50, 1000, 6
    pmrl, synth, Env = _mods()
    T = 2048
    tbl = synth.gbm_ohlc(T, A)
    t0 = synth.episode_offsets(E, T, W, L)
    cfg = pmrl.EnvConfig(num_envs=E, num_assets=A, window_size=W, commission=commission, episode_len=L)
    env = Env(cfg, prices=tbl, t0=t0)
    pick = torch.randperm(E, generator=torch.Generator().manual_seed(3))[:48].sort().values
    small = Env(pmrl.EnvConfig(num_envs=48, num_assets=A, window_size=W, commission=commission, episode_len=L), prices=tbl, t0=t0[pick])
    # ... and the CPU oracle over the same 48 envs: the full-size batch is checked against the reference restatement itself
    ora = OracleEnv(48, A, W, 5, close=tbl[:, :, 3].numpy(), feat=tbl.numpy(), t0=t0[pick].numpy(), episode_len=L,
                    commission=commission)
    pick_d = pick.cuda()
    env.reset(obs=obs_on); small.reset(obs=obs_on)
    g = torch.Generator(device="cuda").manual_seed(1)
    tblc = tbl.cuda()
    soft = torch.ones(E, dtype=torch.bool, device="cuda")
    for s in range(steps):
        v_prev = env.value.clone()
        act = torch.randn(E, A, generator=g, device="cuda")
        obs, r, done = env.step(act, obs=obs_on)
        obs_s, r_s, _ = small.step(act[pick_d].contiguous(), obs=obs_on)
        w = env.weights_last
        # quirk Q1 at scale: a raw score vector whose sum happens to lie within 1.1e-5 of 1 is NOT normalised, negative
        # entries and all (about one env in a million per step at A = 100) — the simplex properties hold for the others.
        # Such an env keeps negative holdings in its ring, which the commission factor of the NEXT step reads: it stays
        # excluded from the sign properties for the rest of the run (the reference behaves the same way).
        soft &= (act.sum(1) - 1.0).abs() > 1e-4
        assert int((~soft).sum()) < 64
        assert torch.all(w[soft] >= 0) and torch.allclose(w[soft].sum(1), torch.ones(int(soft.sum()), device="cuda"), atol=4e-6)
        assert torch.all(env.value[soft] > 0) and not bool(done.any())
        if commission == 0.0:
            assert torch.allclose(r[soft], torch.log(env.value / v_prev)[soft], atol=2e-6)
        else:                                                    # the commission factor only ever shrinks the portfolio
            assert torch.all((torch.exp(r) * v_prev >= env.value * (1 - 1e-6))[soft])
        assert torch.equal(env.t, torch.full((E,), s + 1, dtype=torch.int32, device="cuda"))
        assert torch.equal(env.idx, torch.full((E,), (s + 2) % W, dtype=torch.int32, device="cuda"))
        # sharding identity (bit-exact): the sampled envs as their own batch
        assert torch.equal(r[pick_d], r_s) and torch.equal(env.value[pick_d], small.value)
        assert torch.equal(env.hist[pick_d], small.hist)
        # oracle parity of the full-size batch on the sampled envs
        r_o, d_o = ora.step(act[pick_d].cpu().numpy())
        util.assert_rewards_close(r[pick_d].cpu().numpy(), r_o, f"full-size step {s}")
        util.assert_values_close(env.value[pick_d].cpu().numpy(), ora.value, f"full-size step {s}")
        np.testing.assert_array_equal(env.idx[pick_d].cpu().numpy(), ora.idx)
        np.testing.assert_array_equal(env.t[pick_d].cpu().numpy(), ora.t)
        np.testing.assert_array_equal(done[pick_d].cpu().numpy(), d_o)
        np.testing.assert_allclose(env.hist[pick_d].cpu().numpy(), ora.hist, rtol=util.RTOL_WEIGHT, atol=util.ATOL_WEIGHT)
        if obs_on:
            assert torch.equal(obs[pick_d], obs_s)
            compare_obs(obs[pick_d], ora, f"full-size obs @ {s}")
            e = (s * 7919) % E
            r0 = int(t0[e]) + s + 1
            assert torch.equal(obs[e, :, :, :4], tblc[r0:r0 + W].permute(1, 0, 2))
            assert torch.equal(obs[:, :, W - 1, 4], w)           # ring not full yet: newest column is the last one
            assert bool((obs[:, :, : W - 2 - s, 4] == 0).all())  # zero front padding


def test_empty_batch_is_a_no_op():
    pmrl, synth, Env = _mods()
    cfg = pmrl.EnvConfig(num_envs=0, num_assets=7, window_size=5, episode_len=10)
    env = Env(cfg, prices=synth.gbm_ohlc(64, 7))
    obs = env.reset()
    o, r, d = env.step(torch.zeros(0, 7, device="cuda"))
    assert obs.shape == (0, 7, 5, 5) and r.shape == (0,) and d.shape == (0,)


def test_bad_arguments_raise():
    pmrl, synth, Env = _mods()
    from pmrl_b200._lib import PmrlError
    tbl = synth.gbm_ohlc(64, 5)
    with pytest.raises(ValueError):
        Env(pmrl.EnvConfig(num_envs=2, num_assets=5, window_size=8, episode_len=100), prices=tbl)      # table too short
    with pytest.raises(PmrlError):
        Env(pmrl.EnvConfig(num_envs=2, num_assets=2000, window_size=8, episode_len=0)).step(
            torch.zeros(2, 2000, device="cuda"), y=torch.ones(2, 2000, device="cuda"))                 # A > 1024
    env = Env(pmrl.EnvConfig(num_envs=2, num_assets=5, window_size=8, episode_len=0))
    with pytest.raises(PmrlError):
        env.step(torch.zeros(2, 5, device="cuda"))                                                    # no table, no y
    with pytest.raises(ValueError):                                                                   # tables need an episode length:
        Env(pmrl.EnvConfig(num_envs=2, num_assets=5, window_size=8, episode_len=0), prices=tbl)       # k would run off the table
    # the same guard at the C boundary: a FULL obs gathers rows [t0+k, t0+k+W) — refused without an episode length
    from pmrl_b200 import _lib
    import ctypes as C
    ok = Env(pmrl.EnvConfig(num_envs=2, num_assets=5, window_size=8, episode_len=20), prices=tbl)
    cfg0 = _lib.PmrlEnvCfg(2, 5, 8, 5, ok.T, 0, 0, 16, 1, 25000.0, 0.0, 1.0, 0.04)
    obs = torch.zeros(2, 5, 8, 5, device="cuda")
    rc = ok.lib.pmrl_obs_build(C.byref(cfg0), ok._p_tbl, ok._p_st, obs.data_ptr(), 1, _lib.current_stream())
    assert rc == -2 and b"episode_len" in ok.lib.pmrl_last_error()


def test_nan_inf_and_overflow_propagate_like_the_reference():
    """No asserts / clamps on device: NaN, Inf and exp overflow flow through exactly like the torch reference ops."""
    E, A, W, L = 12, 9, 4, 40
    gpu, ora = make_pair(E, A, W, 5, episode_len=L)
    g = torch.Generator().manual_seed(10)
    act = torch.randn(E, A, generator=g)
    act[0, 3] = float("nan")          # NaN score: sum is NaN, min is NaN -> no normalisation -> NaN value
    act[1, :] = 100.0; act[1, 0] = -1.0   # exp(100) overflows fp32: inf / inf = NaN weights (quirk Q3)
    act[2, 2] = float("inf"); act[2, 0] = -1.0
    act[3] = torch.softmax(act[3], 0)     # well-behaved rows stay finite
    _, r, _ = gpu.step(act.cuda(), obs=False)
    with np.errstate(all="ignore"):
        r_o, _ = ora.step(act.numpy())
    got_v, want_v = gpu.value.cpu().numpy(), ora.value
    np.testing.assert_array_equal(np.isnan(got_v), np.isnan(want_v))
    np.testing.assert_array_equal(np.isnan(r.cpu().numpy()), np.isnan(r_o))
    assert np.isnan(want_v[:3]).all() and np.isfinite(want_v[3:]).all()
    ok = np.isfinite(want_v)
    util.assert_values_close(got_v[ok], want_v[ok])


@pytest.mark.parametrize("A", [129, 513, 1024])
def test_wide_asset_counts(A, tuning):
    """A up to 1024 (32 assets per lane); state-only and fused generic paths."""
    E, W, L = 10, 6, 12
    gpu, ora = make_pair(E, A, W, 5, episode_len=L)
    g = torch.Generator().manual_seed(A)
    obs = gpu.reset()
    for s in range(L + 2):
        act = torch.randn(E, A, generator=g)
        obs, r, done = gpu.step(act.cuda(), obs=(s % 2 == 0))
        r_o, d_o = ora.step(act.numpy())
        util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
        np.testing.assert_array_equal(done.cpu().numpy(), d_o)
    compare_state(gpu, ora, "end")
    compare_obs(gpu.observe(), ora, "end")


@pytest.mark.parametrize("commission", [0.0, 0.0025])
@pytest.mark.parametrize("A,E,shape", [(132, 300, 0), (256, 2600, 0), (500, 300, 0), (500, 4000, 0), (1000, 150, 0), (1024, 2400, 0),
                                       (500, 4000, 3), (500, 300, 3), (260, 3000, 3)])
def test_staged_wide_env_kernel_vs_oracle_and_register_kernel(A, E, shape, commission, tuning):
    """k_env_step_staged (rows staged one env ahead by TMA bulk copies, env_step_staged.cu) against the oracle, and bit for
    bit against the register-load kernel k_env_step on the same inputs — table-driven and external-y, auto-resets included."""
    W, L = 6, 9
    tuning.set_tuning(tuning.TUNE_CTAS_PER_SM, shape)             # CTA shape / staging depth of the staged kernel
    tuning.set_tuning(tuning.TUNE_STAGED, 2)                      # staged at any batch size
    gpu, ora = make_pair(E, A, W, 5, episode_len=L, commission=commission)
    tuning.set_tuning(tuning.TUNE_STAGED, 0)
    ref, _ = make_pair(E, A, W, 5, episode_len=L, commission=commission)
    g = torch.Generator().manual_seed(A + E)
    gpu.reset(obs=False); ref.reset(obs=False)
    n0 = tuning.load().pmrl_launch_count()
    for s in range(2 * L + 3):
        act = torch.randn(E, A, generator=g)
        if s % 5 == 2:
            act = torch.softmax(act, dim=1)
        y = (1 + 0.01 * torch.randn(E, A, generator=g)).cuda() if s % 4 == 3 else None      # external price relatives
        a = act.cuda()
        tuning.set_tuning(tuning.TUNE_STAGED, 2)
        _, r, done = gpu.step(a, y=y, obs=False)
        tuning.set_tuning(tuning.TUNE_STAGED, 0)
        _, r2, done2 = ref.step(a, y=y, obs=False)
        assert torch.equal(r, r2) and torch.equal(done, done2), f"step {s}"
        assert torch.equal(gpu.value, ref.value) and torch.equal(gpu.hist, ref.hist) and torch.equal(gpu.idx, ref.idx)
        r_o, d_o = ora.step(act.numpy(), y.cpu().numpy() if y is not None else None)
        np.testing.assert_array_equal(done.cpu().numpy(), d_o)
        util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
    compare_state(gpu, ora, "end")
    assert tuning.load().pmrl_launch_count() - n0 == 2 * (2 * L + 3)


def test_graphed_step_matches_eager():
    E, A, W, L = 64, 20, 8, 25
    gpu, ora = make_pair(E, A, W, 5, episode_len=L)
    gpu.reset()
    static_actions, replay = gpu.graphed_step(obs=True)
    g = torch.Generator().manual_seed(3)
    for s in range(2 * L + 3):
        act = torch.randn(E, A, generator=g)
        static_actions.copy_(act)
        obs, r, done = replay()
        r_o, d_o = ora.step(act.numpy())
        util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
        np.testing.assert_array_equal(done.cpu().numpy(), d_o)
    compare_state(gpu, ora, "end")
    compare_obs(obs, ora, "end")


@pytest.mark.parametrize("obs", [False, True])
def test_graphed_burst_of_steps_matches_eager(obs):
    """graphed_step(steps=K): K transitions captured in one CUDA graph give the state, rewards and dones of K eager steps
    (episode end and auto-reset inside a burst included)."""
    E, A, W, L, K = 37, 40, 8, 11, 5
    g_env, _ = make_pair(E, A, W, 5, episode_len=L, collect_stats=True)
    e_env, _ = make_pair(E, A, W, 5, episode_len=L, collect_stats=True)
    g_env.reset(obs=obs); e_env.reset(obs=obs)
    static_actions, replay = g_env.graphed_step(obs=obs, steps=K)
    assert static_actions.shape == (K, E, A)
    gen = torch.Generator().manual_seed(9)
    for b in range(4):                                     # 20 steps: crosses the episode end (11) and the auto-reset call
        acts = torch.randn(K, E, A, generator=gen).cuda()
        static_actions.copy_(acts)
        o_g, r_g, d_g = replay()
        for k in range(K):
            o_e, r_e, d_e = e_env.step(acts[k], obs=obs)
            assert torch.equal(r_g[k], r_e) and torch.equal(d_g[k], d_e), f"burst {b} step {k}"
        assert torch.equal(g_env.value, e_env.value) and torch.equal(g_env.hist, e_env.hist) and torch.equal(g_env.t, e_env.t)
        if obs:
            assert torch.equal(o_g, o_e)
    assert g_env.stats()["n_envs"] == e_env.stats()["n_envs"]


def test_window_of_one_is_rejected_like_the_reference_would_fail():
    pmrl, synth, Env = _mods()
    with pytest.raises(ValueError):
        Env(pmrl.EnvConfig(num_envs=2, num_assets=3, window_size=1, episode_len=0))


@pytest.mark.parametrize("E,chunks,pinned", [(50, 1, True), (50, 4, True), (700, 0, True), (700, 0, False), (700, 3, True), (700, -4, False),
                                             (1500, 6, True), (333, -32, True), (4096, 0, True)])
def test_step_host_chunked_matches_step(E, chunks, pinned):
    """Host-buffer step through pmrl_env_step_host (H2D of slice c+1 / D2H of slice c-1 under the kernel of slice c;
    geometric, equal and degenerate slicings) gives the same transition as the oracle."""
    A, W, L = 40, 8, 30
    gpu, ora = make_pair(E, A, W, 5, episode_len=L, collect_stats=True)
    gpu.reset()
    g = torch.Generator().manual_seed(11)
    h_r = torch.empty(E, dtype=torch.float32).pin_memory(); h_d = torch.empty(E, dtype=torch.uint8).pin_memory()
    for s in range(L + 4):
        act = torch.randn(E, A, generator=g)
        act = act.pin_memory() if pinned else act               # pinned + chunks 0 → the kernel reads host memory in place
        obs, r, d = gpu.step_host(act, h_r, h_d, chunks=chunks)
        r_o, d_o = ora.step(act.numpy())
        util.assert_rewards_close(r.numpy(), r_o, f"step {s}")
        np.testing.assert_array_equal(d.numpy(), d_o)
    compare_state(gpu, ora, "end")
    compare_obs(obs, ora, "end")
    assert gpu.stats()["n_envs"] == E * (L + 4) - E      # one auto-reset call per env is not a step


@pytest.mark.parametrize("E,A,W,commission,obs,stream", [(700, 40, 8, 0.0, True, 2), (4096, 100, 12, 0.0, True, 2), (3000, 500, 6, 0.0025, False, 2),
                                                         (3000, 500, 6, 0.0025, True, 2), (9000, 100, 10, 0.0, True, 1), (333, 7, 5, 0.0, True, 2),
                                                         (4096, 100, 12, 0.0, True, 0), (5000, 52, 50, 0.001, False, 2)])
def test_step_host_streamed_actions(E, A, W, commission, obs, stream, tuning):
    """pmrl_env_step_host with page-locked actions: the kernel is launched at once and waits per chunk for the action rows
    the copy engine streams in behind it (PmrlStepIO.actions_ready; stream = 2 forces the path at any size, 1 = by size,
    0 = zero-copy reads) — same transition as the oracle, auto-resets (envs that read no action) included, through the
    fused, the two-kernel, the register-load and the staged wide-env kernels."""
    L = W + 5
    tuning.set_tuning(tuning.TUNE_HOST_STREAM, stream)
    tuning.set_tuning(tuning.TUNE_STAGED, 2)
    gpu, ora = make_pair(E, A, W, 5, episode_len=L, commission=commission)
    gpu.reset(obs=obs)
    g = torch.Generator().manual_seed(E + A)
    h_r = torch.empty(E, dtype=torch.float32).pin_memory(); h_d = torch.empty(E, dtype=torch.uint8).pin_memory()
    acts = [torch.randn(E, A, generator=g).pin_memory() for _ in range(3)]
    for s in range(L + 3):
        act = acts[s % 3]
        o, r, d = gpu.step_host(act, h_r, h_d, obs=obs)
        r_o, d_o = ora.step(act.numpy())
        util.assert_rewards_close(r.numpy(), r_o, f"step {s}")
        np.testing.assert_array_equal(d.numpy(), d_o)
    compare_state(gpu, ora, "end")
    if obs:
        compare_obs(o, ora, "end")


FUZZ_SEEDS = int(os.environ.get("PMRL_FUZZ_SEEDS", "40"))          # a soak run sets this to a few hundred


@pytest.mark.parametrize("seed", range(FUZZ_SEEDS))
def test_random_shape_fuzz(seed):
    """Random (E, A, W, F, commission, reward, episode length) against the oracle: every dispatch path gets exercised
    (RT / register-ring / generic fused kernels, state-only pipelined and plain kernels, partial groups and tiles)."""
    rs = np.random.RandomState(1000 + seed)
    A = int(rs.choice([1, 2, 5, 16, 31, 32, 36, 50, 64, 100, 128, 129, 200, 300]))
    W = int(rs.choice([2, 3, 8, 17, 32, 50, 64, 65]))
    F = int(rs.choice([5, 5, 5, 2, 4, 7]))
    E = int(rs.randint(1, 70))
    L = int(rs.randint(2, W + 6))
    c = float(rs.choice([0.0, 0.0, 0.0025, 0.02]))
    reward = str(rs.choice(["step_log", "returns", "log_returns"]))
    if seed >= 24:                                                  # round-2 paths: wide feature sets through the fused kernel,
        A = int(rs.choice([32, 36, 50, 64, 100, 128, 132, 260, 500, 1000]))    # wide envs through the TMA-staged step
        F = int(rs.choice([5, 9, 9, 13, 17, 6, 12]))
        W = int(rs.choice([2, 8, 17, 32, 50, 64]))
        E = int(rs.randint(1, 70)) if A <= 128 else int(rs.choice([9, 40, 2400]))
        if E == 2400:                                               # more than one env per warp → the staged kernel by default;
            W = int(rs.choice([2, 8, 17]))                          # kept small and state-only so the oracle side stays cheap
        L = int(rs.randint(2, min(W, 12) + 6))
        from pmrl_b200 import _lib
        _lib.set_tuning(_lib.TUNE_STAGED, int(rs.choice([1, 2])))
        _lib.set_tuning(_lib.TUNE_CTAS_PER_SM, int(rs.choice([0, 3])) if A > 128 else 0)
    gpu, ora = make_pair(E, A, W, F, episode_len=L, commission=c, reward=reward, seed=seed)
    g = torch.Generator().manual_seed(seed)
    big = E * A * W > 4_000_000
    obs = gpu.reset(obs=not big)
    if not big:
        compare_obs(obs, ora, "reset")
    for s in range(2 * L + 3):
        kind = s % 4
        act = torch.randn(E, A, generator=g)
        if kind == 1:
            act = torch.softmax(act, dim=1)
        elif kind == 2:
            act = torch.rand(E, A, generator=g) + 1e-3                 # non-negative, not a simplex (quirk Q1): sum in [0.8, 1.2]
            act = act / act.sum(1, keepdim=True) * (0.8 + 0.4 * torch.rand(E, 1, generator=g))
        want_obs = bool(rs.randint(0, 2)) and not big
        obs, r, done = gpu.step(act.cuda(), obs=want_obs)
        r_o, d_o = ora.step(act.numpy())
        np.testing.assert_array_equal(done.cpu().numpy(), d_o, err_msg=f"done @ {s}")
        util.assert_rewards_close(r.cpu().numpy(), r_o, f"step {s}")
        if want_obs:
            compare_obs(obs, ora, f"obs @ {s}")
    compare_state(gpu, ora, "end")
    if seed >= 24:
        _lib.set_tuning(_lib.TUNE_STAGED, 1); _lib.set_tuning(_lib.TUNE_CTAS_PER_SM, 0)


@pytest.mark.parametrize("A,W,commission", [(100, 50, 0.0), (500, 50, 0.0025), (11, 8, 0.02), (200, 16, 0.0)])
def test_price_relatives_table_equals_ieee_division_and_the_external_y_path(A, W, commission):
    """PmrlTables.y_tm is close[t]/close[t-1] divided once per table (pmrl_price_relatives): it must equal the IEEE
    quotient bit for bit, and stepping from the table must give the same bits as stepping with those y rows supplied
    externally (the reference's call shape, trading_env.py:44)."""
    pmrl, synth, Env = _mods()
    E, L = 40, 30
    T = W + L + 40
    tbl = synth.gbm_ohlc(T, A, 77)
    t0 = synth.episode_offsets(E, T, W, L)
    cfg = pmrl.EnvConfig(num_envs=E, num_assets=A, window_size=W, commission=commission, episode_len=L)
    tab = Env(cfg, prices=tbl, t0=t0)
    ref_y = torch.ones_like(tab.close_tm); ref_y[1:] = tab.close_tm[1:] / tab.close_tm[:-1]
    assert torch.equal(tab.y_tm, ref_y)
    ext = Env(pmrl.EnvConfig(num_envs=E, num_assets=A, window_size=W, commission=commission, episode_len=0))
    g = torch.Generator().manual_seed(5)
    rows = t0.cuda().long()
    for s in range(L):
        act = torch.randn(E, A, generator=g).cuda()
        if s % 4 == 1:
            act = torch.softmax(act, dim=1)
        _, r1, _ = tab.step(act, obs=False)
        _, r2, _ = ext.step(act, y=ref_y[rows + s + W], obs=False)
        assert torch.equal(r1, r2), f"step {s}"
        assert torch.equal(tab.value, ext.value) and torch.equal(tab.hist, ext.hist), f"step {s}"


def test_shared_divisor_quotients_equal_ieee_division_bit_for_bit():
    """The kernels divide a whole env by one divisor (softmax sum, new portfolio value) through a shared refined
    reciprocal; pmrl_selftest_division compares that quotient with IEEE division on the same operands."""
    from pmrl_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(2024)
    n = 1 << 25

    def run(num, den):
        out = torch.zeros(2, dtype=torch.int64, device="cuda")
        _lib.check(lib.pmrl_selftest_division(num.data_ptr(), den.data_ptr(), num.numel(), out.data_ptr(), _lib.current_stream()), "selftest")
        bad, tested = out.cpu().tolist()
        return bad, tested

    def rand_float(k, lo_exp, hi_exp):                        # random sign-less mantissa, exponent uniform in [lo, hi)
        mant = torch.randint(0, 1 << 23, (k,), generator=g, device="cuda", dtype=torch.int32)
        ex = torch.randint(lo_exp + 127, hi_exp + 127, (k,), generator=g, device="cuda", dtype=torch.int32)
        return ((ex << 23) | mant).view(torch.float32)

    # (1) the whole accepted range, both signs
    num = rand_float(n, -60, 60) * (torch.randint(0, 2, (n,), generator=g, device="cuda") * 2 - 1).float()
    den = rand_float(n // 32, -60, 60)
    bad, tested = run(num, den)
    assert bad == 0 and tested == n
    # (2) softmax-like: numerators e^x, divisor their sum; (3) drift-like: holdings / their sum
    x = torch.randn(n // 32, 32, generator=g, device="cuda") * 3.0
    e = torch.exp(x)
    bad, tested = run(e.reshape(-1), e.sum(dim=1))
    assert bad == 0 and tested == n
    w = torch.softmax(x, dim=1) * 25000.0 * (1.0 + 0.02 * torch.randn(n // 32, 32, generator=g, device="cuda"))
    bad, tested = run(w.reshape(-1).contiguous(), w.sum(dim=1))
    assert bad == 0 and tested == n
    # (4) mantissa corner cases: all-ones / all-zeros / one-bit mantissas against each other, zeros among the numerators
    corner = torch.tensor([0x3F800000, 0x3FFFFFFF, 0x3F800001, 0x3FC00000, 0x3FFFFFFE, 0x40490FDB, 0x3EAAAAAB, 0x00000000],
                          dtype=torch.int32, device="cuda").view(torch.float32)
    num = corner.repeat(4 * 1024)                             # 32,768 numerators, 32 per divisor
    den = rand_float(num.numel() // 32, -3, 3)
    den[:8] = corner[:8].abs().clamp_min(1.0)
    bad, tested = run(num, den)
    assert bad == 0 and tested == num.numel()
    # out-of-range operands are skipped, not mis-divided
    bad, tested = run(torch.full((64,), 1e-30, device="cuda"), torch.ones(2, device="cuda"))
    assert (bad, tested) == (0, 0)


def test_launch_counter_tells_the_fused_path_from_the_two_kernel_path():
    """pmrl_launch_count (bench.py's gpu_launches): one kernel per step where a specialised fused kernel covers the shape,
    two (step + obs tile kernel) elsewhere, one for a state-only step."""
    from pmrl_b200 import _lib
    lib = _lib.load()

    def launches(A, obs):
        gpu, _ = make_pair(16, A, 8, 5, episode_len=20)
        gpu.reset(obs=obs)
        act = torch.randn(16, A, device="cuda")
        c0 = lib.pmrl_launch_count()
        gpu.step(act, obs=obs)
        return lib.pmrl_launch_count() - c0

    assert launches(40, True) == 1        # RT kernel
    assert launches(12, True) == 1        # register-ring kernel (A < 32)
    assert launches(200, True) == 2       # wide universe: k_env_step + k_obs_build_rows
    assert launches(200, False) == 1
