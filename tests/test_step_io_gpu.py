"""GPU tests of the round-2 entry points: K-step bursts, kernel-written buffer rows (PmrlStepIO sinks), per-batch work
counters under concurrent streams / CUDA graphs, host mirrors of reward / done."""
import numpy as np
import pytest
import torch

from tests import util
from tests.test_env_gpu import make_pair, compare_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("A,W,commission,E", [(50, 50, 0.0, 300), (50, 50, 0.0025, 70), (100, 50, 0.0, 64), (100, 16, 0.0025, 40),
                                              (500, 50, 0.0025, 24), (500, 8, 0.0, 9), (11, 8, 0.0025, 33), (33, 6, 0.0, 20),
                                              (130, 5, 0.0025, 12), (1000, 4, 0.0025, 5)])
def test_step_burst_is_bit_identical_to_eager_steps(A, W, commission, E):
    """pmrl_env_step_burst: 15 steps in one launch (HORIZON, config/dreamer.py:54) = 15 eager state-only steps bit for bit,
    across the episode end and the auto-reset call; and both agree with the oracle."""
    L, K = 2 * W + 3, 15
    b_env, ora = make_pair(E, A, W, 5, episode_len=L, commission=commission, collect_stats=True)
    e_env, _ = make_pair(E, A, W, 5, episode_len=L, commission=commission, collect_stats=True)
    gen = torch.Generator().manual_seed(A + W)
    for b in range(-(-(L + 20) // K)):
        acts = torch.randn(K, E, A, generator=gen)
        if b % 2:
            acts[3] = torch.softmax(acts[3], dim=1)            # a simplex row: pass-through branch inside a burst
        acts_d = acts.cuda()
        r_b, d_b = b_env.step_burst(acts_d)
        for k in range(K):
            _, r_e, d_e = e_env.step(acts_d[k], obs=False)
            assert torch.equal(r_b[k], r_e) and torch.equal(d_b[k], d_e), f"burst {b} step {k}"
            r_o, d_o = ora.step(acts[k].numpy())
            util.assert_rewards_close(r_e.cpu().numpy(), r_o, f"burst {b} step {k}")
            np.testing.assert_array_equal(d_e.cpu().numpy(), d_o)
        for name in ("value", "hist", "idx", "is_full", "t", "ep_return"):
            assert torch.equal(getattr(b_env, name), getattr(e_env, name)), f"{name} after burst {b}"
    compare_state(b_env, ora, "end")
    sb, se = b_env.stats(), e_env.stats()
    assert sb["n_envs"] == se["n_envs"] and sb["n_done"] == se["n_done"]
    np.testing.assert_allclose(sb["sum_r"], se["sum_r"], rtol=1e-9)


def test_step_burst_full_size_config2_properties():
    """BASELINE config 2 (4,096 x 50 x 50) in 15-step bursts: lockstep integer state, simplex weights, value ratio = exp(sum r)."""
    E, A, W, L, K = 4096, 50, 50, 64, 15
    env, _ = make_pair(E, A, W, 5, T=512, episode_len=L)
    g = torch.Generator(device="cuda").manual_seed(1)
    v0 = env.value.clone()
    acts = torch.randn(K, E, A, generator=g, device="cuda")
    r, d = env.step_burst(acts)
    assert torch.equal(env.t, torch.full((E,), K, dtype=torch.int32, device="cuda")) and not bool(d.any())
    w = env.weights_last
    assert torch.all(w >= 0) and torch.allclose(w.sum(1), torch.ones(E, device="cuda"), atol=2e-6)
    assert torch.allclose(torch.log(env.value / v0), r.double().sum(0).float(), atol=1e-5)


def test_step_io_sinks_receive_the_rows_the_buffers_store():
    E, A, W, L = 70, 100, 8, 20
    env, ora = make_pair(E, A, W, 5, episode_len=L)
    g = torch.Generator().manual_seed(3)
    env.reset()
    for s in range(L + 3):                                       # through the auto-reset item too
        act = torch.randn(E, A, generator=g).cuda()
        rr = torch.full((E,), -7.0, device="cuda"); dd = torch.full((E,), 9, dtype=torch.uint8, device="cuda")
        a_s = torch.zeros(E, A, device="cuda"); v_s = torch.zeros(E, device="cuda"); w_s = torch.zeros(E, A, device="cuda")
        i_s = torch.zeros(E, dtype=torch.int32, device="cuda")
        out = torch.zeros(E, A, W, 5, device="cuda")
        was_reset = bool((env.t >= L).all())
        obs, r, d = env.step_io(act, obs=(s % 2 == 0), out=out if s % 2 == 0 else None, reward=rr, done=dd, action_sink=a_s,
                                value_sink=v_s, weight_sink=w_s, index_sink=i_s)
        r_o, d_o = ora.step(act.cpu().numpy())
        assert r.data_ptr() == rr.data_ptr() and d.data_ptr() == dd.data_ptr()
        util.assert_rewards_close(rr.cpu().numpy(), r_o, f"step {s}")
        np.testing.assert_array_equal(dd.cpu().numpy(), d_o)
        assert torch.equal(v_s, env.value) and torch.equal(w_s, env.weights_last)
        assert torch.equal(i_s, env.t0 + env.t)
        if was_reset:                                            # the reset item stores the all-cash row (rollout_buffer.py:36)
            assert bool((a_s[:, 0] == 1).all()) and bool((a_s[:, 1:] == 0).all())
        else:
            assert torch.equal(a_s, act)
        if s % 2 == 0:
            assert obs.data_ptr() == out.data_ptr()
            np.testing.assert_array_equal(out.cpu().numpy()[..., :4], ora.obs()[..., :4])
    compare_state(env, ora, "end")


def test_two_envs_on_two_streams_one_replaying_a_graph():
    """The fused kernel's work counters belong to the env batch (PmrlEnvState.ticket): a 15-step graph replaying on one
    stream and 200 eager fused steps of another env interleaved on a second stream both reproduce their single-stream
    runs bit for bit."""
    E, A, W, L, K = 2500, 40, 8, 60, 15
    ga, _ = make_pair(E, A, W, 5, episode_len=L, seed=1)
    gb, _ = make_pair(E + 300, A, W, 5, episode_len=L, seed=2)
    ra, _ = make_pair(E, A, W, 5, episode_len=L, seed=1)
    rb, _ = make_pair(E + 300, A, W, 5, episode_len=L, seed=2)
    for e in (ga, gb, ra, rb):
        e.reset()
    assert ga._ticket.data_ptr() != gb._ticket.data_ptr()
    gen = torch.Generator(device="cuda").manual_seed(5)
    acts_a = torch.randn(14, K, E, A, generator=gen, device="cuda")
    acts_b = torch.randn(200, E + 300, A, generator=gen, device="cuda")
    # reference runs, one after the other on the default stream
    obs_ra = obs_rb = None
    for n in range(14):
        for k in range(K):
            obs_ra, _, _ = ra.step(acts_a[n, k])
    for n in range(200):
        obs_rb, _, _ = rb.step(acts_b[n])
    obs_ra, obs_rb = obs_ra.clone(), obs_rb.clone()
    torch.cuda.synchronize()
    # concurrent runs
    static_actions, replay = ga.graphed_step(obs=True, steps=K)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    obs_a = obs_b = None
    for n in range(200):
        if n % 14 == 0 and n // 14 < 14:
            with torch.cuda.stream(s1):
                static_actions.copy_(acts_a[n // 14])
                obs_a = replay()[0]
        with torch.cuda.stream(s2):
            obs_b, _, _ = gb.step(acts_b[n])
    torch.cuda.synchronize()
    for name in ("value", "hist", "idx", "t"):
        assert torch.equal(getattr(ga, name), getattr(ra, name)), f"graph env {name}"
        assert torch.equal(getattr(gb, name), getattr(rb, name)), f"eager env {name}"
    assert torch.equal(obs_a, obs_ra) and torch.equal(obs_b, obs_rb)
    assert int(ga._ticket.abs().sum()) == 0 and int(gb._ticket.abs().sum()) == 0      # counters re-armed


@pytest.mark.parametrize("pinned_results", [True, False])
def test_step_host_writes_results_into_mapped_host_memory(pinned_results):
    """Zero-copy host step: with page-locked reward / done buffers the kernel writes them itself (no D2H copy); pageable
    result buffers take the copy path.  Same values either way, and the device-side buffers stay current."""
    E, A, W, L = 700, 100, 50, 60
    env, ora = make_pair(E, A, W, 5, episode_len=L)
    ref, _ = make_pair(E, A, W, 5, episode_len=L)
    env.reset(); ref.reset()
    g = torch.Generator().manual_seed(4)
    h_r = torch.full((E,), -3.0); h_d = torch.full((E,), 7, dtype=torch.uint8)
    if pinned_results:
        h_r, h_d = h_r.pin_memory(), h_d.pin_memory()
    for s in range(L + 2):
        act = torch.randn(E, A, generator=g).pin_memory()
        obs, r, d = env.step_host(act, h_r, h_d)
        obs_r, r_r, d_r = ref.step(act.cuda())
        assert torch.equal(r, r_r.cpu()) and torch.equal(d, d_r.cpu()), f"step {s}"
        assert torch.equal(env.reward, r_r) and torch.equal(env.done, d_r)
        assert torch.equal(obs, obs_r)
