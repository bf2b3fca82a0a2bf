"""Multi-rank host logic on CPU (gloo, world_size 2): env sharding by contiguous slices and the statistics all-reduce.
The per-shard statistics come from the CPU oracle (test infrastructure); the product code under test is pmrl_b200.dist."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pmrl_b200 import dist as pdist


def test_shard_ranges_partition_the_batch():
    E = 1048576
    for world in (1, 2, 4, 8):
        spans = [pdist.shard_range(E, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == E
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert len({b - a for a, b in spans}) == 1
    with pytest.raises(ValueError):
        pdist.shard_range(10, 0, 4)


def test_all_reduce_stats_is_identity_without_a_group():
    v = torch.arange(10, dtype=torch.float64)
    assert torch.equal(pdist.all_reduce_stats(v), v)


def test_async_reducer_is_a_snapshot_without_a_group():
    ar = pdist.AsyncStatsReducer("cpu")
    v = torch.arange(10, dtype=torch.float64)
    ar.launch(v)
    v += 1.0                                  # the running vector moves on; the snapshot does not
    assert not ar.active and torch.equal(ar.result(), torch.arange(10, dtype=torch.float64))


def _shard_stats(first, n, A, W, L, steps, seed):
    """Stats vector of envs [first, first+n) produced by the oracle; actions are keyed by the global env id."""
    from oracle.env_oracle import OracleEnv
    from pmrl_b200 import synth
    T = 256
    tbl = synth.gbm_ohlc(T, A)
    t0 = synth.episode_offsets(n, T, W, L, first_env=first)
    env = OracleEnv(n, A, W, 5, close=tbl[:, :, 3].numpy(), feat=tbl.numpy(), t0=t0.numpy(), episode_len=L)
    st = np.zeros(10); st[8:] = -np.inf
    for s in range(steps):
        act = np.stack([np.random.RandomState(seed + 7919 * (first + e) + s).standard_normal(A) for e in range(n)]).astype(np.float32)
        live = env.t < L
        r, d = env.step(act)
        r64, v64 = r.astype(np.float64)[live], env.value.astype(np.float64)[live]
        st[0] += live.sum(); st[1] += r64.sum(); st[2] += (r64 ** 2).sum(); st[3] += v64.sum(); st[4] += np.log(v64).sum()
        dn = d.astype(bool)
        st[5] += dn.sum(); st[6] += env.ep_return[dn].astype(np.float64).sum(); st[7] += env.t[dn].sum()
        if live.any():
            st[8] = max(st[8], v64.max()); st[9] = max(st[9], (-v64).max())
    return st, env.value.copy()


def _worker(rank, world, port, E, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = pdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = pdist.shard_range(E, rank, world)
    st, vals = _shard_stats(lo, hi - lo, 7, 6, 9, 12, seed=3)
    red = pdist.all_reduce_stats(torch.from_numpy(st))
    # the per-step asynchronous reducer: launch on a running vector several times, the last launch wins
    ar = pdist.AsyncStatsReducer("cpu")
    assert ar.active
    running = torch.from_numpy(st) * 0.0
    for frac in (0.25, 0.5, 1.0):
        running = torch.from_numpy(st) * frac if frac < 1.0 else torch.from_numpy(st)
        ar.launch(running)
    assert ar.launches == 3 and torch.allclose(ar.result(), red, rtol=1e-14, atol=0.0)
    out[rank] = (red.numpy().copy(), vals)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_stats_equal_single_process():
    E = 12
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, E, out), nprocs=2, join=True)
    single, vals = _shard_stats(0, E, 7, 6, 9, 12, seed=3)
    for rank in (0, 1):
        np.testing.assert_allclose(out[rank][0], single, rtol=1e-12)
    # envs are independent and keyed by global id: the concatenated shards reproduce the single-process values bit for bit
    np.testing.assert_array_equal(np.concatenate([out[0][1], out[1][1]]), vals)
