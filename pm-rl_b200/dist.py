"""Multi-GPU plumbing: one process per GPU, the env batch sharded by contiguous slices, no data-path
collective.  The only exchange is the all-reduce of the 10-double statistics vector (NCCL over NVLink on
the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

N_SUM_STATS = 8     # entries [0, 8) are sums, [8, 10) are maxima (include/pmrl_b200.h PMRL_STAT_*)


def shard_range(num_envs_global: int, rank: int, world: int) -> tuple[int, int]:
    """Rank r owns global envs [r*E/R, (r+1)*E/R); E must divide evenly."""
    if num_envs_global % world != 0:
        raise ValueError(f"num_envs {num_envs_global} is not divisible by world size {world}")
    per = num_envs_global // world
    return rank * per, (rank + 1) * per


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise the default process group from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*).
    Returns (rank, world, local_rank); a single-process run needs no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def all_reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """SUM over the first 8 entries, MAX over the last 2; identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    sums = stats[:N_SUM_STATS].clone()
    maxs = stats[N_SUM_STATS:].clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    return torch.cat([sums, maxs])


class AsyncStatsReducer:
    """All-reduce of the statistics vector every step, off the step's critical path (one SUM over 8
    doubles + one MAX over 2, K = 1 for config 4).  launch() snapshots the running vector into private buffers and
    starts both collectives asynchronously (NCCL runs them on its own stream behind the snapshot); the next step's
    kernel is not ordered after them.  result() waits and returns the reduced 10-vector of the last launch."""

    def __init__(self, device, dtype=torch.float64):
        self.sums = torch.zeros(N_SUM_STATS, dtype=dtype, device=device)
        self.maxs = torch.zeros(2, dtype=dtype, device=device)
        self.pending = []
        self.launches = 0
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _drain(self):
        for w in self.pending:
            w.wait()                       # stream-level wait on a collective issued a whole step ago
        self.pending = []

    def launch(self, stats: torch.Tensor):
        self._drain()
        self.sums.copy_(stats[:N_SUM_STATS])
        self.maxs.copy_(stats[N_SUM_STATS:])
        if self.active:
            self.pending = [dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, async_op=True),
                            dist.all_reduce(self.maxs, op=dist.ReduceOp.MAX, async_op=True)]
        self.launches += 1

    def result(self) -> torch.Tensor:
        self._drain()
        return torch.cat([self.sums, self.maxs])
