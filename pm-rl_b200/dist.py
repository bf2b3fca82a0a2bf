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
    """All-reduce of the statistics vector every step, off the step's critical path (SUM over 8
    doubles, MAX over 2, K = 1 for config 4).  launch() snapshots the running vector into a private buffer (one small
    copy on the caller's stream) and starts ONE asynchronous collective — an all-gather of the 10 doubles, so that the
    SUM and the MAX part travel together; NCCL runs it on its own stream behind the snapshot and the next step's kernel
    is not ordered after it.  result() waits and folds the gathered [world, 10] block (sum over ranks / max over ranks).
    Measured on 2 x B200 at config 4: two all-reduces per step cost 1.4 % of the step, see DESIGN.md §6."""

    def __init__(self, device, dtype=torch.float64):
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.active else 1
        self.snap = torch.zeros(N_SUM_STATS + 2, dtype=dtype, device=device)
        self.gathered = torch.zeros(self.world * (N_SUM_STATS + 2), dtype=dtype, device=device)
        self.pending = []
        self.launches = 0

    def _drain(self):
        for w in self.pending:
            w.wait()                       # stream-level wait on a collective issued a whole step ago
        self.pending = []

    def launch(self, stats: torch.Tensor):
        self._drain()
        self.snap.copy_(stats)
        if self.active:
            self.pending = [dist.all_gather_into_tensor(self.gathered, self.snap, async_op=True)]
        self.launches += 1

    def result(self) -> torch.Tensor:
        self._drain()
        if not self.active:
            return self.snap.clone()
        g = self.gathered.view(self.world, N_SUM_STATS + 2)
        return torch.cat([g[:, :N_SUM_STATS].sum(dim=0), g[:, N_SUM_STATS:].max(dim=0).values])
