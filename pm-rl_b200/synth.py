"""Seeded synthetic GBM OHLC tables (SURVEY.md §8(d) recipe) — generated on the CPU so the CPU oracle and
the GPU kernels see identical bits, then copied to the device by the caller."""
from __future__ import annotations

import torch


def gbm_ohlc(T: int, A: int, seed: int = 1234) -> torch.Tensor:
    """[T, A, 4] float32 (open, high, low, close); asset 0 is cash (all ones → y ≡ 1)."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(T, A, generator=g, dtype=torch.float64)
    z2 = torch.randn(T, A, generator=g, dtype=torch.float64).abs()
    z3 = torch.randn(T, A, generator=g, dtype=torch.float64).abs()
    a = torch.arange(A, dtype=torch.float64)
    sigma = 0.01 * (1.0 + (a % 5) / 5.0)
    mu = 2e-4
    logret = (mu - 0.5 * sigma ** 2) + sigma * z
    close = 100.0 * torch.exp(torch.cumsum(logret, dim=0))
    open_ = torch.cat([torch.full((1, A), 100.0, dtype=torch.float64), close[:-1]], dim=0)
    high = torch.maximum(open_, close) * torch.exp(z2 * sigma / 2)
    low = torch.minimum(open_, close) * torch.exp(-z3 * sigma / 2)
    tbl = torch.stack([open_, high, low, close], dim=-1)
    tbl[:, 0, :] = 1.0
    return tbl.to(torch.float32).contiguous()


def episode_offsets(E: int, T: int, W: int, episode_len: int, first_env: int = 0) -> torch.Tensor:
    """t0[e] = (e * 2654435761 mod 2^32) mod (T - W - L_ep), keyed by the GLOBAL env id so that a rank's
    shard reproduces the single-GPU run (SURVEY.md §8(d),(e))."""
    span = T - W - episode_len
    if span < 1:
        raise ValueError("table too short for the requested window + episode length")
    e = torch.arange(first_env, first_env + E, dtype=torch.int64)
    return ((e * 2654435761) % (1 << 32) % span).to(torch.int32)
