"""Per-env evaluation metrics on device (util/eval.py:14-37): Sharpe, Sortino, max drawdown, average turnover."""
from __future__ import annotations

import torch

from . import _lib


def eval_metrics(values, weights=None, rf: float = 0.0, periods: int = 252):
    """values [E, N] portfolio-value history (values[:, 0] = initial cash), weights [E, N, A] or None.
    Returns [E, 4] = (sharpe, sortino, max_drawdown, average_turnover)."""
    lib = _lib.load()
    v = values.to(torch.float32).contiguous()
    E, N = v.shape
    w = weights.to(torch.float32).contiguous() if weights is not None else None
    A = w.shape[2] if w is not None else 1
    out = torch.empty(E, 4, dtype=torch.float32, device=v.device)
    rc = lib.pmrl_eval_metrics(_lib.ptr(v), _lib.ptr(w), E, N, A, float(rf), int(periods), _lib.ptr(out), _lib.current_stream())
    _lib.check(rc, "pmrl_eval_metrics")
    return out
