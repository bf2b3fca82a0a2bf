"""Differentiable batched reward of the PG agent (agent/pg/pg.py:40-82) as one fused forward+backward kernel.

`PGReward.apply(a, _v, _a, p, mode, normalise, commission, scale)` returns the scalar `torch.mean(reward)` the
reference's `_reward` returns, with the analytic gradient w.r.t. the raw action `a` produced in the same launch.
"""
from __future__ import annotations

import torch

from . import _lib

_MODES = {"returns": 1, "log_returns": 2}
# "sharpe_ratio" (pg.py:79-80: mean(ret) / std(ret) * REWARD_SCALE over the batch) is a batch-level statistic: the kernel
# supplies the per-row gross return and its Jacobian wrt the raw action (RETURNS mode), the two batch reductions and the
# chain rule  dr/dret_b = [1/(B s) - m (ret_b - m) / ((B - 1) s^3)] * scale  are a handful of device ops on [B] vectors.


def reference_normalise_condition(a) -> bool:
    """pg.py:52 — `not isclose(sum(a), 1, atol=1e-6) or min(a) < 0` over the WHOLE batch tensor (one host sync)."""
    return bool((not torch.isclose(torch.sum(a), torch.tensor(1.0, device=a.device), atol=1e-6)) or (torch.min(a) < 0))


def pg_reward(a, pv, pa, p, mode: str = "log_returns", normalise: bool = True, commission: float = 0.0,
              scale: float = 1.0, gscale: float = 1.0, mu_max_iter: int = 16, want_grad: bool = True):
    """a [B, A(,1)] raw actions, pv [B(,1,1)] previous values, pa [B, A(,1)] previous weights, p [B, A(,1)] price relatives.
    Returns (rew [B], grad_a [B, A] = gscale * d mean(rew) / d a  or None)."""
    if mode == "sharpe_ratio":
        B = a.shape[0]
        ret, jac = pg_reward(a, pv, pa, p, "returns", normalise, commission, 1.0, float(B), mu_max_iter, want_grad)   # jac = d ret_b / d a_b
        m, sd = ret.mean(), ret.std()                            # torch.std: unbiased (ddof = 1), like the reference
        rew = (m / sd * scale).expand(B)                         # the reference returns ONE scalar; mean over [B] gives it back
        if not want_grad:
            return rew, None
        coef = (1.0 / (B * sd) - m * (ret - m) / ((B - 1) * sd ** 3)) * (scale * gscale)
        return rew, jac * coef[:, None]
    lib = _lib.load()
    B = a.shape[0]
    A = a.numel() // max(B, 1)
    a2 = a.detach().reshape(B, A).to(torch.float32).contiguous()
    pv2 = pv.detach().reshape(B).to(torch.float32).contiguous()
    pa2 = pa.detach().reshape(B, A).to(torch.float32).contiguous() if pa is not None else None
    p2 = p.detach().reshape(B, A).to(torch.float32).contiguous()
    rew = torch.empty(B, dtype=torch.float32, device=a2.device)
    grad = torch.empty(B, A, dtype=torch.float32, device=a2.device) if want_grad else None
    rc = lib.pmrl_pg_reward_fwd_bwd(B, A, _MODES[mode], 1 if normalise else 0, float(commission), float(scale), mu_max_iter,
                                    _lib.ptr(a2), _lib.ptr(pv2), _lib.ptr(pa2), _lib.ptr(p2), _lib.ptr(rew), _lib.ptr(grad),
                                    float(gscale), _lib.current_stream())
    _lib.check(rc, "pmrl_pg_reward_fwd_bwd")
    return rew, grad


class PGReward(torch.autograd.Function):
    """mean_b reward_b with d/da from the fused kernel (drop-in for `PG._reward`, pg.py:40-82)."""

    @staticmethod
    def forward(ctx, a, pv, pa, p, mode="log_returns", normalise=True, commission=0.0, scale=1.0):
        rew, grad = pg_reward(a, pv, pa, p, mode, normalise, commission, scale)
        ctx.save_for_backward(grad.reshape(a.shape))
        return rew.mean()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None, None, None
