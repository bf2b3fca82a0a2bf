"""Harness that imports the LIVE reference (read-only tree at /root/reference) — test infrastructure.

Only usable where the reference tree exists (the build container); it never travels to the GPU box.
It is used (a) by tests/golden/make_golden.py to produce the committed fixtures and (b) by
tests marked `needs_reference` to cross-check the restated oracle against the reference itself.

The reference binds its shape constants by value at import (`from config.base import WINDOW_SIZE,
NUM_ASSETS`, env/sim/weight_buffer.py:1), so the harness patches `config.base` and reloads the env
modules for every shape.
"""
from __future__ import annotations

import importlib
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def locate() -> str | None:
    """the reference tree is looked for at $PMRL_REFERENCE_ROOT, `/root/reference` (the build
    container) and `baseline/_ref/` (it is not pip-installable — no setup.py / pyproject.toml — so that directory only
    exists if an operator puts a checkout there).  None when no candidate holds env/sim/trading_env.py."""
    for cand in (os.environ.get("PMRL_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "env", "sim", "trading_env.py")):
            return cand
    return None


REFERENCE_ROOT = locate() or "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "env", "sim", "trading_env.py"))


def _ensure_path():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True            # the tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_env_module(num_assets: int, window: int, commission: float = 0.0, patch_maximum: bool = False):
    """Returns the reference's `env.sim.trading_env` module bound to (num_assets, window, commission).

    patch_maximum=True wraps `torch.maximum` *inside that module only* so that the broken one-argument
    call at trading_env.py:72 evaluates relu(x) (the upstream PGPortfolio form it cites at :66)."""
    _ensure_path()
    import config.base as cb
    cb.NUM_ASSETS, cb.WINDOW_SIZE, cb.COMISSION = int(num_assets), int(window), float(commission)
    import env.sim.weight_buffer as wb
    import env.reward as rw
    import env.sim.trading_env as te
    importlib.reload(wb)
    importlib.reload(rw)
    importlib.reload(te)
    if patch_maximum:
        import types
        import torch

        class _TorchProxy(types.ModuleType):
            def __getattr__(self, name):
                return getattr(torch, name)

        proxy = _TorchProxy("torch")
        proxy.maximum = lambda x, *a: torch.maximum(x, *a) if a else torch.clamp_min(x, 0)
        te.torch = proxy
    return te


def load_rollout_buffer_module(num_assets: int, window: int, batch_size: int):
    _ensure_path()
    import config.base as cb
    cb.NUM_ASSETS, cb.WINDOW_SIZE, cb.BATCH_SIZE = int(num_assets), int(window), int(batch_size)
    import replay.rollout_buffer as rb
    importlib.reload(rb)
    return rb


def load_pg_agent(num_assets: int, window: int, features: int, seed: int = 0):
    """The reference's on-policy agent, `agent.pg.pg.PG(F)` with its LSRE-CANN policy (net/lsre_cann.py), seeded, eval mode."""
    _ensure_path()
    import torch
    import config.base as cb
    cb.NUM_ASSETS, cb.WINDOW_SIZE = int(num_assets), int(window)
    import net.lsre_cann as net
    import agent.pg.pg as pg
    importlib.reload(net)
    importlib.reload(pg)
    torch.manual_seed(seed)
    agent = pg.PG(features)
    agent.training_mode(False)
    return agent
