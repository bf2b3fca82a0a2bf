"""pmrl_b200 — B200-native portfolio-environment hot path of pm-rl (env step + reward + feature path).

Import name: `pmrl_b200` (the directory is `pm-rl_b200/`; `pmrl_b200/__init__.py` at the repo root maps
the importable name onto it).  Everything computes in hand-written sm_100a CUDA kernels reached through
the C-ABI of libpmrl_b200.so (include/pmrl_b200.h); there is no CPU fallback.
"""
from .config import EnvConfig  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["EnvConfig", "BatchedTradingEnv", "TradingEnv", "load_library"]


def load_library():
    """dlopen the in-tree libpmrl_b200.so (raises if it has not been built)."""
    return _lib.load()


def __getattr__(name):
    # torch-dependent modules are imported lazily so `import pmrl_b200` stays cheap
    if name == "BatchedTradingEnv":
        from .env import BatchedTradingEnv
        return BatchedTradingEnv
    if name == "TradingEnv":
        from .compat import TradingEnv
        return TradingEnv
    raise AttributeError(name)
