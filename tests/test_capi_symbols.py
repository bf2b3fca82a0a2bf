"""The C-ABI library loads and exports every symbol include/pmrl_b200.h declares (no compute calls: CPU-only)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pmrl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmrl_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for must in ("pmrl_env_reset", "pmrl_env_step", "pmrl_obs_build", "pmrl_ffd_weights", "pmrl_ffd_transform",
                 "pmrl_scale_series", "pmrl_pack_features", "pmrl_rollout_add", "pmrl_rollout_gather",
                 "pmrl_replay_add", "pmrl_replay_gather", "pmrl_pg_reward_fwd_bwd", "pmrl_eval_metrics",
                 "pmrl_env_step_io", "pmrl_env_step_burst", "pmrl_env_step_host", "pmrl_rollout_gather_index"):
        assert must in syms


def test_library_builds_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from pmrl_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} is declared in include/pmrl_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared_symbols(), "ctypes signature table and header diverged"
    assert _lib.load().pmrl_abi_version() == 5


def test_ctypes_structs_match_the_c_layout():
    """sizeof / field offsets of the structs that cross the boundary, as the C compiler sees them."""
    from pmrl_b200 import _lib
    lib = _lib.load()
    for which, ct in ((0, _lib.PmrlEnvCfg), (1, _lib.PmrlTables), (2, _lib.PmrlEnvState), (3, _lib.PmrlStepIO)):
        assert lib.pmrl_abi_sizeof(which) == ctypes.sizeof(ct), ct.__name__
    assert lib.pmrl_abi_sizeof(100) == _lib.PmrlStepIO.stats.offset
    assert lib.pmrl_abi_sizeof(101) == _lib.PmrlStepIO.done_host.offset
    assert lib.pmrl_abi_sizeof(102) == _lib.PmrlEnvState.ticket.offset
    assert lib.pmrl_abi_sizeof(103) == _lib.PmrlEnvCfg.initial_cash.offset


def test_argument_validation_needs_no_gpu():
    """Validation happens before any launch, so bad arguments are rejected even without a device."""
    from pmrl_b200 import _lib
    lib = _lib.load()
    cfg = _lib.PmrlEnvCfg(1, 0, 8, 5, 0, 0, 0, 16, 1, 25000.0, 0.0, 1.0, 0.04)          # A = 0
    st = _lib.PmrlEnvState(1, 1, 1, 1, 1, None, None, None, None)
    rc = lib.pmrl_env_step(ctypes.byref(cfg), None, ctypes.byref(st), 1, 1, 1, 1, None, 0, None, None)
    assert rc == -2 and b"A>=1" in lib.pmrl_last_error()
    assert lib.pmrl_ffd_weights(None, 1, 10, 1e-5, None, None, None) == -1
    assert lib.pmrl_eval_metrics(1, None, 1, 1, 1, 0.0, 252, 1, None) == -2


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pm-rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_env_requires_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import pmrl_b200
    from pmrl_b200._lib import PmrlError
    with pytest.raises(PmrlError):
        pmrl_b200.BatchedTradingEnv(pmrl_b200.EnvConfig())
