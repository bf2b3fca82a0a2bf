"""Shared helpers for the test-suite: golden fixture loading and tolerances."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north-star tolerances (BASELINE.json): values / weights within 1e-5 relative in fp32 over 1,000 steps;
# rewards use a mixed bound because relative error is meaningless near r == 0.
RTOL_VALUE = 1e-5
ATOL_WEIGHT = 1e-6
RTOL_WEIGHT = 1e-5
RTOL_REWARD = 1e-5
ATOL_REWARD = 1e-6


def env_fixture_names():
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "env_*.npz")))


def load_env_fixture(name):
    """Returns the fixture as a dict with `actions` / `y` present (regenerated from the seed and checked
    against the stored sha256 when the fixture does not carry them)."""
    from tests.golden.make_golden import input_hash, make_inputs
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    for k in ("A", "W", "F", "S", "seed", "weights_every"):
        d[k] = int(d[k])
    d["kind"] = str(d["kind"])
    d["commission"] = float(d["commission"])
    if "actions" not in d:
        act, y = make_inputs(d["seed"], d["kind"], d["S"], d["A"])
        assert input_hash(act, y) == str(d["input_sha256"]), "seeded input stream drifted from the fixture"
        d["actions"], d["y"] = act, y
    return d


def weight_steps(d):
    S, every = d["S"], d["weights_every"]
    return [s for s in range(S) if s % every == 0 or s == S - 1]


def assert_rewards_close(got, want, msg=""):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    bound = RTOL_REWARD * np.abs(want) + ATOL_REWARD
    bad = ~(np.abs(got - want) <= bound) & ~(np.isnan(got) & np.isnan(want))
    assert not bad.any(), f"{msg} reward mismatch at {np.argwhere(bad)[:5].tolist()}: got {got[bad][:5]} want {want[bad][:5]}"


def assert_values_close(got, want, msg="", rtol=RTOL_VALUE):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    rel = np.abs(got - want) / np.abs(want)
    assert np.all(rel <= rtol), f"{msg} value mismatch: max rel {rel.max():.3e} at {int(rel.argmax())}"
