"""ctypes wrapper of oracle/pmrl_oracle.c (built by `make -C oracle`, also run by __graft_entry__.build()) — TEST /
BASELINE INFRASTRUCTURE.  Same state layout as oracle/env_oracle.OracleEnv (value [E], hist [E,W,A], idx, is_full, t)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(LIB)
        _lib.pmrl_oracle_step.restype = None
    return _lib


class COracleEnv:
    def __init__(self, E, A, W, episode_len=0, initial_cash=25000.0, commission=0.0, reward_scale=1.0, strict=True, mu_max_iter=16):
        self.E, self.A, self.W = E, A, W
        self.episode_len, self.cash, self.c, self.scale = episode_len, initial_cash, commission, reward_scale
        self.strict, self.mu_max_iter = strict, mu_max_iter
        self.value = np.full(E, initial_cash, np.float32)
        self.hist = np.zeros((E, W, A), np.float32); self.hist[:, 0, 0] = 1
        self.idx = np.ones(E, np.int32); self.is_full = np.zeros(E, np.uint8); self.t = np.zeros(E, np.int32)
        self.reward = np.zeros(E, np.float32); self.done = np.zeros(E, np.uint8)
        self._scratch = np.zeros((E, A), np.float32)
        self.lib = load()

    def step(self, actions, y):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.E, self.A)
        yy = np.ascontiguousarray(y, np.float32).reshape(self.E, self.A)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        self.lib.pmrl_oracle_step(C.c_int(self.E), C.c_int(self.A), C.c_int(self.W), p(a), p(yy), p(self.value), p(self.hist),
                                  p(self.idx), p(self.is_full), p(self.t), C.c_int(self.episode_len), C.c_float(self.cash),
                                  C.c_float(self.c), C.c_float(self.scale), C.c_int(1 if self.strict else 0),
                                  C.c_int(self.mu_max_iter), p(self.reward), p(self.done), p(self._scratch))
        return self.reward, self.done


def time_all_cores(A, W, E=32768, steps=20):
    """asset-steps/s of the C port with OpenMP over all cores (state-only transition, inputs resident in RAM)."""
    import time
    rs = np.random.RandomState(0)
    env = COracleEnv(E, A, W, episode_len=1000)
    acts = rs.standard_normal((2, E, A)).astype(np.float32)
    ys = (1 + 0.01 * rs.standard_normal((2, E, A))).astype(np.float32)
    env.step(acts[0], ys[0])
    t0 = time.perf_counter()
    for i in range(steps):
        env.step(acts[i & 1], ys[i & 1])
    return steps * E * A / (time.perf_counter() - t0)
