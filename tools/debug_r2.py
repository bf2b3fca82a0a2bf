import sys, traceback
import numpy as np, torch
sys.path.insert(0, '.')
import pmrl_b200
from pmrl_b200 import synth
from pmrl_b200.env import BatchedTradingEnv as Env
from oracle.env_oracle import OracleEnv

E, A, commission, W, L = 262144, 500, 0.0025, 50, 1000
T = 2048
tbl = synth.gbm_ohlc(T, A)
t0 = synth.episode_offsets(E, T, W, L)
cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, commission=commission, episode_len=L)
env = Env(cfg, prices=tbl, t0=t0)
env.reset(obs=False)
g = torch.Generator(device="cuda").manual_seed(1)
for s in range(6):
    act = torch.randn(E, A, generator=g, device="cuda")
    v_prev = env.value.clone()
    env.step(act, obs=False)
    soft = (act.sum(1) - 1.0).abs() > 1e-4
    bad = (~(env.value > 0)) & soft
    nb = int(bad.sum())
    print("step", s, "bad", nb, "nan", int(torch.isnan(env.value).sum()), "nonsoft", int((~soft).sum()))
    if nb:
        ids = bad.nonzero().flatten()[:8]
        print(" ids", ids.tolist(), "values", env.value[ids].tolist(), "prev", v_prev[ids].tolist())
        print(" warp-relative: id % (1184*8) =", (ids % (1184 * 8)).tolist(), " id // (1184*8) =", (ids // (1184 * 8)).tolist())
        small = Env(pmrl_b200.EnvConfig(num_envs=len(ids), num_assets=A, window_size=W, commission=commission, episode_len=L), prices=tbl, t0=t0[ids.cpu()])
        small.reset(obs=False)
        # replay the same steps for those envs
        g2 = torch.Generator(device="cuda").manual_seed(1)
        for s2 in range(s + 1):
            a2 = torch.randn(E, A, generator=g2, device="cuda")
            small.step(a2[ids].contiguous(), obs=False)
        print(" small-batch values", small.value.tolist())
        print(" act sums", act[ids].sum(1).tolist(), "act min", act[ids].min(1).values.tolist())
        break
