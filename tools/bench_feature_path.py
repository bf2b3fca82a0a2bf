"""Feature-path / off-policy micro-benchmarks (BASELINE config 3): FFD table build (K3), step + index-replay write,
replay gather (K5).  Prints JSON lines; results are recorded under profiles/."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmrl_b200
from pmrl_b200 import features, synth
from pmrl_b200.buffers import DeviceReplayBuffer
from pmrl_b200.env import BatchedTradingEnv


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    A, W, E, L = 100, 50, 65536, 1000
    # ---- K3: FFD(d = 0.4, thres 1e-5) + min-max + packing of 400 series x 12,000 rows ----
    T = 12000
    tbl = synth.gbm_ohlc(T, A).cuda()
    series = tbl.permute(1, 2, 0).contiguous().view(A * 4, T)
    d = [0.4] * (A * 4)
    out, widths, mw = features.ffd_transform(series, d, 1e-5)
    ms = timeit(lambda: features.ffd_transform(series, d, 1e-5), n=5)
    flops = 2.0 * sum(int(w) for w in widths) * (T - mw)
    print(json.dumps({"bench": "ffd_transform", "series": A * 4, "rows": T, "d": 0.4, "thres": 1e-5, "width": int(widths[0]),
                      "max_width": mw, "ms": ms, "gflops": flops / ms / 1e6, "note": "weights + conv (+ one host read of the widths)"}))
    from pmrl_b200 import _lib
    lib = _lib.load()
    wts, wd, d64 = features.ffd_weights(d, T, 1e-5)
    outb = torch.empty(A * 4, T - mw, device="cuda")
    ms_w = timeit(lambda: features.ffd_weights(d, T, 1e-5), n=5)
    ms_cv = timeit(lambda: _lib.check(lib.pmrl_ffd_transform(series.data_ptr(), d64.data_ptr(), wts.data_ptr(), wd.data_ptr(), A * 4, T, mw,
                                                             outb.data_ptr(), _lib.current_stream()), "ffd"), n=10)
    print(json.dumps({"bench": "k_ffd_weights (400 series x 12,000 factors, sequential double cumprod)", "ms": ms_w}))
    print(json.dumps({"bench": "k_ffd_conv alone", "ms": ms_cv, "gflops": flops / ms_cv / 1e6, "fp32_peak_note": "B200 fp32 FMA peak ~75 TFLOP/s"}))
    ms_tab = timeit(lambda: features.build_env_tables(tbl, d=0.4, thres=1e-5, scaler="minmax"), n=5)
    print(json.dumps({"bench": "build_env_tables(ffd + minmax + pack)", "ms": ms_tab}))
    # ---- config 3: off-policy collect on FFD features: step (obs) + index-replay write ----
    tabs = features.build_env_tables(tbl, d=0.4, thres=1e-5, scaler="minmax")
    rows = tabs["rows"]
    cfg = pmrl_b200.EnvConfig(num_envs=E, num_assets=A, window_size=W, episode_len=L)
    t0 = synth.episode_offsets(E, rows, W, L)
    env = BatchedTradingEnv.from_tables(cfg, tabs["close_tm"], tabs["feat_am"], t0=t0)
    rb = DeviceReplayBuffer(tabs["feat_am"], 5, 128 + 2 * (W - 1), E, A, W, buffer_size=128, batch_size=64)
    acts = [torch.randn(E, A, device="cuda") for _ in range(4)]
    env.reset()
    for i in range(W):
        env.step(acts[i % 4], obs=False)
    step_i = [2 * (W - 1)]

    def collect():
        a = acts[step_i[0] % 4]
        _, r, _ = env.step(a)
        rb.add(0, step_i[0], a, r)
        step_i[0] = 2 * (W - 1) + (step_i[0] + 1 - 2 * (W - 1)) % rb.epoch_len

    ms_c = timeit(collect, n=20)
    print(json.dumps({"bench": "config3 collect: fused step(obs) + replay add", "envs": E, "assets": A, "ms_per_step": ms_c,
                      "asset_steps_per_s": E * A / ms_c * 1e3}))
    g = torch.Generator().manual_seed(0)
    ms_g = timeit(lambda: rb.sample(g, sampler="buffer"), n=20)
    print(json.dumps({"bench": "replay sample (B = 64: s, a, r, s')", "ms": ms_g,
                      "GBps": 2 * 64 * A * W * 5 * 4 / ms_g / 1e6}))


if __name__ == "__main__":
    main()
