"""CPU oracle of the feature path — TEST INFRASTRUCTURE (see oracle/env_oracle.py header).

Restates data/ffd.py:38-53,81-88 (fixed-width-window fractional differencing), data/instrument.py:318-336
(per-series scaling through sklearn) and data/instrument.py:339-356 (windowing).

Parity status:
  * FFD transform: PINNED — tests/golden/ffd.npz holds outputs of the reference's own data/ffd.py executed with
    import stubs for its two missing third-party modules (`tensordict` → a dict-of-tensors shim, `statsmodels`
    → an `adfuller` stub; neither touches the transform arithmetic) by tests/golden/make_golden_ffd.py.
  * FFD fit (bisection on the ADF p-value, ffd.py:59-78): OUT OF SCOPE / UNPINNED — statsmodels is not vendored;
    `d` is an input here.
  * scaling: pinned by running sklearn (present in this image) in tests/test_features.py.
"""
from __future__ import annotations

import numpy as np
import torch


def ffd_weights(d: float, length: int, thres: float):
    """ffd.py:38-47 — binomial weights by sequential fp32 cumprod, width = last index above the threshold,
    taps = w[:width] (the tap AT `width` is dropped: quirk Q9)."""
    k = torch.arange(1, length, dtype=torch.float32)
    factor = torch.zeros(length - 1) - (d + 1)
    proto = torch.cat([torch.tensor([1.0]), torch.div(factor, k) + 1])
    w = torch.cumprod(proto, dim=0)
    width = int(torch.where(w.abs() > thres)[0].max())
    return w, width


def ffd_series(x: torch.Tensor, d: float, thres: float):
    """ffd.py:49-53 — valid conv1d with the flipped truncated taps → [len - width + 1]."""
    w, width = ffd_weights(d, x.numel(), thres)
    taps = w[:width].flip(0)
    diff = torch.nn.functional.conv1d(x.reshape(1, 1, -1).float(), taps.reshape(1, 1, -1), padding=0).squeeze()
    return diff, width


def ffd_transform(x: np.ndarray, d: np.ndarray, thres: float):
    """ffd.py:80-89 — all series tail-aligned to len - max_width; series with d <= 0 are passed through."""
    x = torch.as_tensor(np.asarray(x, np.float32))
    N, T = x.shape
    diffs, widths = [], []
    for n in range(N):
        if d[n] > 0:
            df, wd = ffd_series(x[n], float(d[n]), thres)
        else:
            df, wd = None, 0
        diffs.append(df); widths.append(wd)
    mw = int(max(widths))
    out = torch.empty(N, T - mw)
    for n in range(N):
        out[n] = diffs[n][-(T - mw):] if diffs[n] is not None else x[n, mw:]
    return out.numpy(), np.asarray(widths, np.int32), mw


def ffd_transform_f64(x: np.ndarray, d: np.ndarray, thres: float):
    """Same filter with fp64 accumulation over the fp32 taps (accuracy yardstick for the fp32 kernels)."""
    x64 = np.asarray(x, np.float64)
    N, T = x64.shape
    ws = [ffd_weights(float(d[n]), T, thres) if d[n] > 0 else (None, 0) for n in range(N)]
    mw = max(w[1] for w in ws)
    out = np.empty((N, T - mw))
    for n, (w, wd) in enumerate(ws):
        if w is None:
            out[n] = x64[n, mw:]
        else:
            taps = w[:wd].double().numpy()
            full = np.convolve(x64[n], taps, mode="valid")            # len T - wd + 1, index t ↔ x[t + wd - 1 - k]
            out[n] = full[-(T - mw):]
    return out


def scale_series(x: np.ndarray, method: str = "minmax"):
    """instrument.py:331-336 through sklearn, per series over its whole length."""
    from sklearn.preprocessing import MinMaxScaler, StandardScaler
    x = np.asarray(x, np.float32)
    out = np.empty_like(x)
    for n in range(x.shape[0]):
        sc = MinMaxScaler() if method == "minmax" else StandardScaler()
        out[n] = sc.fit_transform(x[n].reshape(-1, 1)).flatten()
    return out


def window(series: np.ndarray, W: int):
    """instrument.py:351-353 — features_[i] = features[i:i+W] → [L-W+1, W] per series."""
    L = series.shape[-1]
    return np.stack([series[..., i:i + W] for i in range(L - W + 1)], axis=-2)
