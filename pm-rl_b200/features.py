"""Feature path on device: fractional differencing (data/ffd.py), per-series scaling (data/instrument.py:318-336)
and packing into the table layouts the env kernels gather from (data/instrument.py:339-356).

`FixedFracDiff` keeps the reference class surface (`FixedFracDiff(data, thres, d_opt).fit_transform()`,
`get_max_width()`, `get_d_opt`); its `fit` (bisection on a statsmodels ADF p-value, ffd.py:59-78) is host-side,
one-time and out of scope — `d_opt` must be supplied.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

SCALERS = {"minmax": 0, "standard": 1}

# features the reference leaves untouched (data/ffd.py:21)
FEAT_IGNORE = ['volume', 'adx', 'adxr', 'apo', 'aroon', 'aroonosc', 'bop', 'cci', 'cmo', 'dx', 'macd', 'macdext', 'macdfix',
               'mfi', 'minus_di', 'minus_dm', 'mom', 'plus_di', 'plus_dm', 'ppo', 'roc', 'rocp', 'rocr', 'rocr100', 'rsi',
               'stoch', 'stochf', 'stochrsi', 'trix', 'ultosc', 'willr']


def _cuda_f32(x, device=None):
    t = torch.as_tensor(x)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.PmrlError("the feature path needs a CUDA device (pmrl_b200 has no CPU fallback)")
        t = t.to(device or "cuda")
    return t.to(torch.float32).contiguous()


def ffd_weights(d, T: int, thres: float, device=None):
    """Binomial weights + widths per series (ffd.py:38-47).  d: [N] host floats (kept in fp64 like Python floats)."""
    lib = _lib.load()
    d64 = torch.tensor(np.asarray(d, np.float64), dtype=torch.float64).to(device or "cuda")
    N = d64.numel()
    w = torch.empty(N, T, dtype=torch.float32, device=d64.device)
    widths = torch.empty(N, dtype=torch.int32, device=d64.device)
    _lib.check(lib.pmrl_ffd_weights(_lib.ptr(d64), N, T, float(thres), _lib.ptr(w), _lib.ptr(widths),
                                    _lib.current_stream()), "pmrl_ffd_weights")
    return w, widths, d64


def ffd_transform(x, d, thres: float = 1e-5):
    """FixedFracDiff.transform for a stack of series (ffd.py:80-89).  x: [N, T]; d: [N].
    Returns (out [N, T - max_width], widths [N] (host ints), max_width).  One host read of the widths sizes the output."""
    lib = _lib.load()
    x = _cuda_f32(x)
    N, T = x.shape
    w, widths, d64 = ffd_weights(d, T, thres, device=x.device)
    wh = widths.cpu().numpy()
    dh = np.asarray(d, np.float64)
    mw = int(np.where(dh > 0, wh, 0).max()) if N else 0
    out = torch.empty(N, T - mw, dtype=torch.float32, device=x.device)
    _lib.check(lib.pmrl_ffd_transform(_lib.ptr(x), _lib.ptr(d64), _lib.ptr(w), _lib.ptr(widths), N, T, mw,
                                      _lib.ptr(out), _lib.current_stream()), "pmrl_ffd_transform")
    return out, np.where(dh > 0, wh, 0).astype(np.int32), mw


def scale_series(x, method: str = "minmax", out=None):
    """Instrument.scale per series over its whole length (instrument.py:331-336): sklearn MinMax / Standard semantics."""
    lib = _lib.load()
    x = _cuda_f32(x)
    if out is None:
        out = torch.empty_like(x)
    N, L = x.shape
    _lib.check(lib.pmrl_scale_series(_lib.ptr(x), N, L, SCALERS[method], _lib.ptr(out), _lib.current_stream()),
               "pmrl_scale_series")
    return out


def pack_tables(series=None, close=None, num_assets: int = 0, channels: int = 0):
    """series [A*C, L] (index a*C + c) → feat_am [A, L, C]; close [A, L] → close_tm [L, A] (env kernel layouts)."""
    lib = _lib.load()
    feat_am = close_tm = None
    A, C, L = num_assets, channels, 0
    if series is not None:
        series = _cuda_f32(series)
        L = series.shape[1]
        if series.shape[0] != A * C:
            raise ValueError(f"series must be [{A * C}, L]")
        feat_am = torch.empty(A, L, C, dtype=torch.float32, device=series.device)
    if close is not None:
        close = _cuda_f32(close)
        A = A or close.shape[0]
        L = close.shape[1]
        close_tm = torch.empty(L, A, dtype=torch.float32, device=close.device)
    _lib.check(lib.pmrl_pack_features(_lib.ptr(series), _lib.ptr(close), A, max(C, 1), L, _lib.ptr(feat_am),
                                      _lib.ptr(close_tm), _lib.current_stream()), "pmrl_pack_features")
    return feat_am, close_tm


def build_env_tables(ohlc, d=0.4, thres: float = 1e-5, scaler: str | None = "minmax", close_channel: int = 3,
                     indicators=None, d_indicators=None, feature_channels: int = 4):
    """The reference pipeline indicators → FFD → scale → window layout (data/data_loader.py:37-43) for an OHLC(V) table
    [T, A, C]:

      1. `indicators` (config.base.INDICATORS form, see `add_indicators`): the indicator windows are computed from the raw
         table and appended as feature channels; every series is clipped by the largest lookback (instrument.py:207-232);
      2. every (asset, channel) series is fractionally differenced — the first `feature_channels` table channels with `d`
         (scalar, [C'] or [A, C']), the indicator outputs with `d_indicators` (scalar or one value per output; default: `d`
         when scalar, else 0) — except the features data/ffd.py:21 ignores BY NAME (the reference compares the full feature
         name, so 'volume' is ignored while a suffixed indicator name such as 'rsi_30' is differenced like any other);
      3. scaled over the whole split (instrument.py:318-336) and packed to feat_am [A, T', F-1]; the raw close plane (for
         y_t) is row-aligned to it → close_tm [T', A].

    F - 1 = feature_channels + n_out: with F - 1 a multiple of four (OHLC + ema + bbands → F = 9) the fused step+obs kernel
    consumes the table directly.  Returns dict(feat_am, close_tm, max_width, lookback, rows, widths, names)."""
    tbl = _cuda_f32(ohlc)
    T, A, C = tbl.shape
    Cf = min(int(feature_channels), C)
    base_names = ["open", "high", "low", "close", "volume"][:Cf] if Cf <= 5 else [f"ch{i}" for i in range(Cf)]
    names, lb, n_out = list(base_names), 0, 0
    series = tbl[:, :, :Cf].permute(1, 2, 0).contiguous()                # [A, Cf, T] series-major (one-time setup copy)
    if indicators:
        ind_names, ind, lb = add_indicators(tbl, indicators)               # [A, n_out, T - lb]
        n_out = ind.shape[1]
        names += ind_names
        series = torch.cat([series[:, :, lb:], ind], dim=1).contiguous()    # [A, Cf + n_out, T - lb]
    Ct = Cf + n_out
    Tl = T - lb
    dd = np.zeros((A, Ct), np.float64)
    dd[:, :Cf] = np.broadcast_to(np.asarray(d, np.float64), (A, Cf))
    if n_out:
        if d_indicators is None:
            d_indicators = float(d) if np.ndim(d) == 0 else 0.0
        dd[:, Cf:] = np.broadcast_to(np.asarray(d_indicators, np.float64), (A, n_out))
    for c, name in enumerate(names):
        if name in FEAT_IGNORE:
            dd[:, c] = 0.0
    out, widths, mw = ffd_transform(series.view(A * Ct, Tl), dd.reshape(-1), thres)
    if scaler is not None:
        out = scale_series(out, scaler, out=out)
    close = tbl[lb + mw:, :, close_channel].t().contiguous()             # [A, T']
    feat_am, close_tm = pack_tables(out, close, num_assets=A, channels=Ct)
    return {"feat_am": feat_am, "close_tm": close_tm, "max_width": mw, "lookback": lb, "rows": Tl - mw, "widths": widths,
            "names": names}


class FixedFracDiff:
    """data/ffd.py:6-101 surface over the CUDA kernels.  `data`: mapping feature name → 1-D tensor [T]."""

    def __init__(self, data, thres: float = 1e-5, d_opt: dict | None = None):
        self.data = data
        self.thres = thres
        self.len_data = len(next(iter(data.values())))
        self.feats = [f for f in data.keys() if f not in FEAT_IGNORE]
        self.d_opt = {f: 0.0 for f in self.feats} if d_opt is None else dict(d_opt)
        self.widths = [0] * len(self.feats)
        self._out = None

    def fit(self):
        missing = [f for f in self.feats if not self.d_opt.get(f, 0.0) > 0.0]
        if missing:
            raise NotImplementedError("FixedFracDiff.fit searches d with a statsmodels ADF test (data/ffd.py:59-78), which is "
                                      f"host-side and out of scope: supply d_opt for {missing}")

    def transform(self):
        x = torch.stack([torch.as_tensor(self.data[f]).float() for f in self.feats])
        d = [float(self.d_opt.get(f, 0.0)) for f in self.feats]
        out, widths, mw = ffd_transform(x, d, self.thres)
        self.widths = [int(w) for w in widths]
        res = {k: torch.as_tensor(v)[mw:].clone() for k, v in self.data.items()}
        host = out.cpu()
        for i, f in enumerate(self.feats):
            if d[i] > 0:
                res[f] = host[i]
        self._out = res
        return res

    def fit_transform(self):
        self.fit()
        return self.transform()

    def get_max_width(self) -> int:
        return int(max(self.widths))

    @property
    def get_d_opt(self) -> dict:
        return self.d_opt


# ---------------------------------------------------------------------------------------------------------------
# Indicator windows (data/instrument.py:207-232; config/base.py:30-44)
# ---------------------------------------------------------------------------------------------------------------
INDICATOR_KINDS = {"sma": 0, "ema": 1, "rsi": 2, "atr": 3, "bbands": 4, "macd": 5, "obv": 6, "adosc": 7, "cci": 8, "stoch": 9, "dx": 10, "adx": 11}
_IND_OUTPUTS = {"sma": ["sma"], "ema": ["ema"], "rsi": ["rsi"], "atr": ["atr"],
                "bbands": ["upperband", "middleband", "lowerband"], "macd": ["macd", "macdsignal", "macdhist"],
                "obv": ["obv"], "adosc": ["adosc"], "cci": ["cci"], "stoch": ["slowk", "slowd"], "dx": ["dx"], "adx": ["adx"]}
_IND_DEFAULT_PERIOD = {"sma": 30, "ema": 30, "rsi": 14, "atr": 14, "bbands": 5, "macd": 0, "obv": 0, "adosc": 0, "cci": 14, "stoch": 0, "dx": 14, "adx": 14}


def add_indicators(ohlcv, indicators):
    """Instrument.add_indicators batched over assets (instrument.py:207-232).

    ohlcv: [T, A, C>=4] (o, h, l, c[, v]); indicators: list of (name, params-dict) or a dict {name: params} like
    config.base.INDICATORS (a list allows the same indicator with several periods).
    Returns (names, out [A, n_out, T - lookback] float32, lookback): the outputs are clipped by the maximum lookback
    exactly like the reference clips the instrument (`self.clip(start=lookback)`, :231)."""
    import ctypes as C
    lib = _lib.load()
    items = list(indicators.items()) if isinstance(indicators, dict) else list(indicators)
    tbl = _cuda_f32(ohlcv)
    T, A, Cn = tbl.shape
    specs, names = [], []
    for name, params in items:
        if name not in INDICATOR_KINDS:
            raise NotImplementedError(f"indicator {name!r} is not implemented on device (have: {sorted(INDICATOR_KINDS)})")
        period = int((params or {}).get("timeperiod", _IND_DEFAULT_PERIOD[name]))
        specs += [INDICATOR_KINDS[name], period]
        suffix = "_".join(str(v) for v in (params or {}).values())
        names += [f"{o if o != name else name}_{suffix}" for o in _IND_OUTPUTS[name]]       # naming of :228-229
    n = len(items)
    arr = (C.c_int32 * (2 * n))(*specs)
    n_out, lb = C.c_int32(0), C.c_int32(0)
    _lib.check(lib.pmrl_indicator_layout(C.cast(arr, C.c_void_p), n, C.cast(C.byref(n_out), C.c_void_p),
                                         C.cast(C.byref(lb), C.c_void_p)), "pmrl_indicator_layout")
    series = tbl.permute(1, 2, 0).contiguous().view(A * Cn, T)
    out = torch.empty(A, max(n_out.value, 1), T, dtype=torch.float32, device=tbl.device)
    _lib.check(lib.pmrl_indicators(_lib.ptr(series), A, Cn, T, C.cast(arr, C.c_void_p), n, _lib.ptr(out), _lib.current_stream()),
               "pmrl_indicators")
    return names, out[:, :n_out.value, lb.value:].contiguous(), lb.value
