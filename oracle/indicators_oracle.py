"""CPU oracle of the indicator windows — TEST INFRASTRUCTURE (see oracle/env_oracle.py header).

The reference computes indicators with TA-Lib's abstract functions on float64 OHLCV (data/instrument.py:207-232).
TA-Lib is neither vendored nor pinned (it is not even in requirements.txt) and is absent from this image → parity
UNPINNED: these functions restate TA-Lib's published algorithms (SMA-seeded EMA, Wilder-smoothed RSI / ATR,
population-std Bollinger bands, 12/26/9 MACD with both EMAs seeded at index 25) in float64, outputs float32 with NaN
over the lookback, exactly the conventions `abstract.Function.run` has."""
from __future__ import annotations

import numpy as np


def sma(c, n):
    c = np.asarray(c, np.float64); out = np.full(c.shape, np.nan)
    cs = np.cumsum(np.insert(c, 0, 0.0))
    out[n - 1:] = (cs[n:] - cs[:-n]) / n
    return out.astype(np.float32)


def ema(c, n):
    c = np.asarray(c, np.float64); out = np.full(c.shape, np.nan); k = 2.0 / (n + 1.0)
    e = c[:n].mean(); out[n - 1] = e
    for t in range(n, len(c)):
        e = (c[t] - e) * k + e; out[t] = e
    return out.astype(np.float32)


def rsi(c, n):
    c = np.asarray(c, np.float64); out = np.full(c.shape, np.nan)
    d = np.diff(c); g = np.maximum(d, 0); l = np.maximum(-d, 0)
    ag, al = g[:n].mean(), l[:n].mean()
    out[n] = 100 * ag / (ag + al) if ag + al != 0 else 0.0
    for t in range(n + 1, len(c)):
        ag = (ag * (n - 1) + g[t - 1]) / n; al = (al * (n - 1) + l[t - 1]) / n
        out[t] = 100 * ag / (ag + al) if ag + al != 0 else 0.0
    return out.astype(np.float32)


def atr(h, l, c, n):
    h, l, c = (np.asarray(x, np.float64) for x in (h, l, c)); out = np.full(c.shape, np.nan)
    tr = np.maximum(h[1:] - l[1:], np.maximum(np.abs(h[1:] - c[:-1]), np.abs(l[1:] - c[:-1])))
    a = tr[:n].mean(); out[n] = a
    for t in range(n + 1, len(c)):
        a = (a * (n - 1) + tr[t - 1]) / n; out[t] = a
    return out.astype(np.float32)


def bbands(c, n):
    c = np.asarray(c, np.float64); up = np.full(c.shape, np.nan); mid = up.copy(); lo = up.copy()
    for t in range(n - 1, len(c)):
        w = c[t - n + 1:t + 1]; m = w.mean(); sd = np.sqrt(max((w * w).mean() - m * m, 0.0))
        up[t], mid[t], lo[t] = m + 2 * sd, m, m - 2 * sd
    return up.astype(np.float32), mid.astype(np.float32), lo.astype(np.float32)


def macd(c):
    c = np.asarray(c, np.float64); L = len(c)
    m = np.full(L, np.nan); sg = m.copy(); hs = m.copy()
    kf, ks, kg = 2 / 13.0, 2 / 27.0, 2 / 10.0
    es, ef = c[:26].mean(), c[14:26].mean()
    raw = [ef - es]
    for t in range(26, L):
        es = (c[t] - es) * ks + es; ef = (c[t] - ef) * kf + ef; raw.append(ef - es)
    raw = np.array(raw)
    eg = raw[:9].mean()
    for q in range(8, len(raw)):
        if q > 8:
            eg = (raw[q] - eg) * kg + eg
        t = q + 25
        m[t], sg[t], hs[t] = raw[q], eg, raw[q] - eg
    return m.astype(np.float32), sg.astype(np.float32), hs.astype(np.float32)


def obv(c, v):
    c, v = np.asarray(c, np.float64), np.asarray(v, np.float64); out = np.empty(len(c)); out[0] = v[0]
    for t in range(1, len(c)):
        out[t] = out[t - 1] + (v[t] if c[t] > c[t - 1] else -v[t] if c[t] < c[t - 1] else 0.0)
    return out.astype(np.float32)


def adosc(h, l, c, v):
    h, l, c, v = (np.asarray(x, np.float64) for x in (h, l, c, v)); out = np.full(len(c), np.nan)
    ad = 0.0; kf, ks = 2 / 4.0, 2 / 11.0
    for t in range(len(c)):
        rng = h[t] - l[t]
        if rng > 0:
            ad += ((c[t] - l[t]) - (h[t] - c[t])) / rng * v[t]
        if t == 0:
            ef = es = ad
        else:
            ef = (ad - ef) * kf + ef; es = (ad - es) * ks + es
        if t >= 9:
            out[t] = ef - es
    return out.astype(np.float32)


def cci(h, l, c, n):
    tp = (np.asarray(h, np.float64) + np.asarray(l, np.float64) + np.asarray(c, np.float64)) / 3.0
    out = np.full(len(tp), np.nan)
    for t in range(n - 1, len(tp)):
        w = tp[t - n + 1:t + 1]; m = w.mean(); md = np.abs(w - m).mean()
        out[t] = (tp[t] - m) / (0.015 * md) if md != 0 else 0.0
    return out.astype(np.float32)


def stoch(h, l, c):
    h, l, c = (np.asarray(x, np.float64) for x in (h, l, c)); L = len(c)
    fk = np.full(L, np.nan)
    for t in range(4, L):
        hh, ll = h[t - 4:t + 1].max(), l[t - 4:t + 1].min()
        fk[t] = 100 * (c[t] - ll) / (hh - ll) if hh != ll else 0.0
    sk = np.full(L, np.nan); sd = np.full(L, np.nan)
    for t in range(6, L):
        sk[t] = fk[t - 2:t + 1].mean()
    for t in range(8, L):
        sd[t] = sk[t - 2:t + 1].mean()
    sk[:8] = np.nan
    return sk.astype(np.float32), sd.astype(np.float32)


def _dm_tr(h, l, c):
    h, l, c = (np.asarray(x, np.float64) for x in (h, l, c))
    dp, dm = h[1:] - h[:-1], l[:-1] - l[1:]
    plus = np.where((dp > 0) & (dp > dm), dp, 0.0); minus = np.where((dm > 0) & (dp < dm), dm, 0.0)
    tr = np.maximum(h[1:] - l[1:], np.maximum(np.abs(h[1:] - c[:-1]), np.abs(l[1:] - c[:-1])))
    return plus, minus, tr                                   # element q belongs to bar q + 1


def dx(h, l, c, n, _with_flags=False):
    """ta_DX.c: bars 1..n-1 accumulate, from bar n Wilder smoothing; a zero denominator repeats the previous value."""
    plus, minus, tr = _dm_tr(h, l, c); L = len(tr) + 1
    out = np.full(L, np.nan); have = np.zeros(L, bool)
    p, m, t_ = plus[:n - 1].sum(), minus[:n - 1].sum(), tr[:n - 1].sum()
    prev = 0.0
    for t in range(n, L):
        p = p - p / n + plus[t - 1]; m = m - m / n + minus[t - 1]; t_ = t_ - t_ / n + tr[t - 1]
        if t_ != 0:
            pdi, mdi = 100 * (p / t_), 100 * (m / t_)
            if pdi + mdi != 0:
                prev = 100 * (abs(mdi - pdi) / (pdi + mdi)); have[t] = True
        out[t] = prev
    return (out, have) if _with_flags else out.astype(np.float32)


def adx(h, l, c, n):
    """ta_ADX.c: mean of the first n DX values at index 2n-1, then (ADX*(n-1) + DX)/n; undefined DX bars add nothing."""
    d, have = dx(h, l, c, n, _with_flags=True); L = len(d)
    out = np.full(L, np.nan)
    a = sum(d[t] for t in range(n, 2 * n) if have[t]) / n
    out[2 * n - 1] = a
    for t in range(2 * n, L):
        if have[t]:
            a = (a * (n - 1) + d[t]) / n
        out[t] = a
    return out.astype(np.float32)
