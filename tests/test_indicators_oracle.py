"""Known answers for the directional-movement restatement (oracle/indicators_oracle.py: dx, adx) — CPU only.
TA-Lib is not available (parity unpinned); these pin the restatement to cases whose value follows from the definition."""
import numpy as np

from oracle import indicators_oracle as io


def _bars(close, spread=1.0):
    c = np.asarray(close, np.float64)
    return c + spread, c - spread, c                      # high, low, close


def test_pure_trends_saturate_and_flat_markets_vanish():
    n, L = 5, 60
    up = _bars(100.0 + 2.0 * np.arange(L))                # only +DM: DX = 100 from the first defined bar
    down = _bars(300.0 - 2.0 * np.arange(L))              # only -DM
    flat = _bars(np.full(L, 50.0))                        # no directional movement at all: 0/0 → 0 by TA-Lib's convention
    for h, l, c in (up, down):
        d, a = io.dx(h, l, c, n), io.adx(h, l, c, n)
        assert np.isnan(d[:n]).all() and np.allclose(d[n:], 100.0)
        assert np.isnan(a[:2 * n - 1]).all() and np.allclose(a[2 * n - 1:], 100.0)
    d, a = io.dx(*flat, n), io.adx(*flat, n)
    assert (d[n:] == 0).all() and (a[2 * n - 1:] == 0).all()


def test_first_values_follow_the_wilder_recurrence_by_hand():
    # n = 2: bar 1 accumulates, bar 2 is the first Wilder step → DX[2]; ADX[3] = (DX[2] + DX[3]) / 2
    h = np.array([10.0, 11.0, 13.0, 12.5, 14.0]); l = np.array([9.0, 9.5, 11.0, 10.0, 12.0]); c = np.array([9.5, 10.5, 12.0, 11.0, 13.5])
    # bar1: +DM 1.0 (dp 1.0 > dm -0.5), TR = max(1.5, |11-9.5|, |9.5-9.5|) = 1.5
    # bar2: dp 2.0, dm -1.5 → +DM 2.0; TR = max(2.0, |13-10.5|, |11-10.5|) = 2.5 → +DM14 = 1 - 0.5 + 2 = 2.5, TR14 = 1.5 - 0.75 + 2.5 = 3.25, -DM = 0 → DX = 100
    # bar3: dp -0.5, dm 1.0 → -DM 1.0; TR = max(2.5, |12.5-12|, |10-12|) = 2.5 → +DM = 1.25, -DM = 1.0, TR = 4.125
    pdi, mdi = 100 * 1.25 / 4.125, 100 * 1.0 / 4.125
    dx3 = 100 * abs(mdi - pdi) / (pdi + mdi)
    d, a = io.dx(h, l, c, 2), io.adx(h, l, c, 2)
    assert np.isnan(d[:2]).all() and d[2] == np.float32(100.0) and np.isclose(d[3], dx3, rtol=1e-6)
    assert np.isnan(a[:3]).all() and np.isclose(a[3], (100.0 + dx3) / 2, rtol=1e-6)


def test_bounded_on_random_walks():
    rs = np.random.RandomState(4)
    for _ in range(5):
        c = 100 * np.exp(np.cumsum(rs.normal(0, 0.02, 300)))
        h = c * (1 + np.abs(rs.normal(0, 0.01, 300))); l = c * (1 - np.abs(rs.normal(0, 0.01, 300)))
        for n in (5, 14, 30):
            d, a = io.dx(h, l, c, n), io.adx(h, l, c, n)
            assert np.nanmin(d) >= 0 and np.nanmax(d) <= 100 + 1e-4 and np.nanmin(a) >= 0 and np.nanmax(a) <= 100 + 1e-4
