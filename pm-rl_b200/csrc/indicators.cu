// indicators.cu — technical-indicator windows of the feature path (data/instrument.py:207-232 runs TA-Lib's abstract
// functions on the float64 OHLCV and appends the float32 outputs; the indicator set is config/base.py:30-44).
// TA-Lib itself is not vendored with the reference (no version pin) → the formulas below restate TA-Lib's published
// algorithms (ta_SMA.c, ta_EMA.c, ta_RSI.c, ta_ATR.c, ta_BBANDS.c, ta_MACD.c, ta_DX.c, ta_ADX.c …; default unstable period 0) in double
// precision like TA-Lib computes them; parity against TA-Lib is UNPINNED.
//
// One thread per (asset, indicator): every indicator is a sequential scan over the series (one-time preprocessing).
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "pmrl_b200.h"
#include "host_util.h"

namespace pmrl {

constexpr int kMaxSpecs = 16;
struct IndSpecs { int n; int kind[kMaxSpecs]; int period[kMaxSpecs]; int out0[kMaxSpecs]; };

__device__ __forceinline__ float nanf32() { return __int_as_float(0x7fc00000); }

__global__ void k_indicators(const float* __restrict__ series, int A, int C, int L, IndSpecs sp, int n_out,
                             float* __restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (a >= A || s >= sp.n) return;
    const float* __restrict__ hi = series + ((size_t)a * C + 1) * L;
    const float* __restrict__ lo = series + ((size_t)a * C + 2) * L;
    const float* __restrict__ cl = series + ((size_t)a * C + 3) * L;
    const float* __restrict__ vo = series + ((size_t)a * C + (C > 4 ? 4 : 3)) * L;   // volume (needs C >= 5)
    float* __restrict__ o0 = out + ((size_t)a * n_out + sp.out0[s]) * L;
    float* __restrict__ o1 = o0 + L;
    float* __restrict__ o2 = o1 + L;
    const int n = sp.period[s];
    const float nan = nanf32();
    switch (sp.kind[s]) {
    case PMRL_IND_SMA: {
        double sum = 0.0;
        for (int t = 0; t < L; ++t) {
            sum += (double)cl[t];
            if (t >= n) sum -= (double)cl[t - n];
            o0[t] = t >= n - 1 ? (float)(sum / n) : nan;
        }
    } break;
    case PMRL_IND_EMA: {
        const double k = 2.0 / (n + 1.0);
        double sum = 0.0, ema = 0.0;
        for (int t = 0; t < L; ++t) {
            const double x = (double)cl[t];
            if (t < n - 1) { sum += x; o0[t] = nan; }
            else if (t == n - 1) { sum += x; ema = sum / n; o0[t] = (float)ema; }            // seeded with the SMA
            else { ema = (x - ema) * k + ema; o0[t] = (float)ema; }
        }
    } break;
    case PMRL_IND_RSI: {                                                                      // Wilder smoothing
        double ag = 0.0, al = 0.0;
        if (L > 0) o0[0] = nan;
        for (int t = 1; t < L; ++t) {
            const double d = (double)cl[t] - (double)cl[t - 1];
            const double g = d > 0 ? d : 0.0, l = d < 0 ? -d : 0.0;
            if (t <= n) {
                ag += g; al += l;
                if (t == n) { ag /= n; al /= n; const double den = ag + al; o0[t] = (float)(den != 0.0 ? 100.0 * ag / den : 0.0); }
                else o0[t] = nan;
            } else {
                ag = (ag * (n - 1) + g) / n; al = (al * (n - 1) + l) / n;
                const double den = ag + al;
                o0[t] = (float)(den != 0.0 ? 100.0 * ag / den : 0.0);
            }
        }
    } break;
    case PMRL_IND_ATR: {                                                                      // Wilder smoothing of the true range
        double atr = 0.0;
        if (L > 0) o0[0] = nan;
        for (int t = 1; t < L; ++t) {
            const double h = hi[t], l = lo[t], pc = cl[t - 1];
            const double tr = fmax(h - l, fmax(fabs(h - pc), fabs(l - pc)));
            if (t <= n) { atr += tr; if (t == n) { atr /= n; o0[t] = (float)atr; } else o0[t] = nan; }
            else { atr = (atr * (n - 1) + tr) / n; o0[t] = (float)atr; }
        }
    } break;
    case PMRL_IND_BBANDS: {                                                                   // SMA ± 2 population std
        double sum = 0.0, sq = 0.0;
        for (int t = 0; t < L; ++t) {
            const double x = (double)cl[t];
            sum += x; sq += x * x;
            if (t >= n) { const double y = (double)cl[t - n]; sum -= y; sq -= y * y; }
            if (t >= n - 1) {
                const double m = sum / n;
                double var = sq / n - m * m;
                if (var < 0.0) var = 0.0;
                const double sd = sqrt(var);
                o0[t] = (float)(m + 2.0 * sd); o1[t] = (float)m; o2[t] = (float)(m - 2.0 * sd);
            } else { o0[t] = nan; o1[t] = nan; o2[t] = nan; }
        }
    } break;
    case PMRL_IND_MACD: {                                                                     // 12 / 26 / 9
        const int nf = 12, ns = 26, ng = 9;
        const double kf = 2.0 / (nf + 1.0), ks = 2.0 / (ns + 1.0), kg = 2.0 / (ng + 1.0);
        double sf = 0.0, ss = 0.0, ef = 0.0, es = 0.0, sg = 0.0, eg = 0.0;
        for (int t = 0; t < L; ++t) {
            const double x = (double)cl[t];
            if (t < ns) { ss += x; if (t >= ns - nf) sf += x; }                               // both EMAs seeded at index ns-1
            if (t < ns - 1) { o0[t] = nan; o1[t] = nan; o2[t] = nan; continue; }
            if (t == ns - 1) { es = ss / ns; ef = sf / nf; }
            else { es = (x - es) * ks + es; ef = (x - ef) * kf + ef; }
            const double macd = ef - es;
            const int q = t - (ns - 1);                                                       // macd sample index
            if (q < ng - 1) { sg += macd; o0[t] = nan; o1[t] = nan; o2[t] = nan; }
            else {
                if (q == ng - 1) { sg += macd; eg = sg / ng; } else eg = (macd - eg) * kg + eg;
                o0[t] = (float)macd; o1[t] = (float)eg; o2[t] = (float)(macd - eg);
            }
        }
    } break;
    case PMRL_IND_OBV: {                                                                      // on-balance volume
        double obv = 0.0;
        for (int t = 0; t < L; ++t) {
            const double v = (double)vo[t];
            if (t == 0) obv = v;
            else if (cl[t] > cl[t - 1]) obv += v;
            else if (cl[t] < cl[t - 1]) obv -= v;
            o0[t] = (float)obv;
        }
    } break;
    case PMRL_IND_ADOSC: {                                                                    // Chaikin A/D oscillator, EMA3 - EMA10 of the A/D line
        const double kf = 2.0 / 4.0, ks = 2.0 / 11.0;
        double ad = 0.0, ef = 0.0, es = 0.0;
        for (int t = 0; t < L; ++t) {
            const double h = hi[t], l = lo[t], c = cl[t], v = vo[t];
            const double rng = h - l;
            if (rng > 0.0) ad += (((c - l) - (h - c)) / rng) * v;
            if (t == 0) { ef = ad; es = ad; }
            else { ef = (ad - ef) * kf + ef; es = (ad - es) * ks + es; }
            o0[t] = t >= 9 ? (float)(ef - es) : nan;
        }
    } break;
    case PMRL_IND_CCI: {                                                                      // (TP - SMA(TP)) / (0.015 * mean deviation)
        for (int t = 0; t < L; ++t) {
            if (t < n - 1) { o0[t] = nan; continue; }
            double sum = 0.0;
            for (int i = t - n + 1; i <= t; ++i) sum += ((double)hi[i] + (double)lo[i] + (double)cl[i]) / 3.0;
            const double m = sum / n;
            double md = 0.0;
            for (int i = t - n + 1; i <= t; ++i) md += fabs(((double)hi[i] + (double)lo[i] + (double)cl[i]) / 3.0 - m);
            md /= n;
            const double tp = ((double)hi[t] + (double)lo[t] + (double)cl[t]) / 3.0;
            o0[t] = (float)(md != 0.0 ? (tp - m) / (0.015 * md) : 0.0);
        }
    } break;
    case PMRL_IND_STOCH: {                                                                    // fastK 5, slowK = SMA3(fastK), slowD = SMA3(slowK)
        double fk[3] = {0, 0, 0}, sk[3] = {0, 0, 0};
        for (int t = 0; t < L; ++t) {
            if (t < 4) { o0[t] = nan; o1[t] = nan; continue; }
            double hh = hi[t], ll = lo[t];
            for (int i = t - 4; i < t; ++i) { hh = fmax(hh, (double)hi[i]); ll = fmin(ll, (double)lo[i]); }
            const double diff = hh - ll;
            fk[t % 3] = diff != 0.0 ? 100.0 * ((double)cl[t] - ll) / diff : 0.0;
            if (t < 6) { o0[t] = nan; o1[t] = nan; continue; }
            const double slowk = (fk[0] + fk[1] + fk[2]) / 3.0;
            sk[t % 3] = slowk;
            if (t < 8) { o0[t] = nan; o1[t] = nan; continue; }
            o0[t] = (float)slowk; o1[t] = (float)((sk[0] + sk[1] + sk[2]) / 3.0);
        }
    } break;
    case PMRL_IND_DX:
    case PMRL_IND_ADX: {                                                                      // Wilder directional movement (ta_DX.c / ta_ADX.c)
        // bars 1..n-1 accumulate +DM, -DM and TR; from bar n on each is Wilder-smoothed (x ← x − x/n + today) and
        // DX = 100·|+DI − −DI| / (+DI + −DI) (first value at index n, a zero denominator repeats the previous DX);
        // ADX = mean of the first n DX values (index 2n−1), then (ADX·(n−1) + DX) / n.
        const bool adx = sp.kind[s] == PMRL_IND_ADX;
        double pdm = 0.0, mdm = 0.0, tr = 0.0, dx = 0.0, sum_dx = 0.0, ax = 0.0;
        if (L > 0) o0[0] = nan;
        for (int t = 1; t < L; ++t) {
            const double h = hi[t], l = lo[t], ph = hi[t - 1], pl = lo[t - 1], pc = cl[t - 1];
            const double dp = h - ph, dm = pl - l;
            const double plus = (dp > 0.0 && dp > dm) ? dp : 0.0, minus = (dm > 0.0 && dp < dm) ? dm : 0.0;
            const double trt = fmax(h - l, fmax(fabs(h - pc), fabs(l - pc)));
            if (t < n) { pdm += plus; mdm += minus; tr += trt; o0[t] = nan; continue; }
            pdm = pdm - pdm / n + plus; mdm = mdm - mdm / n + minus; tr = tr - tr / n + trt;
            bool have = false;
            if (tr != 0.0) {
                const double pdi = 100.0 * (pdm / tr), mdi = 100.0 * (mdm / tr), sm = pdi + mdi;
                if (sm != 0.0) { dx = 100.0 * (fabs(mdi - pdi) / sm); have = true; }
            }
            if (!adx) { o0[t] = (float)dx; continue; }          // DX: an undefined bar repeats the previous value (0 at the start)
            if (t < 2 * n) {                                     // the first n DX values seed the ADX (undefined bars add nothing)
                if (have) sum_dx += dx;
                if (t == 2 * n - 1) { ax = sum_dx / n; o0[t] = (float)ax; } else o0[t] = nan;
            } else {
                if (have) ax = (ax * (n - 1) + dx) / n;
                o0[t] = (float)ax;
            }
        }
    } break;
    default: break;
    }
}

}  // namespace pmrl

using namespace pmrl;

static int spec_outputs(int kind) { return (kind == PMRL_IND_BBANDS || kind == PMRL_IND_MACD) ? 3 : (kind == PMRL_IND_STOCH ? 2 : 1); }
static int spec_lookback(int kind, int n) {
    switch (kind) {
        case PMRL_IND_SMA: case PMRL_IND_EMA: case PMRL_IND_BBANDS: return n - 1;
        case PMRL_IND_RSI: case PMRL_IND_ATR: return n;
        case PMRL_IND_MACD: return 25 + 8;
        case PMRL_IND_OBV: return 0;
        case PMRL_IND_ADOSC: return 9;
        case PMRL_IND_CCI: return n - 1;
        case PMRL_IND_STOCH: return 8;
        case PMRL_IND_DX: return n;
        case PMRL_IND_ADX: return 2 * n - 1;
        default: return -1;
    }
}

extern "C" int pmrl_indicator_layout(const int32_t* specs, int32_t n_specs, int32_t* n_out, int32_t* lookback) {
    if (!specs || n_specs < 0 || n_specs > kMaxSpecs) return pmrl_fail(PMRL_E_ARG, "indicator_layout: need 0 <= n_specs <= 16");
    int no = 0, lb = 0;
    for (int i = 0; i < n_specs; ++i) {
        const int kind = specs[2 * i], n = specs[2 * i + 1];
        const int l = spec_lookback(kind, n);
        const bool fixed = kind == PMRL_IND_MACD || kind == PMRL_IND_OBV || kind == PMRL_IND_ADOSC || kind == PMRL_IND_STOCH;
        if (l < 0 || (!fixed && n < 2)) return pmrl_fail(PMRL_E_ARG, "indicator_layout: unknown indicator or period < 2");
        no += spec_outputs(kind);
        if (l > lb) lb = l;
    }
    if (n_out) *n_out = no;
    if (lookback) *lookback = lb;
    return 0;
}

extern "C" int pmrl_indicators(const float* series, int32_t A, int32_t C, int32_t L, const int32_t* specs, int32_t n_specs,
                               float* out, void* stream) {
    if (!series || !out || !specs) return pmrl_fail(PMRL_E_ARG, "indicators: NULL pointer");
    if (A < 1 || C < 4 || L < 2) return pmrl_fail(PMRL_E_SHAPE, "indicators: need A >= 1, C >= 4 (o,h,l,c[,v]) and L >= 2");
    int n_out = 0, lb = 0;
    if (int rc = pmrl_indicator_layout(specs, n_specs, &n_out, &lb)) return rc;
    if (n_specs == 0) return 0;
    for (int i = 0; i < n_specs; ++i)
        if ((specs[2 * i] == PMRL_IND_OBV || specs[2 * i] == PMRL_IND_ADOSC) && C < 5) return pmrl_fail(PMRL_E_SHAPE, "indicators: OBV / ADOSC need the volume channel (C >= 5)");
    IndSpecs sp;
    sp.n = n_specs;
    int o = 0;
    for (int i = 0; i < n_specs; ++i) { sp.kind[i] = specs[2 * i]; sp.period[i] = specs[2 * i + 1]; sp.out0[i] = o; o += spec_outputs(sp.kind[i]); }
    dim3 grid((A + 63) / 64, n_specs);
    k_indicators<<<grid, 64, 0, (cudaStream_t)stream>>>(series, A, C, L, sp, n_out, out);
    return pmrl_check_launch("k_indicators");
}
