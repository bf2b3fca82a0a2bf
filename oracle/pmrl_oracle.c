/* pmrl_oracle.c — plain-C restatement of the reference transition — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * One call advances E independent envs by one step exactly like TradingEnv.step (env/sim/trading_env.py:54-100) with
 * ActionBuffer.update/get_last (env/sim/weight_buffer.py:13-30): fp32 arithmetic in the reference's association,
 * un-stabilised softmax under the AND-condition (quirks Q1-Q3), commission fixed point with the upstream relu
 * (trading_env.py:67-75), log-return reward (:99), auto-reset on the call after `done` (train/on_policy.py:60-61).
 * Parity: pinned for c == 0 by tests/test_c_oracle.py (golden fixtures of the live reference + the numpy oracle);
 * c > 0 unpinned like the numpy oracle (the reference raises TypeError, trading_env.py:72).
 * Envs are independent → `#pragma omp parallel for` over envs gives the multi-core CPU baseline of bench.py.
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC → oracle/_build/liboracle.so)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* Sums over the asset axis use 8 interleaved partial sums combined by a tree — the shape of the vectorised reductions
 * torch / numpy run on the CPU; a single sequential accumulator drifts ~1e-5 relative in V over 1,000 compounding steps. */
static float sum8(const float* x, int n) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    for (; i + 8 <= n; i += 8)
        for (int k = 0; k < 8; ++k) acc[k] += x[i + k];
    for (int k = 0; i < n; ++i, ++k) acc[k] += x[i];
    return ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
}

static int isclose_one(float s) {
    const float allowed = 1e-6f + fabsf(1e-5f * 1.0f);
    const float err = fabsf(s - 1.0f);
    return (s == 1.0f) || (isfinite(err) && err <= allowed);
}

void pmrl_oracle_step(int E, int A, int W, const float* actions /*[E,A]*/, const float* y /*[E,A]*/,
                      float* value /*[E]*/, float* hist /*[E,W,A]*/, int32_t* idx, uint8_t* is_full, int32_t* t,
                      int episode_len, float initial_cash, float commission, float reward_scale, int strict,
                      int mu_max_iter, float* reward /*[E]*/, uint8_t* done /*[E]*/, float* scratch /*[threads? no: E*A]*/) {
    const float c = commission;
    const float mu0 = (float)(1.0 - 2.0 * (double)c + (double)c * (double)c);
    const float c2 = (float)(2.0 * (double)c - (double)c * (double)c);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < E; ++e) {
        float* h = hist + (size_t)e * W * A;
        if (episode_len > 0 && t[e] >= episode_len) {                     /* auto-reset instead of a step */
            memset(h, 0, sizeof(float) * (size_t)W * A);
            h[0] = 1.0f;
            value[e] = initial_cash; idx[e] = 1; is_full[e] = 0; t[e] = 0; reward[e] = 0.0f; done[e] = 0;
            continue;
        }
        const float* a = actions + (size_t)e * A;
        const float* ye = y + (size_t)e * A;
        float* w = scratch + (size_t)e * A;
        float mn = a[0];
        int has_nan = 0;
        for (int i = 0; i < A; ++i) { if (a[i] != a[i]) has_nan = 1; if (a[i] < mn) mn = a[i]; }
        const float s = sum8(a, A);
        if (has_nan) mn = NAN;                                            /* torch.min propagates NaN */
        const int not_close = !isclose_one(s), has_neg = mn < 0.0f;
        const int normalise = strict ? (not_close && has_neg) : (not_close || has_neg);
        if (normalise) {                                                  /* trading_env.py:59-60 */
            float mx = 0.0f;
            if (!strict) { mx = a[0]; for (int i = 1; i < A; ++i) if (a[i] > mx) mx = a[i]; }
            for (int i = 0; i < A; ++i) w[i] = expf(a[i] - mx);
            const float se = sum8(w, A);
            for (int i = 0; i < A; ++i) w[i] = w[i] / se;
        } else {
            for (int i = 0; i < A; ++i) w[i] = a[i];
        }
        const int i0 = idx[e];
        const float* wl = h + (size_t)((i0 - 1 + W) % W) * A;             /* weight_buffer.py:30 */
        float V = value[e];
        if (c > 0.0f) {                                                   /* trading_env.py:67-75 */
            float mu_last = 1.0f, mu = mu0;
            const float denom = 1.0f - c * w[0], cw = c * wl[0];
            int it = 0;
            while (fabsf(mu - mu_last) > 1e-10f && it < mu_max_iter) {
                mu_last = mu;
                float part = 0.0f;                                        /* (second scratch row not needed: A-1 terms, blocked) */
                float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int i = 1; i < A; ++i) { const float d = wl[i] - mu * w[i]; acc[(i - 1) & 7] += d > 0.0f ? d : 0.0f; }
                part = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
                mu = ((1.0f - cw) - c2 * part) / denom;
                ++it;
            }
            V = mu * V;
        }
        for (int i = 0; i < A; ++i) w[i] = V * (w[i] * ye[i]);           /* :78 */
        const float Vn = sum8(w, A);                                      /* :79 */
        float* row = h + (size_t)i0 * A;
        for (int i = 0; i < A; ++i) row[i] = w[i] / Vn;                   /* :83, weight_buffer.py:21 */
        const float ret = Vn / V;                                         /* :88 */
        const int i1 = (i0 + 1) % W;
        idx[e] = i1;
        if (i1 == 0) is_full[e] = 1;
        value[e] = Vn;
        t[e] += 1;
        reward[e] = logf(ret) * reward_scale;                             /* :99 */
        done[e] = (episode_len > 0 && t[e] == episode_len) ? 1 : 0;
    }
}
