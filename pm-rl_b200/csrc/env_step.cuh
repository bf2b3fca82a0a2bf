// env_step.cuh — one warp advances one environment by one step.
//
// Restates, batched and in registers, the reference transition
//   TradingEnv.step            env/sim/trading_env.py:54-100
//   ActionBuffer.update/get_last  env/sim/weight_buffer.py:13-30
//   Reward.*                   env/reward.py:15-31
//   y_t = close_t / close_{t-1}   data/instrument.py:79  (precomputed once per table: pmrl_price_relatives)
// keeping the reference association of every fp32 operation (no FMA contraction, IEEE div,
// expf/logf without fast-math) so values stay within 1e-5 relative over 1,000 compounding steps.
//
// The transition is split into three stages so that the state-only kernel can software-pipeline them
// across consecutive envs of a warp (scalars of env n+2, vectors of env n+1 and the arithmetic of env n
// are in flight together):
//   env_load_scalars   V, ring pointer, is_full, local step, episode offset
//   env_load_vectors   raw action, the price-relative row (or external y), previous weights (commission only)
//   env_compute_store  everything else
#pragma once
#include "pmrl_device.cuh"

namespace pmrl {

// Per-warp partial of the PMRL_STAT_* vector lives in shared memory (10 doubles per warp, touched only by
// lane 0) so that it costs no registers across the long-lived persistent loops.
__device__ __forceinline__ void stats_init_block(double* sm /* [warps*10] */, int nwarps) {
    for (int q = threadIdx.x; q < nwarps * PMRL_STATS_LEN; q += blockDim.x)
        sm[q] = (q % PMRL_STATS_LEN >= PMRL_STAT_MAX_V) ? -INFINITY : 0.0;
    __syncthreads();
}

struct StepOut {       // warp-uniform result of one env transition
    int idx_new;       // ring pointer after the step
    int slot_written;  // ring row that received w' (or -1 on auto-reset)
    int is_full;
    int k;             // local step after the call
    int did_reset;
    float V, reward;
    int done;
};

struct EnvScalars { float V; int i, full, k, t0e; };

template <int NPL, bool HASC>
struct EnvVectors {
    float a[NPL];                   // raw action → weights → holdings → w'  (updated in place)
    float y[NPL];                   // price relative close_t / close_{t-1} (table row or external)
    float wl[HASC ? NPL : 1];       // previous post-drift weights (only read when commission > 0)
};

// Is asset slot j of this lane a real asset?  With TAIL the caller guarantees A > 32·(NPL−1) — every slot row but the
// last is full — so the test folds to `true` at compile time for j < NPL−1 and only the last row is guarded
// (all BASELINE shapes: A = 11, 50, 100, 500 ↔ NPL = 1, 2, 4, 16).
template <int NPL, bool TAIL>
__device__ __forceinline__ bool slot_ok(int j, int lane, int A) {
    if (TAIL && j < NPL - 1) return true;
    return lane + 32 * j < A;
}

// Zero the ring of env e and set the all-cash row (weight_buffer.py:46-50).
__device__ __forceinline__ void ring_reset_warp(float* __restrict__ hist_e, int W, int A, int lane) {
    const int n = W * A;
    for (int i = lane; i < n; i += 32) hist_e[i] = (i == 0) ? 1.0f : 0.0f;
}

__device__ __forceinline__ void env_load_scalars(const StepParams& p, int e, EnvScalars& s) {
    s.V = p.value[e];
    s.i = p.idx[e];
    s.full = p.is_full[e];
    s.k = p.t[e];
    s.t0e = p.t0 ? p.t0[e] : 0;
}

__device__ __forceinline__ bool env_needs_reset(const StepParams& p, const EnvScalars& s) {
    return p.episode_len > 0 && s.k >= p.episode_len;           // train/on_policy.py:60-61
}

template <int NPL, bool HASC, bool TAIL = false>
__device__ __forceinline__ void env_load_vectors(const StepParams& p, int e, int lane, const EnvScalars& s,
                                                 EnvVectors<NPL, HASC>& v) {
    if (env_needs_reset(p, s)) return;                           // nothing is read on the auto-reset call
    const int A = p.A, W = p.W;
    const size_t eA = (size_t)e * A;
    const uint64_t pol_once = l2_policy_evict_first();
    // one validity test per slot shared by the three loads (slots beyond A read as 0 everywhere)
    const float* __restrict__ act = p.actions + eA;
    const float* __restrict__ wrow = p.hist;
    if (HASC) wrow += ((size_t)e * W + (s.i - 1 + W) % W) * A;          // weight_buffer.py:30
    if (p.y_ext) {
        const float* __restrict__ yrow = p.y_ext + eA;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int a = lane + 32 * j;
            const bool ok = slot_ok<NPL, TAIL>(j, lane, A);
            v.a[j] = ok ? ld_once(act + a, pol_once) : 0.0f;
            v.y[j] = ok ? ld_once(yrow + a, pol_once) : 0.0f;
            if (HASC) v.wl[j] = ok ? wrow[a] : 0.0f;
        }
    } else {                                                     // row t0 + k_new + W - 1: last row of the new window
        const uint64_t pol_keep = l2_policy_evict_last();
        const float* __restrict__ yrow = p.y_tm + (size_t)(s.t0e + s.k + W) * A;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int a = lane + 32 * j;
            const bool ok = slot_ok<NPL, TAIL>(j, lane, A);
            v.a[j] = ok ? ld_once(act + a, pol_once) : 0.0f;
            v.y[j] = ok ? ld_keep(yrow + a, pol_keep) : 0.0f;
            if (HASC) v.wl[j] = ok ? wrow[a] : 0.0f;
        }
    }
}

// Pull the DRAM-resident rows env e will read (its raw action, and its previous weights when commission > 0) into L2
// one env ahead of the warp, without holding registers for them: one 128-byte line per lane.  The price-relative row
// lives in the L2-resident table already.
template <bool HASC>
__device__ __forceinline__ void env_prefetch_vectors(const StepParams& p, int e, int lane, const EnvScalars& s) {
    if (env_needs_reset(p, s)) return;
    const int A = p.A;
    for (int a = lane * 32; a < A; a += 32 * 32) prefetch_l2(p.actions + (size_t)e * A + a);
    if (HASC) {
        const int last_slot = (s.i - 1 + p.W) % p.W;
        const float* __restrict__ wrow = p.hist + ((size_t)e * p.W + last_slot) * A;
        for (int a = lane * 32; a < A; a += 32 * 32) prefetch_l2(wrow + a);
    }
}

// On return v.a[] holds w' (the post-drift weights, lane-strided: asset a = lane + 32*j).
template <int NPL, bool HASC, bool TAIL = false>
__device__ __forceinline__ void env_compute_store(const StepParams& p, int e, int lane, const EnvScalars& sc,
                                                  EnvVectors<NPL, HASC>& v, StepOut& out,
                                                  double* __restrict__ acc /* smem [10] of this warp */) {
    const int A = p.A, W = p.W;
    float* __restrict__ hist_e = p.hist + (size_t)e * W * A;

    // ---- auto-reset instead of a step (train/on_policy.py:60-61) ----
    if (env_needs_reset(p, sc)) {
        ring_reset_warp(hist_e, W, A, lane);
        if (lane == 0) {
            p.value[e] = p.initial_cash;
            p.idx[e] = 1; p.is_full[e] = 0; p.t[e] = 0;
            p.reward[e] = 0.0f; p.done[e] = 0;
            if (p.sharpe) { p.sharpe[3 * (size_t)e] = 0.0; p.sharpe[3 * (size_t)e + 1] = 0.0; p.sharpe[3 * (size_t)e + 2] = 0.0; }
            if (p.ep_return) p.ep_return[e] = 0.0f;
        }
#pragma unroll
        for (int j = 0; j < NPL; ++j) v.a[j] = (lane + 32 * j == 0) ? 1.0f : 0.0f;
        out.idx_new = 1; out.slot_written = -1; out.is_full = 0; out.k = 0; out.did_reset = 1;
        out.V = p.initial_cash; out.reward = 0.0f; out.done = 0;
        return;
    }

    float V = sc.V;
    const int i = sc.i;
    int full = sc.full;
    const int k_new = sc.k + 1;

    // ---- normalise (trading_env.py:58-60; quirks Q1-Q3) ----
    // slots beyond A hold 0: neutral for the sum, and for the minimum too, which is only ever compared with 0 (has_neg)
    // and with the −41 bound of the shared-reciprocal softmax below (a smaller minimum only makes that test stricter)
    float s = 0.0f, mn = INFINITY;
#pragma unroll
    for (int j = 0; j < NPL; ++j) { s = __fadd_rn(s, v.a[j]); mn = nanmin(mn, v.a[j]); }
    s = warp_sum(s);
    mn = warp_min_nan(mn);
    const bool strict = (p.flags & PMRL_FLAG_STRICT_REFERENCE) != 0;
    const bool not_close = !isclose_one(s);
    const bool has_neg = mn < 0.0f;
    const bool normalise = strict ? (not_close && has_neg) : (not_close || has_neg);
    if (normalise) {
        float mx = 0.0f;
        if (!strict) {                                           // stabilised softmax (agent/pg/pg.py:53)
            mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NPL; ++j) if (slot_ok<NPL, TAIL>(j, lane, A)) mx = fmaxf(mx, v.a[j]);
            mx = warp_max(mx);
        }
        float se = 0.0f;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            v.a[j] = slot_ok<NPL, TAIL>(j, lane, A) ? expf(__fsub_rn(v.a[j], mx)) : 0.0f;
            se = __fadd_rn(se, v.a[j]);
        }
        se = warp_sum(se);
        // e_j / Σe over one divisor.  min_j e_j = exp(min_j a_j − mx) and max_j e_j ≤ Σe, so the shared-reciprocal form is
        // exact when the smallest raw score is above −41 (e^-41 > 2^-60) and Σe is in range; NaNs fail both tests.
        if (__fsub_rn(mn, mx) >= -41.0f && se >= kUniDivLo && se <= kUniDivHi) {
            const UniDiv d = unidiv_make(se);
#pragma unroll
            for (int j = 0; j < NPL; ++j) v.a[j] = unidiv(v.a[j], d);
        } else {
#pragma unroll
            for (int j = 0; j < NPL; ++j) v.a[j] = __fdiv_rn(v.a[j], se);
        }
    }

    // ---- transaction remainder factor mu (trading_env.py:67-75; upstream PGPortfolio relu form) ----
    const float V_prev = V;
    if (HASC) {
        const float c = p.commission;
        const float w0 = __shfl_sync(PMRL_FULL_MASK, v.a[0], 0);
        const float wl0 = __shfl_sync(PMRL_FULL_MASK, v.wl[0], 0);
        const float denom = __fsub_rn(1.0f, __fmul_rn(c, w0));
        const float cw = __fmul_rn(c, wl0);
        float mu_last = 1.0f, mu = p.mu0;
        int it = 0;
        // the sum runs over assets i >= 1: asset 0 (cash) is taken out by a −inf previous weight, the slots beyond A hold
        // wl = w = 0 and contribute relu(0) = 0 — no per-slot guard inside the iteration
        if (lane == 0) v.wl[0] = -INFINITY;
        while (fabsf(__fsub_rn(mu, mu_last)) > 1e-10f && it < p.mu_max_iter) {
            mu_last = mu;
            float part = 0.0f;
#pragma unroll
            for (int j = 0; j < NPL; ++j)
                part = __fadd_rn(part, fmaxf(__fsub_rn(v.wl[j], __fmul_rn(mu, v.a[j])), 0.0f));
            part = warp_sum(part);
            const float numer = __fsub_rn(__fsub_rn(1.0f, cw), __fmul_rn(p.c2, part));
            mu = __fdiv_rn(numer, denom);
            ++it;
        }
        V = __fmul_rn(mu, V);                                    // trading_env.py:75
    }

    // ---- value, drift, return (trading_env.py:78-90) ----
    float part = 0.0f, lo = INFINITY, hi = 0.0f;           // lo/hi: range of |port_j| for the shared-reciprocal division
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        v.a[j] = slot_ok<NPL, TAIL>(j, lane, A) ? __fmul_rn(V, __fmul_rn(v.a[j], v.y[j])) : 0.0f;
        part = __fadd_rn(part, v.a[j]);
    }
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        const float ap = slot_ok<NPL, TAIL>(j, lane, A) ? fabsf(v.a[j]) : 1.0f;
        lo = fminf(lo, ap);
        hi = fmaxf(hi, ap);
    }
    const float Vn = warp_sum(part);
    // w' = port / V' (trading_env.py:83): one divisor for the whole env.  A zero, denormal, huge or NaN holding anywhere in
    // the env (fminf/fmaxf drop NaNs, so those are caught through V') sends the warp down the plain IEEE division.
    const bool uni = __all_sync(PMRL_FULL_MASK, lo >= kUniDivLo && hi <= kUniDivHi) && unidiv_in_range(Vn);
    if (uni) {
        const UniDiv d = unidiv_make(Vn);
#pragma unroll
        for (int j = 0; j < NPL; ++j) v.a[j] = unidiv(v.a[j], d);
    } else {
#pragma unroll
        for (int j = 0; j < NPL; ++j) v.a[j] = __fdiv_rn(v.a[j], Vn);
    }
    const float ret = __fdiv_rn(Vn, V);

    // ---- ring write (weight_buffer.py:21-26) ----
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        const int a = lane + 32 * j;
        if (slot_ok<NPL, TAIL>(j, lane, A)) hist_e[(size_t)i * A + a] = v.a[j];
    }
    const int i_new = (i + 1 == W) ? 0 : i + 1;
    if (i_new == 0) full = 1;

    // ---- reward (trading_env.py:99; env/reward.py:15-31) ----
    float r;
    if (p.reward_mode == PMRL_REWARD_STEP_LOG) {
        r = __fmul_rn(logf(ret), p.reward_scale);
    } else if (p.reward_mode == PMRL_REWARD_RETURNS) {
        r = __fmul_rn(__fdiv_rn(Vn, V_prev), p.reward_scale);
    } else if (p.reward_mode == PMRL_REWARD_LOG_RETURNS) {
        r = __fmul_rn(logf(__fdiv_rn(Vn, V_prev)), p.reward_scale);
    } else {  // running Sharpe over the episode's value history, fp64 like numpy (reward.py:26-31)
        r = 0.0f;
        if (lane == 0) {
            double* sh = p.sharpe + 3 * (size_t)e;
            double n = sh[0], mean = sh[1], m2 = sh[2];
            const double g = (double)Vn / (double)V_prev;
            n += 1.0;
            const double d1 = g - mean;
            mean += d1 / n;
            m2 += d1 * (g - mean);
            sh[0] = n; sh[1] = mean; sh[2] = m2;
            const double sd = sqrt(m2 / (n - 1.0));              // ddof=1 → NaN at n == 1 (quirk Q11)
            r = (float)(((mean - (double)p.risk_free) / sd) * (double)p.reward_scale);
        }
        r = __shfl_sync(PMRL_FULL_MASK, r, 0);
    }

    const int dn = (p.episode_len > 0 && k_new == p.episode_len) ? 1 : 0;
    if (lane == 0) {
        p.value[e] = Vn;
        p.idx[e] = i_new; p.is_full[e] = (uint8_t)full; p.t[e] = k_new;
        p.reward[e] = r; p.done[e] = (uint8_t)dn;
        float epr = 0.0f;
        if (p.ep_return) { epr = __fadd_rn(p.ep_return[e], r); p.ep_return[e] = epr; }
        if (p.stats) {
            acc[PMRL_STAT_N_ENVS] += 1.0; acc[PMRL_STAT_SUM_R] += r; acc[PMRL_STAT_SUM_R2] += (double)r * r;
            acc[PMRL_STAT_SUM_V] += Vn; acc[PMRL_STAT_SUM_LNV] += (double)logf(Vn);   // fp32 log: a double-precision log costs ~100 instructions per env
            if (dn) { acc[PMRL_STAT_N_DONE] += 1.0; acc[PMRL_STAT_SUM_EPRET] += epr; acc[PMRL_STAT_SUM_EPLEN] += k_new; }
            acc[PMRL_STAT_MAX_V] = fmax(acc[PMRL_STAT_MAX_V], (double)Vn);
            acc[PMRL_STAT_MAX_NEGV] = fmax(acc[PMRL_STAT_MAX_NEGV], -(double)Vn);
        }
    }
    out.idx_new = i_new; out.slot_written = i; out.is_full = full; out.k = k_new; out.did_reset = 0;
    out.V = Vn; out.reward = r; out.done = dn;
}

// The three stages back to back (fused kernel phase 1).
template <int NPL, bool HASC>
__device__ __forceinline__ void env_step_warp(const StepParams& p, int e, int lane, EnvVectors<NPL, HASC>& v,
                                              StepOut& out, double* __restrict__ acc) {
    EnvScalars sc;
    env_load_scalars(p, e, sc);
    env_load_vectors<NPL, HASC>(p, e, lane, sc, v);
    env_compute_store<NPL, HASC>(p, e, lane, sc, v, out, acc);
}

// Block-level flush of the per-warp partials: thread q < 10 folds column q over the warps → 10 atomics per CTA.
__device__ __forceinline__ void stats_flush_block(double* __restrict__ stats, const double* sm /* [warps*10] */, int nwarps) {
    __syncthreads();
    if (threadIdx.x < PMRL_STATS_LEN) {
        const int q = threadIdx.x;
        double v = sm[q];
        for (int wi = 1; wi < nwarps; ++wi) {
            const double o = sm[wi * PMRL_STATS_LEN + q];
            v = (q >= PMRL_STAT_MAX_V) ? fmax(v, o) : v + o;
        }
        if (q >= PMRL_STAT_MAX_V) { if (v > -INFINITY) atomic_max_double(stats + q, v); }
        else if (v != 0.0) atomicAdd(stats + q, v);
    }
}

}  // namespace pmrl
