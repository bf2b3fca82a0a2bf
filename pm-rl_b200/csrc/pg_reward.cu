// pg_reward.cu — the differentiable batched restatement of the transition that the PG agent uses as its
// loss (agent/pg/pg.py:40-82), forward + analytic gradient w.r.t. the raw action in one kernel.
//
// One warp per batch row b:
//   a' = softmax(a_b) if `normalise` (torch.softmax(a, dim=1), pg.py:53) else a_b
//   mu : commission fixed point (pg.py:57-65, relu form), pv' = mu * pv
//   v  = Σ_i pv'·(a'_i·p_i);  ret = v / pv'                                   (pg.py:68-72)
//   rew_b = ret·scale | ln(ret)·scale                                         (pg.py:75-78)
// Loss side: R = mean_b rew_b.  ret does not depend on mu mathematically (it cancels in v/pv'), so
//   dR/da'_i = g·p_i with g = scale/B (returns) or scale/(B·Σ_i a'_i p_i) (log-returns) and, through the softmax,
//   dR/da_j = a'_j (g p_j − Σ_i a'_i g p_i).   grad_a = gscale · dR/da  (gscale = −1 for loss = −R, pg.py:101).
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "pmrl_b200.h"
#include "pmrl_device.cuh"
#include "host_util.h"

namespace pmrl {

constexpr int kPgThreads = 256;

template <int NPL>
__global__ void __launch_bounds__(kPgThreads) k_pg_reward(int B, int A, int mode, int normalise, float c, float mu0,
                                                          float c2, float scale, int mu_max_iter,
                                                          const float* __restrict__ a, const float* __restrict__ pv,
                                                          const float* __restrict__ pa, const float* __restrict__ p,
                                                          float* __restrict__ rew, float* __restrict__ grad_a, float gscale) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * kPgThreads + threadIdx.x) >> 5;
    const int nw = (gridDim.x * kPgThreads) >> 5;
    for (int b = gw; b < B; b += nw) {
        const size_t o = (size_t)b * A;
        float w[NPL], y[NPL], wl[NPL];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int i = lane + 32 * j;
            w[j] = (i < A) ? a[o + i] : 0.0f;
            y[j] = (i < A) ? p[o + i] : 0.0f;
            wl[j] = (c > 0.0f && i < A) ? pa[o + i] : 0.0f;
            if (i < A) mx = fmaxf(mx, w[j]);
        }
        if (normalise) {
            mx = warp_max(mx);
            float se = 0.0f;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                w[j] = (lane + 32 * j < A) ? expf(__fsub_rn(w[j], mx)) : 0.0f;
                se = __fadd_rn(se, w[j]);
            }
            se = warp_sum(se);
#pragma unroll
            for (int j = 0; j < NPL; ++j) w[j] = __fdiv_rn(w[j], se);
        }
        float V = pv[b];
        if (c > 0.0f) {
            const float w0 = __shfl_sync(PMRL_FULL_MASK, w[0], 0);
            const float wl0 = __shfl_sync(PMRL_FULL_MASK, wl[0], 0);
            const float denom = __fsub_rn(1.0f, __fmul_rn(c, w0));
            const float cw = __fmul_rn(c, wl0);
            float mu_last = 1.0f, mu = mu0;
            int it = 0;
            while (fabsf(__fsub_rn(mu, mu_last)) > 1e-10f && it < mu_max_iter) {
                mu_last = mu;
                float part = 0.0f;
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    const int i = lane + 32 * j;
                    if (i >= 1 && i < A) part = __fadd_rn(part, fmaxf(__fsub_rn(wl[j], __fmul_rn(mu, w[j])), 0.0f));
                }
                part = warp_sum(part);
                mu = __fdiv_rn(__fsub_rn(__fsub_rn(1.0f, cw), __fmul_rn(c2, part)), denom);
                ++it;
            }
            V = __fmul_rn(mu, V);
        }
        float part = 0.0f, dot = 0.0f;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const float wy = __fmul_rn(w[j], y[j]);
            part = __fadd_rn(part, __fmul_rn(V, wy));
            dot = __fadd_rn(dot, wy);
        }
        const float v = warp_sum(part);
        dot = warp_sum(dot);                                  // Σ a'_i p_i (the exact-arithmetic ret)
        const float ret = __fdiv_rn(v, V);
        const float r = (mode == PMRL_REWARD_RETURNS) ? __fmul_rn(ret, scale) : __fmul_rn(logf(ret), scale);
        if (lane == 0) rew[b] = r;
        if (grad_a) {
            const float g = (mode == PMRL_REWARD_RETURNS) ? scale / (float)B : scale / ((float)B * dot);
            const float gd = g * dot;                         // Σ_i a'_i g p_i
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const int i = lane + 32 * j;
                if (i < A) {
                    const float gi = g * y[j];
                    grad_a[o + i] = gscale * (normalise ? w[j] * (gi - gd) : gi);
                }
            }
        }
    }
}

}  // namespace pmrl

using namespace pmrl;

template <int NPL>
static int launch_pg(int B, int A, int mode, int normalise, float c, float scale, int it,
                     const float* a, const float* pv, const float* pa, const float* p, float* rew, float* grad_a,
                     float gscale, cudaStream_t s) {
    const double cd = (double)c;
    const int warps = kPgThreads / 32;
    int blocks = (B + warps - 1) / warps;
    const int cap = pmrl_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_pg_reward<NPL><<<blocks, kPgThreads, 0, s>>>(B, A, mode, normalise, c, (float)(1.0 - 2.0 * cd + cd * cd),
                                                  (float)(2.0 * cd - cd * cd), scale, it, a, pv, pa, p, rew, grad_a, gscale);
    return pmrl_check_launch("k_pg_reward");
}

extern "C" int pmrl_pg_reward_fwd_bwd(int32_t B, int32_t A, int32_t mode, int32_t normalise, float commission, float scale,
                                      int32_t mu_max_iter,
                                      const float* a, const float* pv, const float* pa, const float* p,
                                      float* rew, float* grad_a, float gscale, void* stream) {
    if (B < 0 || A < 1) return pmrl_fail(PMRL_E_SHAPE, "pg_reward: bad sizes");
    if (A > 1024) return pmrl_fail(PMRL_E_SHAPE, "pg_reward: A > 1024 is not supported");
    if (mode != PMRL_REWARD_RETURNS && mode != PMRL_REWARD_LOG_RETURNS) return pmrl_fail(PMRL_E_ARG, "pg_reward: mode must be RETURNS or LOG_RETURNS");
    if (!a || !pv || !p || !rew || (commission > 0.0f && !pa)) return pmrl_fail(PMRL_E_ARG, "pg_reward: NULL pointer");
    if (B == 0) return 0;
    const int it = mu_max_iter > 0 ? mu_max_iter : 16;
    cudaStream_t s = (cudaStream_t)stream;
#define PG_CASE(N) return launch_pg<N>(B, A, mode, normalise, commission, scale, it, a, pv, pa, p, rew, grad_a, gscale, s)
    if (A <= 32) PG_CASE(1);
    if (A <= 64) PG_CASE(2);
    if (A <= 128) PG_CASE(4);
    if (A <= 256) PG_CASE(8);
    if (A <= 512) PG_CASE(16);
    PG_CASE(32);
#undef PG_CASE
}
