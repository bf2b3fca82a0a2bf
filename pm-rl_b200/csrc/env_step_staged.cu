// env_step_staged.cu — state-only step ("Mode S") for WIDE envs (129..1024 assets, A % 4 == 0): the three rows an env reads —
// raw action, previous weights (commission only), price relatives — enter the SM as TMA bulk copies (cp.async.bulk
// global→shared on an mbarrier, SASS UBLKCP) one env AHEAD of the warp that consumes them, double-buffered per warp.
//
// Why: a 500-asset env is 16 slots per lane; with action + previous weights + price relatives in registers there is no room
// for a second env's rows, so k_env_step (env_kernels.cu) could only pull the next env's rows into L2 and then paid an L2
// round trip per row at the head of every env (ncu, config 5: 3.0 warps per scheduler stalled on the long scoreboard, issue
// slots 59 % busy, 8 spilled registers at the 80-register / 3-CTA build: 0.390 ms on config 5).  Here the DRAM latency is
// taken by the TMA unit: a warp finds its rows in shared memory (ld.shared.v4, conflict-free: lane-consecutive 16-byte groups), the price
// relatives are read after the mu iteration straight from the staged row (no registers held for them), and 107 registers
// at two CTAs per SM leave nothing spilled: 0.335 ms (0.72 of the HBM roofline, instruction-issue-bound).  The arithmetic is env_compute_rows of env_step.cuh — the same operations in
// the same order as every other step kernel (cross-kernel bit-identity is asserted in tests/test_env_gpu.py).
//
// Shared memory per warp: 2 stages x R rows x A floats (R = 3 with commission, else 2): 12,000 B at A = 500 → 96 KB per
// 8-warp CTA, two CTAs per SM.  One mbarrier per (warp, stage); lane 0 issues the copies of env n+1 before the warp
// starts on env n (the stage it overwrites was read one env earlier; every lane has passed the warp-wide reductions of
// that env since, and lane 0 orders the generic reads before the async-proxy writes with fence.proxy.async).
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "pmrl_b200.h"
#include "pmrl_device.cuh"
#include "env_step.cuh"
#include "env_launch.h"
#include "host_util.h"

namespace pmrl {

template <int NPL, bool HASC, bool TAIL, int kStgWarps, int MINB, bool NOSINKS>
__global__ void __launch_bounds__(kStgWarps * 32, MINB) k_env_step_staged(const StepParams p) {
    constexpr int VEC = 4;
    constexpr int R = HASC ? 3 : 2;                                  // staged rows per env: action, price relatives[, previous weights]
    extern __shared__ __align__(128) float s_rows[];                 // [warps][2 stages][R][A]
    __shared__ double s_stats[kStgWarps * PMRL_STATS_LEN];
    __shared__ __align__(8) uint64_t s_bar[kStgWarps][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kStgWarps + warp;
    const int nw = gridDim.x * kStgWarps;
    const int A = p.A, W = p.W;
    const uint32_t row_bytes = (uint32_t)A * 4u;
    float* const my_rows = s_rows + (size_t)warp * 2 * R * A;
    const uint32_t my_rows_s = smem_u32(my_rows);
    if (lane == 0) { mbar_init(&s_bar[warp][0], 1); mbar_init(&s_bar[warp][1], 1); mbar_fence_init(); }
    __syncwarp();
    WarpStats ws;
    wstats_init(ws);
    StepOut so;

    // lane 0: request the rows of env e (scalars s) into stage st
    auto issue = [&](int e, const EnvScalars& s, int st) {
        if (env_needs_reset(p, s)) return;                           // the auto-reset call reads nothing
        if (lane == 0) {
            if (!NOSINKS && p.act_ready) { env_wait_actions_lane(p, e); asm volatile("fence.proxy.async;" ::: "memory"); }
            fence_proxy_async_smem();
            float* dst = my_rows + (size_t)st * R * A;
            uint64_t* bar = &s_bar[warp][st];
            mbar_arrive_expect_tx(bar, R * row_bytes);
            bulk_load_g2s(dst, p.actions + (size_t)e * A, row_bytes, bar, kPolicyEvictFirst);
            if (p.y_ext) bulk_load_g2s(dst + A, env_y_row(p, e, s), row_bytes, bar, kPolicyEvictFirst);
            else bulk_load_g2s(dst + A, env_y_row(p, e, s), row_bytes, bar, kPolicyEvictLast);
            if (HASC) bulk_load_g2s(dst + 2 * A, p.hist + ((size_t)e * W + (s.i == 0 ? W : s.i) - 1) * A, row_bytes, bar, kPolicyEvictFirst);
        }
    };

    EnvScalars s0, s1, s2;
    int e = gw;
    if (e < p.E) { env_load_scalars(p, e, s0); issue(e, s0, 0); }
    if (e + nw < p.E) env_load_scalars(p, e + nw, s1);
    uint32_t uses0 = 0, uses1 = 0;                                   // completed waits per stage → mbarrier phase parity
    for (int n = 0; e < p.E; e += nw, ++n) {
        const int st = n & 1;
        if (e + 2 * nw < p.E) env_load_scalars(p, e + 2 * nw, s2);
        __syncwarp();                                                // every lane is done with the stage about to be refilled
        if (e + nw < p.E) issue(e + nw, s1, st ^ 1);
        float va[NPL], vy[NPL], vwl[HASC ? NPL : 1];
        const uint32_t stage_s = my_rows_s + (uint32_t)st * R * row_bytes;
        if (!env_needs_reset(p, s0)) {
            const uint32_t par = (st ? uses1 : uses0) & 1u;
            mbar_wait(&s_bar[warp][st], par);
            if (st) ++uses1; else ++uses0;
#pragma unroll
            for (int g = 0; g < NPL / VEC; ++g) {
                const uint32_t off = 4u * (uint32_t)((g * 32 + lane) * VEC);
                if (slot_ok<NPL, VEC, TAIL>(g * VEC, lane, A)) {
                    lds_v<VEC>(stage_s + off, &va[g * VEC]);
                    if (HASC) lds_v<VEC>(stage_s + 2 * row_bytes + off, &vwl[g * VEC]);
                } else {
#pragma unroll
                    for (int c = 0; c < VEC; ++c) { va[g * VEC + c] = 0.0f; if (HASC) vwl[g * VEC + c] = 0.0f; }
                }
            }
        }
        env_compute_rows<NPL, HASC, VEC, TAIL, /*LOADY=*/false, true, false, NoHook, /*YSTAGE=*/true, NOSINKS>(
            p, e, lane, s0, va, vy, vwl, so, ws, stage_s + row_bytes);
        s0 = s1; s1 = s2;
    }
    if (p.stats) { wstats_store(ws, s_stats + warp * PMRL_STATS_LEN, lane); stats_flush_block(p.stats, s_stats, kStgWarps); }
}

// Single-stage shape: ONE copy of the rows per warp (half the shared memory → three CTAs per SM).  The action row is
// copied to registers at the head of an env and its buffer refilled at once with the next env's action; the previous
// weights stay in shared memory and are re-read by every mu iteration (WLS: 16 fewer live registers), and they and the
// price relatives are refilled for the next env as soon as the price relatives have been read (after the mu iteration).
struct StagedHook {
    uint64_t* bar; uint32_t parity; int lane; bool wait_here, refill;
    float* dst_y; const float* src_y; float* dst_wl; const float* src_wl; uint32_t row_bytes; uint64_t pol_y;
    __device__ __forceinline__ void before_y() const { if (wait_here) mbar_wait(bar, parity); }
    __device__ __forceinline__ void after_y() const {
        __syncwarp();
        if (refill && lane == 0) {
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(bar, (dst_wl ? 2u : 1u) * row_bytes);
            bulk_load_g2s(dst_y, src_y, row_bytes, bar, pol_y);
            if (dst_wl) bulk_load_g2s(dst_wl, src_wl, row_bytes, bar, kPolicyEvictFirst);
        }
    }
};

template <int NPL, bool HASC, bool TAIL, int kStgWarps, int MINB, bool NOSINKS, bool WLSMEM = true>
__global__ void __launch_bounds__(kStgWarps * 32, MINB) k_env_step_staged1(const StepParams p) {
    constexpr int VEC = 4;
    constexpr int R = HASC ? 3 : 2;
    constexpr bool WLS = HASC && WLSMEM;                             // previous weights re-read from shared memory by the mu iteration
    constexpr bool WLR = HASC && !WLSMEM;                            // ... or copied to registers with the action row
    extern __shared__ __align__(128) float s_rows[];                 // [warps][R][A]: action, price relatives, previous weights
    __shared__ double s_stats[kStgWarps * PMRL_STATS_LEN];
    __shared__ __align__(8) uint64_t s_bar[kStgWarps][2];            // [0]: action row, [1]: price relatives (+ previous weights)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kStgWarps + warp;
    const int nw = gridDim.x * kStgWarps;
    const int A = p.A, W = p.W;
    const uint32_t row_bytes = (uint32_t)A * 4u;
    float* const my_rows = s_rows + (size_t)warp * R * A;
    const uint32_t my_rows_s = smem_u32(my_rows);
    uint64_t* const bar_a = &s_bar[warp][0];
    uint64_t* const bar_y = &s_bar[warp][1];
    if (lane == 0) { mbar_init(bar_a, 1); mbar_init(bar_y, 1); mbar_fence_init(); }
    __syncwarp();
    WarpStats ws;
    wstats_init(ws);
    StepOut so;
    const uint64_t pol_y = p.y_ext ? kPolicyEvictFirst : kPolicyEvictLast;
    auto wl_row = [&](int e, const EnvScalars& s) { return p.hist + ((size_t)e * W + (s.i == 0 ? W : s.i) - 1) * A; };

    EnvScalars s0, s1, s2;
    int e = gw;
    if (e < p.E) {
        env_load_scalars(p, e, s0);
        if (!env_needs_reset(p, s0) && lane == 0) {
            if (!NOSINKS && p.act_ready) { env_wait_actions_lane(p, e); asm volatile("fence.proxy.async;" ::: "memory"); }
            mbar_arrive_expect_tx(bar_a, (WLR ? 2u : 1u) * row_bytes);
            bulk_load_g2s(my_rows, p.actions + (size_t)e * A, row_bytes, bar_a, kPolicyEvictFirst);
            if (WLR) bulk_load_g2s(my_rows + 2 * A, wl_row(e, s0), row_bytes, bar_a, kPolicyEvictFirst);
            mbar_arrive_expect_tx(bar_y, (WLS ? 2u : 1u) * row_bytes);
            bulk_load_g2s(my_rows + A, env_y_row(p, e, s0), row_bytes, bar_y, pol_y);
            if (WLS) bulk_load_g2s(my_rows + 2 * A, wl_row(e, s0), row_bytes, bar_y, kPolicyEvictFirst);
        }
    }
    if (e + nw < p.E) env_load_scalars(p, e + nw, s1);
    uint32_t uses = 0;                                               // both barriers complete one phase per stepped env
    for (; e < p.E; e += nw) {
        if (e + 2 * nw < p.E) env_load_scalars(p, e + 2 * nw, s2);
        const bool live = !env_needs_reset(p, s0);
        const bool next_live = (e + nw < p.E) && !env_needs_reset(p, s1);
        float va[NPL], vy[NPL], vwl[WLR ? NPL : 1];
        StagedHook hook;
        hook.bar = bar_y; hook.parity = uses & 1u; hook.lane = lane; hook.wait_here = !WLS; hook.refill = next_live;
        hook.dst_y = my_rows + A; hook.dst_wl = WLS ? my_rows + 2 * A : nullptr; hook.row_bytes = row_bytes; hook.pol_y = pol_y;
        hook.src_y = next_live ? env_y_row(p, e + nw, s1) : nullptr;
        hook.src_wl = (WLS && next_live) ? wl_row(e + nw, s1) : nullptr;
        if (live) {
            mbar_wait(bar_a, uses & 1u);
#pragma unroll
            for (int g = 0; g < NPL / VEC; ++g) {
                if (slot_ok<NPL, VEC, TAIL>(g * VEC, lane, A)) {
                    lds_v<VEC>(my_rows_s + 16u * (uint32_t)(g * 32 + lane), &va[g * VEC]);
                    if (WLR) lds_v<VEC>(my_rows_s + 2 * row_bytes + 16u * (uint32_t)(g * 32 + lane), &vwl[g * VEC]);
                } else {
#pragma unroll
                    for (int c = 0; c < VEC; ++c) { va[g * VEC + c] = 0.0f; if (WLR) vwl[g * VEC + c] = 0.0f; }
                }
            }
            if (WLS) mbar_wait(bar_y, uses & 1u);                    // the mu iteration reads the previous weights from their row
            ++uses;
        }
        __syncwarp();                                                // every lane holds its part of the action row
        if (next_live && lane == 0) {
            if (!NOSINKS && p.act_ready) { env_wait_actions_lane(p, e + nw); asm volatile("fence.proxy.async;" ::: "memory"); }
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(bar_a, (WLR ? 2u : 1u) * row_bytes);
            bulk_load_g2s(my_rows, p.actions + (size_t)(e + nw) * A, row_bytes, bar_a, kPolicyEvictFirst);
            if (WLR) bulk_load_g2s(my_rows + 2 * A, wl_row(e + nw, s1), row_bytes, bar_a, kPolicyEvictFirst);
        }
        if (!live) {                                                 // auto-reset: nothing staged was consumed, so the refill of the
            hook.after_y();                                          // y / previous-weight rows for the next env happens here
        }
        env_compute_rows<NPL, HASC, VEC, TAIL, /*LOADY=*/false, true, WLS, StagedHook, /*YSTAGE=*/true, NOSINKS>(
            p, e, lane, s0, va, vy, vwl, so, ws, my_rows_s + row_bytes, my_rows_s + 2 * row_bytes, hook);
        s0 = s1; s1 = s2;
    }
    if (p.stats) { wstats_store(ws, s_stats + warp * PMRL_STATS_LEN, lane); stats_flush_block(p.stats, s_stats, kStgWarps); }
}

}  // namespace pmrl

using namespace pmrl;

template <int NPL, bool HASC, bool TAIL, int WARPS, int MINB, int STAGES_>
static int launch_staged_t(const StepParams& p, cudaStream_t s) {
    constexpr int STAGES = STAGES_ == 2 ? 2 : 1;                      // STAGES_ == 3: single stage, previous weights in registers
    const size_t smem = (size_t)WARPS * STAGES * (HASC ? 3 : 2) * p.A * 4;
    if (smem > (size_t)224 * 1024) return -100;
    int per_sm = (int)(((size_t)226 * 1024) / (smem + 2048));
    if (per_sm > MINB) per_sm = MINB;
    void (*kern)(const StepParams);
    // (NOSINKS instantiations also drop the wait for streamed-in action rows: the two rarely-used features share one variant)
    const bool nosinks = !p.action_sink && !p.weight_sink && !p.value_sink && !p.index_sink && !p.reward_host && !p.act_ready;
    if constexpr (STAGES == 2) kern = nosinks ? k_env_step_staged<NPL, HASC, TAIL, WARPS, MINB, true> : k_env_step_staged<NPL, HASC, TAIL, WARPS, MINB, false>;
    else kern = nosinks ? k_env_step_staged1<NPL, HASC, TAIL, WARPS, MINB, true, STAGES_ == 1> : k_env_step_staged1<NPL, HASC, TAIL, WARPS, MINB, false, STAGES_ == 1>;
    static bool attr_done[64][2] = {{false}};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev][nosinks]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return pmrl_fail((int)e, "cudaFuncSetAttribute(k_env_step_staged) failed");
        attr_done[dev][nosinks] = true;
    }
    const int want = (p.E + WARPS - 1) / WARPS;
    const int cap = pmrl_sm_count() * per_sm;
    kern<<<want < cap ? want : cap, WARPS * 32, smem, s>>>(p);
    return pmrl_check_launch("k_env_step_staged");
}

// Returns -100 when the shape is not covered (caller falls back to k_env_step).  `shape` (PMRL_TUNE_CTAS_PER_SM): default = two
// 8-warp CTAs per SM with double-buffered rows (107 registers, no spills); 3 = three 8-warp CTAs with single-stage rows and
// the previous weights re-read from shared memory (80 registers).  Measured on config 5 (262,144 x 500, c = 0.0025): both
// 0.335 ms — the kernel is bound by instruction issue, and the 12 % more instructions of shape 3 cancel its better latency
// hiding (issue slots 72 % vs 65 % busy).  CTAs whose warp count is not a multiple of four per SM sub-partition (2 x 9,
// 3 x 6, 3 x 7 warps) lose 10-16 %: the envs are statically divided over the warps, so the fullest sub-partition sets the time.
int pmrl_launch_step_staged(const StepParams& p, int npl, int vec, int shape, cudaStream_t s) {
    if (npl < 8 || vec != 4 || p.A % 4 != 0) return -100;
    if (((uintptr_t)p.actions | (uintptr_t)p.hist | (uintptr_t)p.y_tm | (uintptr_t)p.y_ext) % 16 != 0) return -100;
    const bool hasc = p.commission > 0.0f;
#define STG_GO(N, WARPS, MINB, ST)                                                                                                  \
    {                                                                                                                                \
        const bool tail = env_tail_ok<N, 4>(p.A);                                                                                    \
        if (hasc) return tail ? launch_staged_t<N, true, true, WARPS, MINB, ST>(p, s) : launch_staged_t<N, true, false, WARPS, MINB, ST>(p, s);   \
        return tail ? launch_staged_t<N, false, true, WARPS, MINB, ST>(p, s) : launch_staged_t<N, false, false, WARPS, MINB, ST>(p, s);           \
    }
    if (npl == 16) {
        if (shape == 3) STG_GO(16, 8, 3, 1)
        STG_GO(16, 8, 2, 2)
    }
    if (npl == 8) STG_GO(8, 8, 2, 2)
    if (npl == 32) STG_GO(32, 8, 2, 2)
#undef STG_GO
    return -100;
}
