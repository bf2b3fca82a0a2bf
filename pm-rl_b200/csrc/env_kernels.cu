// env_kernels.cu — environment reset / step / observation kernels + their C-ABI entry points.
//
// Kernels (sm_100a):
//   k_env_step<NPL>      state-only transition ("Mode S"): persistent grid of warps, one env per warp.
//   k_env_reset          masked re-initialisation, one env per warp.
//   k_obs_build          one CTA per (env, asset-tile): gather window + weight channel into a shared
//                        tile that is the byte image of obs[e, a0:a0+na, :, :], then one TMA bulk store.
//   k_env_step_obs<NPL>  fused "Mode O": persistent CTAs, step math + obs tiles in one pass (see below).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include "pmrl_b200.h"
#include "pmrl_device.cuh"
#include "env_step.cuh"
#include "obs_tile.cuh"
#include "host_util.h"
#include "env_launch.h"

namespace pmrl {

constexpr int kStepThreads = 256;
constexpr int kStepWarps = kStepThreads / 32;
constexpr int kObsThreads = 256;

// ------------------------------------------------------------------------------------------------
// Mode S: state-only step.
// ------------------------------------------------------------------------------------------------
template <int NPL, bool HASC, bool PIPE, bool TAIL>
__global__ void __launch_bounds__(kStepThreads) k_env_step(const StepParams p) {
    __shared__ double s_stats[kStepWarps * PMRL_STATS_LEN];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kStepWarps + warp;
    const int nw = gridDim.x * kStepWarps;
    if (p.stats) stats_init_block(s_stats, kStepWarps);
    double* const acc = s_stats + warp * PMRL_STATS_LEN;
    StepOut so;
    if constexpr (PIPE) {
        // three envs of this warp in flight: scalars of e+2nw, vectors of e+nw, arithmetic of e
        EnvScalars s0, s1, s2;
        EnvVectors<NPL, HASC> v0, v1;
        int e = gw;
        if (e < p.E) { env_load_scalars(p, e, s0); env_load_vectors<NPL, HASC, TAIL>(p, e, lane, s0, v0); }
        if (e + nw < p.E) env_load_scalars(p, e + nw, s1);
        for (; e < p.E; e += nw) {
            if (e + 2 * nw < p.E) env_load_scalars(p, e + 2 * nw, s2);
            if (e + nw < p.E) env_load_vectors<NPL, HASC, TAIL>(p, e + nw, lane, s1, v1);
            env_compute_store<NPL, HASC, TAIL>(p, e, lane, s0, v0, so, acc);
            s0 = s1; s1 = s2; v0 = v1;
        }
    } else {
        // wide envs (NPL >= 8): no registers for a second env's vectors.  Keep the scalars of the next two envs in
        // registers and pull the next env's DRAM rows into L2 while this one computes: the dependent chain
        // scalars → row addresses → vectors then costs one L2 hit per env instead of two DRAM round trips.
        EnvScalars s0, s1, s2;
        int e = gw;
        if (e < p.E) env_load_scalars(p, e, s0);
        if (e + nw < p.E) env_load_scalars(p, e + nw, s1);
        for (; e < p.E; e += nw) {
            EnvVectors<NPL, HASC> v;
            if (e + 2 * nw < p.E) env_load_scalars(p, e + 2 * nw, s2);
            env_load_vectors<NPL, HASC, TAIL>(p, e, lane, s0, v);
            if (e + nw < p.E) env_prefetch_vectors<HASC>(p, e + nw, lane, s1);
            env_compute_store<NPL, HASC, TAIL>(p, e, lane, s0, v, so, acc);
            s0 = s1; s1 = s2;
        }
    }
    if (p.stats) stats_flush_block(p.stats, s_stats, kStepWarps);
}

// ------------------------------------------------------------------------------------------------
// Masked reset (trading_env.py:28-29, weight_buffer.py:46-50).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStepThreads) k_env_reset(const StepParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kStepWarps + warp;
    const int nw = gridDim.x * kStepWarps;
    for (int e = gw; e < p.E; e += nw) {
        if (p.mask && !p.mask[e]) continue;
        ring_reset_warp(p.hist + (size_t)e * p.W * p.A, p.W, p.A, lane);
        if (lane == 0) {
            p.value[e] = p.initial_cash;
            p.idx[e] = 1; p.is_full[e] = 0; p.t[e] = 0;
            if (p.sharpe) { p.sharpe[3 * (size_t)e] = 0.0; p.sharpe[3 * (size_t)e + 1] = 0.0; p.sharpe[3 * (size_t)e + 2] = 0.0; }
            if (p.ep_return) p.ep_return[e] = 0.0f;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Observation tiles from the current state.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kObsThreads) k_obs_build(const StepParams p) {
    extern __shared__ __align__(128) float tile[];
    const int e = blockIdx.x / p.tiles_per_env;
    const int ti = blockIdx.x - e * p.tiles_per_env;
    if (p.mask && !p.mask[e]) return;
    const int a0 = ti * p.tile_assets;
    const int na = min(p.tile_assets, p.A - a0);
    const int tid = threadIdx.x;
    const int idx = p.idx[e], full = p.is_full[e];
    if (p.obs_mode == PMRL_OBS_FULL) {
        const int row0 = (p.t0 ? p.t0[e] : 0) + p.t[e];
        obs_tile_fill_features(p, tile, a0, na, row0, tid, kObsThreads);
    }
    obs_tile_fill_weights(p, tile, p.hist + (size_t)e * p.W * p.A, a0, na, idx, full, nullptr, -1, tid, kObsThreads);
    fence_proxy_async_smem();
    __syncthreads();
    obs_tile_store(p, tile, e, a0, na, tid, kObsThreads);
    if (p.obs_mode == PMRL_OBS_FULL && p.obs_bulk_ok && tid == 0) bulk_wait_read<0>();
}


// The same tile for the common shape (4 feature channels + weight slot, W <= 64, 32-asset tiles), without the
// per-element index divisions of the generic fill: a warp owns an asset row of the feature window (lanes over window
// rows, one float4 per row), and lanes run over assets for the weight channel (coalesced 128-byte ring rows).  All
// loads of a thread are issued before the first shared store; the L2-resident table loads go first so that no
// DRAM-latency ring load sits ahead of them in the L1 return queue.
__global__ void __launch_bounds__(kObsThreads) k_obs_build_rows(const StepParams p) {
    extern __shared__ __align__(128) float tile[];                  // [32][W][5]
    const int A = p.A, W = p.W, T = p.T;
    const int e = blockIdx.x / p.tiles_per_env;
    const int ti = blockIdx.x - e * p.tiles_per_env;
    if (p.mask && !p.mask[e]) return;
    const int a0 = ti * 32;
    const int na = min(32, A - a0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int idx = p.idx[e], full = p.is_full[e];
    const int row0 = p.t0[e] + p.t[e];
    const int shift = full ? 0 : (W - idx);                         // weight_buffer.py:38-42
    const bool w0 = lane < W, w1 = lane + 32 < W;
    const float4* __restrict__ tbl = reinterpret_cast<const float4*>(p.feat_am) + ((size_t)a0 * T + row0 + lane);
    float4 fv[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int al = warp + 8 * i;
        if (al < na) {
            const float4* __restrict__ src = tbl + (size_t)al * T;
            if (w0) fv[i][0] = ld_keep4(src, kPolicyEvictLast);
            if (w1) fv[i][1] = ld_keep4(src + 32, kPolicyEvictLast);
        }
    }
    const float* __restrict__ ring = p.hist + (size_t)e * W * A + a0 + lane;
    float rv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int w = warp + 8 * j, slot = w - shift;
        rv[j] = (w < W && slot >= 0 && lane < na) ? ld_once(ring + (size_t)slot * A, kPolicyEvictFirst) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int al = warp + 8 * i;
        if (al < na) {
            float* d = tile + (al * W + lane) * 5;
            if (w0) { const float4 v = fv[i][0]; d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
            if (w1) { const float4 v = fv[i][1]; d[160] = v.x; d[161] = v.y; d[162] = v.z; d[163] = v.w; }
        }
    }
    if (lane < na) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int w = warp + 8 * j;
            if (w < W) tile[(lane * W + w) * 5 + 4] = rv[j];
        }
    }
    fence_proxy_async_smem();
    __syncthreads();
    float* __restrict__ dst = p.obs + ((size_t)e * A + a0) * W * 5;
    const int n = na * W * 5;
    if (((((uintptr_t)dst) | ((size_t)n * 4)) & 15) == 0) {
        if (tid == 0) { bulk_store_s2g(dst, tile, (uint32_t)n * 4u, kPolicyEvictFirst); bulk_commit(); bulk_wait_read<0>(); }
    } else {
        for (int q = tid; q < n; q += kObsThreads) dst[q] = tile[q];
    }
}

// ------------------------------------------------------------------------------------------------
// Mode O, fused: step + observation in one pass.
//
// A CTA owns a *group* of G consecutive envs at a time (persistent, grid-stride over groups):
//   phase 1  warp w advances env e0+w (env_step_warp) and leaves w', the new ring pointer and the
//            window row in shared memory;
//   phase 2  all warps stream the group's observations.  The obs of G consecutive envs is one
//            contiguous [G*A, W, F] slab, so tiles are TA asset-rows cut across env boundaries (no
//            ragged per-env tiles); each tile is assembled in one of two shared buffers and leaves
//            as a single TMA bulk store while the next tile is being filled.
// The row written by this step comes from shared memory (never re-read from global), so the ring is
// read exactly once and written exactly once per step: 4·W + 4 bytes per asset-step of ring traffic.
// ------------------------------------------------------------------------------------------------
constexpr int kFusedThreads = 256;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kMaxGroup = kFusedWarps;

struct GroupEnv { int row0, shift, fresh_slot, pad; };   // per env of the group (phase 1 → phase 2)

template <int NPL, bool HASC, int MINB>
__global__ void __launch_bounds__(kFusedThreads, MINB) k_env_step_obs(const StepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_stats[kFusedWarps * PMRL_STATS_LEN];
    __shared__ GroupEnv s_env[kMaxGroup];
    const int A = p.A, W = p.W, F = p.F, T = p.T, TA = p.tile_assets, G = p.group_envs;
    const int tile_floats = TA * W * F;
    float* const tile0 = reinterpret_cast<float*>(smem_raw);
    float* const tile1 = tile0 + tile_floats;
    float* const s_wnew = tile1 + tile_floats;           // [G, A]   w' of the group's envs, indexed by asset-row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_groups = (p.E + G - 1) / G;
    const size_t row_floats = (size_t)W * F;
    if (p.stats) stats_init_block(s_stats, kFusedWarps);
    const uint64_t pol_keep = kPolicyEvictLast, pol_once = kPolicyEvictFirst;
    int buf = 0;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int e0 = grp * G;
        const int ne = min(G, p.E - e0);
        // ---------------- phase 1: one warp per env ----------------
        if (warp < ne) {
            const int e = e0 + warp;
            EnvVectors<NPL, HASC> ev;
            StepOut so;
            env_step_warp<NPL, HASC>(p, e, lane, ev, so, s_stats + warp * PMRL_STATS_LEN);
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const int a = lane + 32 * j;
                if (a < A) s_wnew[warp * A + a] = ev.a[j];
            }
            if (lane == 0) {
                GroupEnv ge;
                ge.row0 = p.t0[e] + so.k;
                ge.shift = so.is_full ? 0 : (W - so.idx_new);            // weight_buffer.py:38-42
                ge.fresh_slot = so.did_reset ? 0 : so.slot_written;       // rows written by this launch come from smem
                ge.pad = 0;
                s_env[warp] = ge;
            }
        }
        __syncthreads();
        // ---------------- phase 2: tiles over the group's [ne*A] asset-rows ----------------
        const int R = ne * A;
        const int ntiles = (R + TA - 1) / TA;
        float* const obs_grp = p.obs + (size_t)e0 * A * row_floats;

        for (int ti = 0; ti < ntiles; ++ti) {
            float* const tile = buf ? tile1 : tile0;
            if (tid == 0) bulk_wait_read<1>();            // the store that last used this buffer has drained
            __syncthreads();
            const int r0 = ti * TA;
            const int nr = min(TA, R - r0);
            // -- feature channels: one warp per asset-row, lanes over the window rows --
            if (F - 1 == 4) {
                const float4* __restrict__ tbl = reinterpret_cast<const float4*>(p.feat_am);
                for (int ar = warp; ar < nr; ar += kFusedWarps) {
                    const int gar = r0 + ar;
                    const int el = gar / A, a = gar - el * A;
                    const float4* __restrict__ src = tbl + (size_t)a * T + s_env[el].row0;
                    float* __restrict__ dst = tile + (size_t)ar * W * 5;
                    for (int w = lane; w < W; w += 32) {
                        const float4 v = ld_keep4(src + w, pol_keep);
                        float* d = dst + w * 5;
                        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                    }
                }
            } else {
                const int Fm1 = F - 1, per = W * Fm1;
                for (int ar = warp; ar < nr; ar += kFusedWarps) {
                    const int gar = r0 + ar;
                    const int el = gar / A, a = gar - el * A;
                    const float* __restrict__ src = p.feat_am + ((size_t)a * T + s_env[el].row0) * Fm1;
                    float* __restrict__ dst = tile + (size_t)ar * W * F;
                    for (int q = lane; q < per; q += 32) {
                        const int w = q / Fm1, c = q - w * Fm1;
                        dst[w * F + c] = ld_keep(src + q, pol_keep);
                    }
                }
            }
            // -- weight channel: lane = asset-row (coalesced ring rows), warps over the window columns --
            for (int ar = lane; ar < nr; ar += 32) {
                const int gar = r0 + ar;
                const int el = gar / A, a = gar - el * A;
                const GroupEnv ge = s_env[el];
                const float* __restrict__ hist_ea = p.hist + ((size_t)(e0 + el) * W) * A + a;
                const float fresh = s_wnew[el * A + a];
                float* __restrict__ dst = tile + (size_t)ar * W * F + (F - 1);
#pragma unroll 4
                for (int w = warp; w < W; w += kFusedWarps) {
                    const int slot = w - ge.shift;
                    float v = 0.0f;
                    if (slot >= 0) v = (slot == ge.fresh_slot) ? fresh : ld_once(hist_ea + (size_t)slot * A, pol_once);
                    dst[w * F] = v;
                }
            }
            fence_proxy_async_smem();
            __syncthreads();
            // -- tile → global --
            float* const gdst = obs_grp + (size_t)r0 * row_floats;
            const int n = nr * W * F;
            if (((((uintptr_t)gdst) | ((size_t)n * 4)) & 15) == 0) {
                if (tid == 0) { bulk_store_s2g(gdst, tile, (uint32_t)n * 4u, pol_once); bulk_commit(); }
            } else {
                for (int q = tid; q < n; q += kFusedThreads) gdst[q] = tile[q];
            }
            buf ^= 1;
        }
    }
    if (tid == 0) bulk_wait_read<0>();
    if (p.stats) stats_flush_block(p.stats, s_stats, kFusedWarps);
}

}  // namespace pmrl

// ================================================================================================
// Host side: validation, tiling choice, launches.
// ================================================================================================
using namespace pmrl;

static int npl_for(int A) {
    if (A <= 32) return 1;
    if (A <= 64) return 2;
    if (A <= 128) return 4;
    if (A <= 256) return 8;
    if (A <= 512) return 16;
    if (A <= 1024) return 32;
    return 0;
}

// Pick the asset-tile of the obs kernels: the biggest tile ≤ cap bytes whose start offsets stay
// 16-byte aligned (so the TMA bulk store applies) with the least ragged last tile.
static void choose_obs_tile(int A, int W, int F, size_t cap_bytes, const void* obs, StepParams& p) {
    const size_t per = (size_t)W * F * 4;
    int max_ta = (int)(cap_bytes / per);
    if (max_ta > A) max_ta = A;
    if (max_ta < 1) max_ta = 1;
    const bool env_aligned = (((size_t)A * W * F) % 4 == 0) && (((uintptr_t)obs) % 16 == 0);
    int best = -1; long best_waste = 0;
    if (env_aligned) {
        for (int pass = 0; pass < 2 && best < 0; ++pass) {
            const int lo = pass == 0 ? (max_ta + 1) / 2 : 1;
            for (int ta = max_ta; ta >= lo; --ta) {
                if (((size_t)ta * W * F) % 4 != 0 && ta < A) continue;
                const long tiles = (A + ta - 1) / ta;
                const long waste = tiles * ta - A;
                if (best < 0 || waste < best_waste) { best = ta; best_waste = waste; }
            }
        }
    }
    if (best > 0) { p.tile_assets = best; p.obs_bulk_ok = 1; }
    else { p.tile_assets = max_ta; p.obs_bulk_ok = 0; }
    p.tiles_per_env = (A + p.tile_assets - 1) / p.tile_assets;
}

static int fill_params(const PmrlEnvCfg* cfg, const PmrlTables* tbl, const PmrlEnvState* st, StepParams& p) {
    if (!cfg || !st) return pmrl_fail(PMRL_E_ARG, "cfg/state is NULL");
    if (cfg->E < 0 || cfg->A < 1 || cfg->W < 1 || cfg->F < 1) return pmrl_fail(PMRL_E_SHAPE, "E>=0, A>=1, W>=1, F>=1 required");
    if (!st->value || !st->hist || !st->idx || !st->is_full || !st->t) return pmrl_fail(PMRL_E_ARG, "state pointer is NULL");
    memset(&p, 0, sizeof(p));
    p.E = cfg->E; p.A = cfg->A; p.W = cfg->W; p.F = cfg->F; p.T = cfg->T;
    p.episode_len = cfg->episode_len; p.reward_mode = cfg->reward_mode;
    p.mu_max_iter = cfg->mu_max_iter > 0 ? cfg->mu_max_iter : 16;
    p.flags = cfg->flags;
    p.initial_cash = cfg->initial_cash; p.commission = cfg->commission;
    p.reward_scale = cfg->reward_scale; p.risk_free = cfg->risk_free;
    const double c = (double)cfg->commission;
    p.mu0 = (float)(1.0 - 2.0 * c + c * c);
    p.c2 = (float)(2.0 * c - c * c);
    if (tbl) { p.y_tm = tbl->y_tm; p.feat_am = tbl->feat_am; }
    p.value = st->value; p.hist = st->hist; p.idx = st->idx; p.is_full = st->is_full; p.t = st->t;
    p.t0 = st->t0; p.sharpe = st->sharpe; p.ep_return = st->ep_return;
    return 0;
}

static int check_obs_args(const StepParams& p, const float* obs, int obs_mode) {
    if (obs_mode != PMRL_OBS_NONE && obs_mode != PMRL_OBS_FULL && obs_mode != PMRL_OBS_WEIGHTS)
        return pmrl_fail(PMRL_E_ARG, "bad obs_mode");
    if (obs_mode == PMRL_OBS_NONE) return 0;
    if (!obs) return pmrl_fail(PMRL_E_ARG, "obs is NULL but obs_mode != NONE");
    if (p.F < 2 && obs_mode == PMRL_OBS_FULL) return pmrl_fail(PMRL_E_SHAPE, "F >= 2 required for a full obs");
    if (obs_mode == PMRL_OBS_FULL) {
        if (!p.feat_am) return pmrl_fail(PMRL_E_ARG, "feat_am is NULL but obs_mode == FULL");
        if (p.F - 1 == 4 && ((uintptr_t)p.feat_am) % 16 != 0) return pmrl_fail(PMRL_E_ALIGN, "feat_am must be 16-byte aligned");
        if (p.T < p.W) return pmrl_fail(PMRL_E_SHAPE, "T < W");
    }
    if ((size_t)p.W * p.F * 4 > kObsTileCapBytes) return pmrl_fail(PMRL_E_SHAPE, "W*F*4 exceeds the obs tile capacity");
    return 0;
}

// Launch-shape knobs (pmrl_set_tuning).
static int g_tune_rows = 0, g_tune_group = 0, g_tune_ctas_per_sm = 0, g_tune_fused = 1, g_tune_fast = 1, g_tune_tma = 0, g_tune_stages = 0, g_tune_rt = 1, g_tune_tm = 0;

static int launch_obs(StepParams& p, float* obs, int obs_mode, cudaStream_t s) {
    p.obs = obs; p.obs_mode = obs_mode;
    if (obs_mode == PMRL_OBS_FULL && p.F == 5 && p.W <= 64 && p.t0 && g_tune_fast) {      // division-free tile kernel
        if (p.E == 0) return 0;
        p.tile_assets = 32;
        p.tiles_per_env = (p.A + 31) / 32;
        const long grid = (long)p.E * p.tiles_per_env;
        if (grid > 2147483647L) return pmrl_fail(PMRL_E_SHAPE, "E * tiles_per_env exceeds the grid limit");
        k_obs_build_rows<<<(unsigned)grid, kObsThreads, (size_t)32 * p.W * 5 * 4, s>>>(p);
        return pmrl_check_launch("k_obs_build_rows");
    }
    choose_obs_tile(p.A, p.W, p.F, kObsTileCapBytes, obs, p);
    if (p.E == 0) return 0;
    const size_t smem = (size_t)p.tile_assets * p.W * p.F * 4;
    const long grid = (long)p.E * p.tiles_per_env;
    if (grid > 2147483647L) return pmrl_fail(PMRL_E_SHAPE, "E * tiles_per_env exceeds the grid limit");
    k_obs_build<<<(unsigned)grid, kObsThreads, smem, s>>>(p);
    return pmrl_check_launch("k_obs_build");
}

template <int NPL, bool HASC>
static int launch_step_s(const StepParams& p, cudaStream_t s) {
    constexpr bool PIPE = (NPL <= 4);                 // 3-stage software pipeline while the registers allow it
    const int want = (p.E + kStepWarps - 1) / kStepWarps;
    const int cap = pmrl_sm_count() * 8;
    // TAIL: every 32-asset slot row but the last is full → only the last row carries validity guards
    if (p.A > 32 * (NPL - 1)) k_env_step<NPL, HASC, PIPE, true><<<want < cap ? want : cap, kStepThreads, 0, s>>>(p);
    else k_env_step<NPL, HASC, PIPE, false><<<want < cap ? want : cap, kStepThreads, 0, s>>>(p);
    return pmrl_check_launch("k_env_step");
}

template <bool HASC>
static int launch_step_npl(const StepParams& p, int npl, cudaStream_t s) {
    switch (npl) {
        case 1: return launch_step_s<1, HASC>(p, s);
        case 2: return launch_step_s<2, HASC>(p, s);
        case 4: return launch_step_s<4, HASC>(p, s);
        case 8: return launch_step_s<8, HASC>(p, s);
        case 16: return launch_step_s<16, HASC>(p, s);
        default: return launch_step_s<32, HASC>(p, s);
    }
}


extern "C" int pmrl_set_tuning(int32_t key, int32_t value) {
    switch (key) {
        case PMRL_TUNE_TILE_ROWS: g_tune_rows = value; return 0;
        case PMRL_TUNE_GROUP_ENVS: g_tune_group = value; return 0;
        case PMRL_TUNE_CTAS_PER_SM: g_tune_ctas_per_sm = value; return 0;
        case PMRL_TUNE_FUSED: g_tune_fused = value; return 0;
        case PMRL_TUNE_FAST_FILL: g_tune_fast = value; return 0;
        case PMRL_TUNE_TMA_PIPELINE: g_tune_tma = value; return 0;
        case PMRL_TUNE_TMA_STAGES: g_tune_stages = value; return 0;
        case PMRL_TUNE_FAST_VARIANT: pmrl_set_fast_variant(value); return 0;
        case PMRL_TUNE_RING_TMA: g_tune_rt = value; return 0;
        case PMRL_TUNE_TENSORMAP: g_tune_tm = value; return 0;
        default: return pmrl_fail(PMRL_E_ARG, "unknown tuning key");
    }
}

template <int NPL, bool HASC, int MINB>
static int launch_fused_t(StepParams& p, size_t smem, int grid, cudaStream_t s) {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_env_step_obs<NPL, HASC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return pmrl_fail((int)e, "cudaFuncSetAttribute(k_env_step_obs) failed");
        attr_done[dev] = true;
    }
    k_env_step_obs<NPL, HASC, MINB><<<grid, kFusedThreads, smem, s>>>(p);
    return pmrl_check_launch("k_env_step_obs");
}

template <bool HASC>
static int launch_fused_npl(StepParams& p, size_t smem, int grid, int npl, cudaStream_t s) {
    switch (npl) {
        case 1: return launch_fused_t<1, HASC, 3>(p, smem, grid, s);
        case 2: return launch_fused_t<2, HASC, 3>(p, smem, grid, s);
        case 4: return launch_fused_t<4, HASC, 3>(p, smem, grid, s);
        case 8: return launch_fused_t<8, HASC, 2>(p, smem, grid, s);
        case 16: return launch_fused_t<16, HASC, 1>(p, smem, grid, s);
        default: return launch_fused_t<32, HASC, 1>(p, smem, grid, s);
    }
}

static int launch_fused(StepParams& p, float* obs, int npl, cudaStream_t s) {
    p.obs = obs; p.obs_mode = PMRL_OBS_FULL;
    if (g_tune_fast && g_tune_tma) {                  // warp-specialised TMA pipeline (env_step_tma.cu)
        const int rc = pmrl_launch_step_obs_tma(p, npl, g_tune_stages, g_tune_group, s);
        if (rc != -100) return rc;                    // -100: shape not covered by that variant → fall through
    }
    if (g_tune_fast && g_tune_tm) {                   // all loads through TMA: tensor-map feature boxes + ring bulk loads (env_step_tm.cu)
        const int rc = pmrl_launch_step_obs_tm(p, npl, g_tune_group, g_tune_ctas_per_sm, s);
        if (rc != -100) return rc;
    }
    if (g_tune_fast && g_tune_rt) {                   // ring rows through one TMA bulk load per env (env_step_rt.cu)
        p.tma_stages = (g_tune_rt == 2) ? 0 : 1;      // RT reuses this field: 1 = prefetch the next group's phase-1 inputs into L2
        const int rc = pmrl_launch_step_obs_rt(p, npl, g_tune_group, g_tune_ctas_per_sm, s);
        if (rc != -100) return rc;
    }
    if (g_tune_fast && (g_tune_rows <= 0 || g_tune_rows == 32)) {   // register-staged fill (env_step_fast.cu): F == 5, W <= 64, A <= 128
        const int rc = pmrl_launch_step_obs_fast(p, npl, g_tune_group, g_tune_ctas_per_sm, s);
        if (rc != -100) return rc;
    }
    // No specialised kernel covers this shape.  The generic fused kernel below loses to the two-kernel path (state-only
    // step, then k_obs_build) on every shape measured — 32,768 x 500 x 50: 20.9 ms vs 4.5 ms; 131,072 x 100 x 50 with the
    // specialised kernels disabled: 6.9 ms vs 3.8 ms — so it only runs when asked for (PMRL_TUNE_FUSED = 2).
    if (g_tune_fused != 2) return -100;
    const size_t row_bytes = (size_t)p.W * p.F * 4;
    int rows = g_tune_rows > 0 ? g_tune_rows : 32;
    while (rows > 1 && rows * row_bytes > 36 * 1024) rows >>= 1;
    p.tile_assets = rows;
    int ctas_per_sm = g_tune_ctas_per_sm > 0 ? g_tune_ctas_per_sm : (npl <= 4 ? 3 : (npl <= 8 ? 2 : 1));
    const int slots = pmrl_sm_count() * ctas_per_sm;
    int G = g_tune_group > 0 ? g_tune_group : kMaxGroup;
    if (G > kMaxGroup) G = kMaxGroup;
    while (G > 1 && (p.E + G - 1) / G < 2 * slots) G >>= 1;     // small batches: more, smaller groups
    p.group_envs = G;
    const size_t smem = 2 * rows * row_bytes + (size_t)G * p.A * 4;
    if (smem > 200 * 1024) return pmrl_fail(PMRL_E_SHAPE, "fused step: shared-memory budget exceeded");
    const int n_groups = (p.E + G - 1) / G;
    const int grid = n_groups < slots ? n_groups : slots;
    return p.commission > 0.0f ? launch_fused_npl<true>(p, smem, grid, npl, s) : launch_fused_npl<false>(p, smem, grid, npl, s);
}

extern "C" int pmrl_env_reset(const PmrlEnvCfg* cfg, const PmrlTables* tbl, const PmrlEnvState* st,
                              const uint8_t* mask, float* obs, int32_t obs_mode, void* stream) {
    if (cfg && cfg->E == 0) return 0;                  // empty batch: nothing to do (state pointers may be NULL)
    StepParams p;
    if (int rc = fill_params(cfg, tbl, st, p)) return rc;
    if (int rc = check_obs_args(p, obs, obs_mode)) return rc;
    if (obs_mode == PMRL_OBS_FULL && !p.t0) return pmrl_fail(PMRL_E_ARG, "t0 is NULL but obs_mode == FULL");
    p.mask = mask;
    cudaStream_t s = (cudaStream_t)stream;
    if (p.E == 0) return 0;
    const int want = (p.E + kStepWarps - 1) / kStepWarps;
    const int cap = pmrl_sm_count() * 8;
    k_env_reset<<<want < cap ? want : cap, kStepThreads, 0, s>>>(p);
    if (int rc = pmrl_check_launch("k_env_reset")) return rc;
    if (obs_mode != PMRL_OBS_NONE) return launch_obs(p, obs, obs_mode, s);
    return 0;
}

// y_tm[t, a] = close_tm[t, a] / close_tm[t-1, a] (data/instrument.py:79), row 0 = 1: the IEEE division of the reference,
// done once per table instead of once per env-step.
__global__ void k_price_relatives(const float* __restrict__ close_tm, int T, int A, float* __restrict__ y_tm) {
    const size_t n = (size_t)T * A;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y_tm[i] = (i < (size_t)A) ? 1.0f : __fdiv_rn(close_tm[i], close_tm[i - A]);
}

extern "C" int pmrl_price_relatives(const float* close_tm, int32_t T, int32_t A, float* y_tm, void* stream) {
    if (!close_tm || !y_tm) return pmrl_fail(PMRL_E_ARG, "price_relatives: null pointer");
    if (T < 1 || A < 1) return pmrl_fail(PMRL_E_SHAPE, "price_relatives: T >= 1, A >= 1 required");
    const size_t n = (size_t)T * A;
    const size_t want = (n + 255) / 256, cap = (size_t)pmrl_sm_count() * 16;
    k_price_relatives<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(close_tm, T, A, y_tm);
    return pmrl_check_launch("k_price_relatives");
}

// Bit-for-bit comparison of the shared-reciprocal quotient (pmrl_device.cuh: unidiv) with IEEE division on caller-supplied
// operands: den[i / 32] divides num[i]; pairs outside the range the kernels accept are skipped and counted separately.
__global__ void k_selftest_division(const float* __restrict__ num, const float* __restrict__ den, long long n,
                                    unsigned long long* __restrict__ out /* [2]: mismatches, pairs tested */) {
    unsigned long long bad = 0, tested = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = num[i], b = den[i >> 5];
        if (!unidiv_in_range(b) || !(a == 0.0f || unidiv_in_range(a))) continue;
        const UniDiv d = unidiv_make(b);
        const float q = unidiv(a, d), ref = __fdiv_rn(a, b);
        ++tested;
        if (__float_as_uint(q) != __float_as_uint(ref)) ++bad;
    }
    if (bad) atomicAdd(out, bad);
    if (tested) atomicAdd(out + 1, tested);
}

extern "C" int pmrl_selftest_division(const float* num, const float* den, int64_t n, uint64_t* out, void* stream) {
    if (!num || !den || !out) return pmrl_fail(PMRL_E_ARG, "selftest_division: null pointer");
    if (n < 0) return pmrl_fail(PMRL_E_SHAPE, "selftest_division: n < 0");
    if (n == 0) return 0;
    k_selftest_division<<<pmrl_sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(num, den, (long long)n, (unsigned long long*)out);
    return pmrl_check_launch("k_selftest_division");
}

extern "C" int pmrl_obs_build(const PmrlEnvCfg* cfg, const PmrlTables* tbl, const PmrlEnvState* st,
                              float* obs, int32_t obs_mode, void* stream) {
    if (cfg && cfg->E == 0) return 0;                  // empty batch: nothing to do (state pointers may be NULL)
    StepParams p;
    if (int rc = fill_params(cfg, tbl, st, p)) return rc;
    if (obs_mode == PMRL_OBS_NONE) return pmrl_fail(PMRL_E_ARG, "obs_mode NONE makes no obs");
    if (int rc = check_obs_args(p, obs, obs_mode)) return rc;
    if (obs_mode == PMRL_OBS_FULL && !p.t0) return pmrl_fail(PMRL_E_ARG, "t0 is NULL but obs_mode == FULL");
    return launch_obs(p, obs, obs_mode, (cudaStream_t)stream);
}

extern "C" int pmrl_env_step(const PmrlEnvCfg* cfg, const PmrlTables* tbl, const PmrlEnvState* st,
                             const float* actions, const float* y_ext,
                             float* reward, uint8_t* done, float* obs, int32_t obs_mode,
                             double* stats, void* stream) {
    if (cfg && cfg->E == 0) return 0;                  // empty batch: nothing to do (state pointers may be NULL)
    StepParams p;
    if (int rc = fill_params(cfg, tbl, st, p)) return rc;
    if (!actions || !reward || !done) return pmrl_fail(PMRL_E_ARG, "actions/reward/done is NULL");
    if (!y_ext) {
        if (!p.y_tm) return pmrl_fail(PMRL_E_ARG, "need y_tm (pmrl_price_relatives) or y_ext");
        if (!p.t0) return pmrl_fail(PMRL_E_ARG, "t0 is NULL but y comes from the price table");
        if (p.episode_len <= 0) return pmrl_fail(PMRL_E_SHAPE, "episode_len must be > 0 when y comes from the price table");
    }
    if (cfg->reward_mode < 0 || cfg->reward_mode > PMRL_REWARD_SHARPE) return pmrl_fail(PMRL_E_ARG, "bad reward_mode");
    if (cfg->reward_mode == PMRL_REWARD_SHARPE && !p.sharpe) return pmrl_fail(PMRL_E_ARG, "sharpe state is NULL");
    if (stats && !p.ep_return) return pmrl_fail(PMRL_E_ARG, "ep_return is NULL but stats requested");
    if (int rc = check_obs_args(p, obs, obs_mode)) return rc;
    if (obs_mode == PMRL_OBS_FULL && !p.t0) return pmrl_fail(PMRL_E_ARG, "t0 is NULL but obs_mode == FULL");
    const int npl = npl_for(p.A);
    if (!npl) return pmrl_fail(PMRL_E_SHAPE, "A > 1024 is not supported");
    p.actions = actions; p.y_ext = y_ext; p.reward = reward; p.done = done; p.stats = stats;
    cudaStream_t s = (cudaStream_t)stream;
    if (p.E == 0) return 0;
    if (obs_mode == PMRL_OBS_FULL && g_tune_fused && (size_t)p.W * p.F * 4 <= 36 * 1024) {
        const int rc = launch_fused(p, obs, npl, s);
        if (rc != -100) return rc;                    // -100: no fused kernel for this shape → step kernel + obs kernel
    }
    int rc = p.commission > 0.0f ? launch_step_npl<true>(p, npl, s) : launch_step_npl<false>(p, npl, s);
    if (rc) return rc;
    if (obs_mode != PMRL_OBS_NONE) return launch_obs(p, obs, obs_mode, s);
    return 0;
}
