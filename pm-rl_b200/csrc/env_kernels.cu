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

namespace pmrl {

constexpr int kStepThreads = 256;
constexpr int kStepWarps = kStepThreads / 32;
constexpr int kObsThreads = 256;

// ------------------------------------------------------------------------------------------------
// Mode S: state-only step.
// ------------------------------------------------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(kStepThreads) k_env_step(const StepParams p) {
    __shared__ double s_stats[kStepWarps * PMRL_STATS_LEN];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kStepWarps + warp;
    const int nw = gridDim.x * kStepWarps;
    StatAcc acc;
    for (int e = gw; e < p.E; e += nw) {
        float wn[NPL];
        StepOut so;
        env_step_warp<NPL>(p, e, lane, wn, so, acc);
    }
    if (p.stats) stats_flush_block(acc, p.stats, s_stats, lane, warp, kStepWarps);
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
    if (tbl) { p.close_tm = tbl->close_tm; p.feat_am = tbl->feat_am; }
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

static int launch_obs(StepParams& p, float* obs, int obs_mode, cudaStream_t s) {
    p.obs = obs; p.obs_mode = obs_mode;
    choose_obs_tile(p.A, p.W, p.F, kObsTileCapBytes, obs, p);
    if (p.E == 0) return 0;
    const size_t smem = (size_t)p.tile_assets * p.W * p.F * 4;
    const long grid = (long)p.E * p.tiles_per_env;
    if (grid > 2147483647L) return pmrl_fail(PMRL_E_SHAPE, "E * tiles_per_env exceeds the grid limit");
    k_obs_build<<<(unsigned)grid, kObsThreads, smem, s>>>(p);
    return pmrl_check_launch("k_obs_build");
}

template <int NPL>
static int launch_step_s(const StepParams& p, cudaStream_t s) {
    const int want = (p.E + kStepWarps - 1) / kStepWarps;
    const int cap = pmrl_sm_count() * 8;
    k_env_step<NPL><<<want < cap ? want : cap, kStepThreads, 0, s>>>(p);
    return pmrl_check_launch("k_env_step");
}

extern "C" int pmrl_env_reset(const PmrlEnvCfg* cfg, const PmrlTables* tbl, const PmrlEnvState* st,
                              const uint8_t* mask, float* obs, int32_t obs_mode, void* stream) {
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

extern "C" int pmrl_obs_build(const PmrlEnvCfg* cfg, const PmrlTables* tbl, const PmrlEnvState* st,
                              float* obs, int32_t obs_mode, void* stream) {
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
    StepParams p;
    if (int rc = fill_params(cfg, tbl, st, p)) return rc;
    if (!actions || !reward || !done) return pmrl_fail(PMRL_E_ARG, "actions/reward/done is NULL");
    if (!y_ext) {
        if (!p.close_tm) return pmrl_fail(PMRL_E_ARG, "need close_tm or y_ext");
        if (!p.t0) return pmrl_fail(PMRL_E_ARG, "t0 is NULL but y comes from close_tm");
        if (p.episode_len <= 0) return pmrl_fail(PMRL_E_SHAPE, "episode_len must be > 0 when y comes from close_tm");
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
    int rc = 0;
    switch (npl) {
        case 1: rc = launch_step_s<1>(p, s); break;
        case 2: rc = launch_step_s<2>(p, s); break;
        case 4: rc = launch_step_s<4>(p, s); break;
        case 8: rc = launch_step_s<8>(p, s); break;
        case 16: rc = launch_step_s<16>(p, s); break;
        default: rc = launch_step_s<32>(p, s); break;
    }
    if (rc) return rc;
    if (obs_mode != PMRL_OBS_NONE) return launch_obs(p, obs, obs_mode, s);
    return 0;
}
