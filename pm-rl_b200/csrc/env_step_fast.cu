// env_step_fast.cu — fused step + observation kernel, register-staged fill (the default Mode-O path for F == 5, W <= 64).
//
// A persistent CTA (3 per SM) owns a group of G <= 8 consecutive envs at a time:
//   phase 1  warp w advances env e0+w in registers (env_step_warp: trading_env.py:54-100) and leaves w', the ring
//            pointer and the window row of the env in shared memory;
//   phase 2  the group's obs is one contiguous [G*A, W, 5] slab; it is streamed as 32-asset-row tiles (cut across
//            env boundaries).  A thread owns a fixed set of 8 float4 table loads + 8 ring loads per tile; all 16
//            are issued back to back, and the loads of tile i+1 are issued before tile i is handed to the TMA
//            store engine.  The tile in shared memory is the exact byte image of obs[e, a0:a0+32, :, :] and leaves
//            as ONE cp.async.bulk (UBLKCP) store with an L2 evict_first hint while the next tile is being filled.
// Weight channel semantics: ActionBuffer.get_all (weight_buffer.py:32-44) — zero front padding while the ring is not
// full, raw ring order once it is; the row written by this very step comes from shared memory, never from global.
//
// Template knobs (PMRL_TUNE_FAST_VARIANT selects them for A/B measurement; the defaults are the measured best):
//   WT     window size as a compile-time constant (50, 32) so every shared-memory offset of the hot loop is an
//          immediate; 0 = runtime W
//   RING2  (variant) predicated branch-free ring loads kept behind the table loads by a warp barrier: the L1 returns
//          loads in issue order, so DRAM-latency ring loads queued in front of L2-hit table loads make every load
//          of the tile wait for DRAM
// Measured: predicated branch-free ring loads WITHOUT the warp barrier let ptxas hoist them in front of
// the table loads (3.2 → 4.4 ms); with it they tie with the branchy form (3.27 vs 3.26 ms); 2 CTAs/SM at 123 registers
// without spills ties with 3 CTAs/SM at 80 (3.29 ms); issuing the proxy fence before the prefetch loads changes nothing;
// writing the tile out with plain 16-byte stores instead of the TMA bulk store is slower (5.0 ms).
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "pmrl_b200.h"
#include "pmrl_device.cuh"
#include "env_step.cuh"
#include "env_launch.h"
#include "host_util.h"

namespace pmrl {

constexpr int kFusedThreads = 256;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kMaxGroup = kFusedWarps;

struct GroupEnv { int row0, shift, fresh_slot, pad; };   // per env of the group (phase 1 → phase 2)

struct FeatRegs { float4 fv[4][2]; };     // one tile of in-flight table loads of a thread
struct RingRegs { float wv[8]; float fresh; int shift, wf; };   // one tile of in-flight ring loads of a thread

template <int NPL, bool HASC, int VEC, int WT, bool RING2, int MINB>
__global__ void __launch_bounds__(kFusedThreads, MINB) k_env_step_obs_fast(const StepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_stats[kFusedWarps * PMRL_STATS_LEN];
    __shared__ GroupEnv s_env[kMaxGroup];
    const int W = WT ? WT : p.W;
    const int A = p.A, T = p.T, G = p.group_envs;
    const int tile_floats = 32 * W * 5;
    float* const tile0 = reinterpret_cast<float*>(smem_raw);
    float* const tile1 = tile0 + tile_floats;
    float* const s_wnew = tile1 + tile_floats;           // [G, A]   w' of the group's envs, indexed by asset-row
    int* const s_ea = reinterpret_cast<int*>(s_wnew + G * A);   // [G, A]   (env-in-group << 16) | asset per asset-row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_groups = (p.E + G - 1) / G;
    const size_t row_floats = (size_t)W * 5;
    WarpStats ws;
    wstats_init(ws);
    const uint64_t pol_keep = kPolicyEvictLast, pol_once = kPolicyEvictFirst;   // immediates: no per-load register moves
    int buf = 0;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int e0 = grp * G;
        const int ne = min(G, p.E - e0);
        // ---------------- phase 1: one warp per env ----------------
        if (warp < ne) {
            const int e = e0 + warp;
            EnvVectors<NPL, HASC, VEC> ev;
            StepOut so;
            env_step_warp<NPL, HASC, VEC>(p, e, lane, ev, so, ws);
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const int a = asset_of<VEC>(lane, j);
                if (a < A) { s_wnew[warp * A + a] = ev.a[j]; s_ea[warp * A + a] = (warp << 16) | a; }
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
        const int ntiles = (R + 31) >> 5;
        float* const obs_grp = p.obs + (size_t)e0 * A * row_floats;
        {
            const float4* __restrict__ tbl = reinterpret_cast<const float4*>(p.feat_am);
            const float* __restrict__ hist_g = p.hist + (size_t)e0 * W * A;
            // thread-invariant shared-memory offsets of this thread's 8 feature rows / 8 weight columns
            const int fbase = (warp * W + lane) * 5, fstep = 8 * W * 5;
            const bool w0 = lane < W, w1 = lane + 32 < W;
            const int wbase = (lane * W + warp) * 5 + 4;
            const int nj = min(8, max(0, (W - warp + 7) >> 3));
            const int WA = W * A;
            auto load_feat = [&](FeatRegs& fr, int r0, auto partial) {
                constexpr bool PARTIAL = decltype(partial)::value;
                const int nr = PARTIAL ? R - r0 : 32;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!PARTIAL || warp + 8 * i < nr) {
                        const int ea = s_ea[r0 + warp + 8 * i];
                        const float4* __restrict__ src = tbl + ((ea & 0xffff) * T + s_env[ea >> 16].row0) + lane;
                        if (w0) fr.fv[i][0] = ld_keep4(src, pol_keep);
                        if (w1) fr.fv[i][1] = ld_keep4(src + 32, pol_keep);
                    }
                }
            };
            auto load_ring = [&](RingRegs& rr, int r0) {
                if (r0 + lane < R) {
                    const int ea = s_ea[r0 + lane];
                    const int el = ea >> 16;
                    const GroupEnv ge = s_env[el];
                    const float* __restrict__ base = hist_g + (el * WA + (ea & 0xffff));
                    const float fresh = s_wnew[r0 + lane];
                    if constexpr (RING2) {
                        // predicated loads (no branch per row); zero padding and the fresh row are resolved when the tile is
                        // written.  The caller separates them from the table loads with a warp barrier so that ptxas cannot
                        // hoist these DRAM-latency loads in front of the L2-hit table loads (the L1 returns in issue order).
                        rr.fresh = fresh; rr.shift = ge.shift; rr.wf = ge.fresh_slot + ge.shift;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int cs = max(warp + 8 * j - ge.shift, 0);
                            if (j < nj && cs != ge.fresh_slot) rr.wv[j] = ld_once(base + cs * A, pol_once);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {             // zero padding / fresh row resolved here, at load time
                            const int slot = warp + 8 * j - ge.shift;
                            float v = 0.0f;
                            if (j < nj && slot >= 0) v = (slot == ge.fresh_slot) ? fresh : ld_once(base + slot * A, pol_once);
                            rr.wv[j] = v;
                        }
                    }
                }
            };
            auto spill_tile = [&](const FeatRegs& fr, const RingRegs& rr, float* __restrict__ tile, int r0, auto partial) {
                constexpr bool PARTIAL = decltype(partial)::value;
                const int nr = PARTIAL ? R - r0 : 32;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!PARTIAL || warp + 8 * i < nr) {
                        float* d = tile + fbase + i * fstep;
                        if (w0) { const float4 v = fr.fv[i][0]; d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
                        if (w1) { const float4 v = fr.fv[i][1]; d[160] = v.x; d[161] = v.y; d[162] = v.z; d[163] = v.w; }
                    }
                }
                if (!PARTIAL || lane < nr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < nj) tile[wbase + 40 * j] = RING2 ? ((warp + 8 * j >= rr.shift) ? rr.wv[j] : 0.0f) : rr.wv[j];
                    if constexpr (RING2) { if ((rr.wf & 7) == warp) tile[(lane * W + rr.wf) * 5 + 4] = rr.fresh; }
                }
            };
            const int nfull = R >> 5;                      // tiles with all 32 asset-rows
            constexpr int RA = 1;
            auto emit = [&](FeatRegs& fr, RingRegs& rr, int ti) {   // registers → shared tile → TMA store; refill the registers
                // one CTA barrier per tile: thread 0 waits for the previous tile's store (other buffer) to finish reading shared
                // memory before it arrives at THIS tile's barrier, so past that barrier the other buffer is free (env_step_rt.cu)
                float* const tile = buf ? tile1 : tile0;
                const int r0 = ti * 32;
                if (ti < nfull) spill_tile(fr, rr, tile, r0, std::false_type{}); else spill_tile(fr, rr, tile, r0, std::true_type{});
                if (ti + 1 < ntiles) {                    // table loads of the next tile first (L2 hits) ...
                    if (ti + 1 < nfull) load_feat(fr, r0 + 32, std::false_type{}); else load_feat(fr, r0 + 32, std::true_type{});
                }
                if (RING2) __syncwarp();                  // keeps the ring loads behind the table loads in issue order
                if (ti + RA < ntiles) load_ring(rr, r0 + 32 * RA);   // ... then the DRAM-latency ring loads
                fence_proxy_async_smem();
                if (tid == 0) bulk_wait_read<0>();
                __syncthreads();
                const int nr = min(32, R - r0);
                float* const gdst = obs_grp + (size_t)r0 * row_floats;
                const int n = nr * W * 5;
                if (((((uintptr_t)gdst) | ((size_t)n * 4)) & 15) == 0) {
                    if (tid == 0) { bulk_store_s2g(gdst, tile, (uint32_t)n * 4u, pol_once); bulk_commit(); }
                } else {
                    for (int q = tid; q < n; q += kFusedThreads) gdst[q] = tile[q];
                }
                buf ^= 1;
            };
            FeatRegs fr;
            RingRegs rrA;
            if (nfull > 0) load_feat(fr, 0, std::false_type{}); else load_feat(fr, 0, std::true_type{});
            if (RING2) __syncwarp();
            load_ring(rrA, 0);
            for (int ti = 0; ti < ntiles; ++ti) emit(fr, rrA, ti);
        }
    }
    if (tid == 0) bulk_wait_read<0>();
    if (p.stats) { wstats_store(ws, s_stats + warp * PMRL_STATS_LEN, lane); stats_flush_block(p.stats, s_stats, kFusedWarps); }
}

}  // namespace pmrl

using namespace pmrl;

template <int NPL, bool HASC, int VEC, int WT>
static int launch_fast_t(StepParams& p, size_t smem, int grid, cudaStream_t s) {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_env_step_obs_fast<NPL, HASC, VEC, WT, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return pmrl_fail((int)e, "cudaFuncSetAttribute(k_env_step_obs_fast) failed");
        attr_done[dev] = true;
    }
    k_env_step_obs_fast<NPL, HASC, VEC, WT, false, 3><<<grid, kFusedThreads, smem, s>>>(p);
    return pmrl_check_launch("k_env_step_obs_fast");
}

template <int NPL, int VEC>
static int launch_fast_w(StepParams& p, size_t smem, int grid, cudaStream_t s) {
    const bool hasc = p.commission > 0.0f;
    if (p.W == 50) return hasc ? launch_fast_t<NPL, true, VEC, 50>(p, smem, grid, s) : launch_fast_t<NPL, false, VEC, 50>(p, smem, grid, s);
    return hasc ? launch_fast_t<NPL, true, VEC, 0>(p, smem, grid, s) : launch_fast_t<NPL, false, VEC, 0>(p, smem, grid, s);
}

int pmrl_launch_step_obs_fast(StepParams& p, int npl, int vec, int group, int ctas_per_sm, cudaStream_t s) {
    if (p.F != 5 || p.W > 64 || npl > 4) return -100;              // larger A: register budget of 3 CTAs/SM does not hold
    if ((size_t)p.A * p.T >= (1u << 31) || p.A >= 65536) return -100;
    const int slots = pmrl_sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 3);
    int G = group > 0 ? group : kMaxGroup;
    if (G > kMaxGroup) G = kMaxGroup;
    while (G > 1 && (p.E + G - 1) / G < 2 * slots) G >>= 1;         // small batches: more, smaller groups
    p.group_envs = G;
    p.tile_assets = 32;
    const size_t smem = (size_t)2 * 32 * p.W * 5 * 4 + (size_t)G * p.A * 8;
    if (smem > 200 * 1024) return -100;
    const int n_groups = (p.E + G - 1) / G;
    const int grid = n_groups < slots ? n_groups : slots;
#define FAST_CASE(N, V) if (npl == N && vec == V) return launch_fast_w<N, V>(p, smem, grid, s)
    FAST_CASE(1, 1); FAST_CASE(2, 1); FAST_CASE(2, 2); FAST_CASE(4, 1); FAST_CASE(4, 2); FAST_CASE(4, 4);
#undef FAST_CASE
    return -100;
}
