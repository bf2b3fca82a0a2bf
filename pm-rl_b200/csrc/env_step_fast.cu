// env_step_fast.cu — fused step + observation kernel, register-staged fill (the default Mode-O path for F == 5, W <= 64).
//
// A persistent CTA (3 per SM) owns a group of G <= 8 consecutive envs at a time:
//   phase 1  warp w advances env e0+w in registers (env_step_warp: trading_env.py:54-100) and leaves, per asset-row
//            of the group, w' and the table offset of its feature window in shared memory;
//   phase 2  the group's obs is one contiguous [G*A, W, 5] slab; it is streamed as 32-asset-row tiles (cut across
//            env boundaries).  A thread owns a fixed set of 8 float4 table loads + 8 ring loads per tile; all 16
//            are issued back to back, and the loads of tile i+1 are issued before tile i is handed to the TMA
//            store engine, so L2/DRAM latency overlaps the shared-memory interleave, the barriers and the store.
//            The tile in shared memory is the exact byte image of obs[e, a0:a0+32, :, :] and leaves as ONE
//            cp.async.bulk (UBLKCP) store with an L2 evict_first hint.
// The window size is a template parameter for the common W (50, 32) so that every shared-memory offset and
// loop bound of the hot loop is an immediate; WT == 0 is the runtime-W instance.
// Weight channel semantics: ActionBuffer.get_all (weight_buffer.py:32-44) — zero front padding while the ring is not
// full, raw ring order once it is; the row written by this very step comes from shared memory, never from global.
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "pmrl_b200.h"
#include "pmrl_device.cuh"
#include "env_step.cuh"
#include "env_launch.h"
#include "host_util.h"

namespace pmrl {

constexpr int kFastThreads = 256;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kFastGroup = kFastWarps;

struct FastEnv { int shift, wf; };        // per env of the group: ring shift and the window column of the fresh row

struct FastRegs {                         // one tile's worth of in-flight loads of a thread
    float4 fv[4][2];
    float wv[8];
    float fresh;
    int shift, wf;
};

__device__ __forceinline__ float4 ld_table4(const float4* p) {      // table rows: keep in L2 (constant evict_last policy)
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(kPolicyEvictLast));
    return v;
}

template <int NPL, bool HASC, int WT>
__global__ void __launch_bounds__(kFastThreads, 3) k_env_step_obs_fast(const StepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_stats[kFastWarps * PMRL_STATS_LEN];
    __shared__ FastEnv s_env[kFastGroup];
    const int W = WT ? WT : p.W;
    const int A = p.A, T = p.T, G = p.group_envs;
    const int tile_floats = 32 * W * 5;
    float* const tile0 = reinterpret_cast<float*>(smem_raw);
    float* const tile1 = tile0 + tile_floats;
    float* const s_wnew = tile1 + tile_floats;                    // [G*A]  w' per asset-row of the group
    int* const s_roff = reinterpret_cast<int*>(s_wnew + G * A);   // [G*A]  float4 offset of the asset-row's window in feat_am
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_groups = (p.E + G - 1) / G;
    const int WA = W * A;
    if (p.stats) stats_init_block(s_stats, kFastWarps);

    // thread-invariant shared-memory offsets: features rows warp+8i (i<4) × window rows lane, lane+32;
    // weights: asset-row = lane, window columns warp+8j (j<8)
    const int fbase = (warp * W + lane) * 5;
    const int wbase = (lane * W + warp) * 5 + 4;
    const bool w0 = lane < W, w1 = lane + 32 < W;
    const int nj = min(8, max(0, (W - warp + 7) >> 3));
    const float4* __restrict__ tbl = reinterpret_cast<const float4*>(p.feat_am) + lane;

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
            const int row0 = p.t0[e] + so.k;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const int a = lane + 32 * j;
                if (a < A) { s_wnew[warp * A + a] = ev.a[j]; s_roff[warp * A + a] = a * T + row0; }
            }
            if (lane == 0) {
                FastEnv fe;
                fe.shift = so.is_full ? 0 : (W - so.idx_new);                          // weight_buffer.py:38-42
                fe.wf = (so.did_reset ? 0 : so.slot_written) + fe.shift;               // column showing the row of this step
                s_env[warp] = fe;
            }
        }
        __syncthreads();
        // ---------------- phase 2: 32-asset-row tiles over the group's [ne*A] rows ----------------
        const int R = ne * A;
        const int ntiles = (R + 31) >> 5, nfull = R >> 5;
        const float* __restrict__ hist_g = p.hist + (size_t)e0 * WA;
        float* gdst = p.obs + (size_t)e0 * A * (W * 5);
        int wel = lane / A, wa = lane - wel * A;                   // (env-in-group, asset) of this lane's weight row

        auto load_tile = [&](FastRegs& tr, int r0, auto partial) {
            constexpr bool PARTIAL = decltype(partial)::value;
            const int nr = PARTIAL ? R - r0 : 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!PARTIAL || warp + 8 * i < nr) {
                    const float4* __restrict__ src = tbl + s_roff[r0 + warp + 8 * i];
                    if (w0) tr.fv[i][0] = ld_table4(src);
                    if (w1) tr.fv[i][1] = ld_table4(src + 32);
                }
            }
            if (!PARTIAL || lane < nr) {
                const FastEnv fe = s_env[wel];
                const float* __restrict__ base = hist_g + (wel * WA + wa);
                tr.fresh = s_wnew[r0 + lane];
                tr.shift = fe.shift;
                tr.wf = fe.wf;
#pragma unroll
                for (int j = 0; j < 8; ++j)                        // raw ring values; padding / fresh row are resolved at spill time
                    if (j < nj) tr.wv[j] = ld_once_c(base + max(warp + 8 * j - fe.shift, 0) * A);
            }
            wa += 32;
            while (wa >= A) { wa -= A; ++wel; }
        };
        auto spill_tile = [&](const FastRegs& tr, float* __restrict__ tile, int r0, auto partial) {
            constexpr bool PARTIAL = decltype(partial)::value;
            const int nr = PARTIAL ? R - r0 : 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!PARTIAL || warp + 8 * i < nr) {
                    float* d = tile + fbase + i * (8 * W * 5);
                    if (w0) { const float4 v = tr.fv[i][0]; d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
                    if (w1) { const float4 v = tr.fv[i][1]; d[160] = v.x; d[161] = v.y; d[162] = v.z; d[163] = v.w; }
                }
            }
            if (!PARTIAL || lane < nr) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < nj) tile[wbase + 40 * j] = (warp + 8 * j >= tr.shift) ? tr.wv[j] : 0.0f;
                if ((tr.wf & 7) == warp) tile[(lane * W + tr.wf) * 5 + 4] = tr.fresh;
            }
        };

        FastRegs tr;
        if (nfull > 0) load_tile(tr, 0, std::false_type{}); else load_tile(tr, 0, std::true_type{});
        for (int ti = 0; ti < ntiles; ++ti) {
            float* const tile = buf ? tile1 : tile0;
            if (tid == 0) bulk_wait_read<1>();                     // the store that last used this buffer has drained
            __syncthreads();
            const int r0 = ti * 32;
            if (ti < nfull) spill_tile(tr, tile, r0, std::false_type{}); else spill_tile(tr, tile, r0, std::true_type{});
            if (ti + 1 < ntiles) {                                 // next tile's loads fly during the barrier and the store
                if (ti + 1 < nfull) load_tile(tr, r0 + 32, std::false_type{}); else load_tile(tr, r0 + 32, std::true_type{});
            }
            fence_proxy_async_smem();
            __syncthreads();
            const int n = min(32, R - r0) * W * 5;
            if (((((uintptr_t)gdst) | ((size_t)n * 4)) & 15) == 0) {
                if (tid == 0) { bulk_store_s2g(gdst, tile, (uint32_t)n * 4u, kPolicyEvictFirst); bulk_commit(); }
            } else {
                for (int q = tid; q < n; q += kFastThreads) gdst[q] = tile[q];
            }
            gdst += 32 * W * 5;
            buf ^= 1;
        }
    }
    if (tid == 0) bulk_wait_read<0>();
    if (p.stats) stats_flush_block(p.stats, s_stats, kFastWarps);
}

}  // namespace pmrl

using namespace pmrl;

template <int NPL, bool HASC, int WT>
static int launch_fast_t(StepParams& p, size_t smem, int grid, cudaStream_t s) {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_env_step_obs_fast<NPL, HASC, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return pmrl_fail((int)e, "cudaFuncSetAttribute(k_env_step_obs_fast) failed");
        attr_done[dev] = true;
    }
    k_env_step_obs_fast<NPL, HASC, WT><<<grid, kFastThreads, smem, s>>>(p);
    return pmrl_check_launch("k_env_step_obs_fast");
}

template <int NPL, bool HASC>
static int launch_fast_w(StepParams& p, size_t smem, int grid, cudaStream_t s) {
    if (p.W == 50) return launch_fast_t<NPL, HASC, 50>(p, smem, grid, s);
    if (p.W == 32) return launch_fast_t<NPL, HASC, 32>(p, smem, grid, s);
    return launch_fast_t<NPL, HASC, 0>(p, smem, grid, s);
}

int pmrl_launch_step_obs_fast(StepParams& p, int npl, int group, int ctas_per_sm, cudaStream_t s) {
    if (p.F != 5 || p.W > 64 || npl > 4) return -100;              // larger A: register budget of 3 CTAs/SM does not hold
    if ((size_t)p.A * p.T >= (1u << 31)) return -100;
    const int slots = pmrl_sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 3);
    int G = group > 0 ? group : kFastGroup;
    if (G > kFastGroup) G = kFastGroup;
    while (G > 1 && (p.E + G - 1) / G < 2 * slots) G >>= 1;         // small batches: more, smaller groups
    p.group_envs = G;
    p.tile_assets = 32;
    const size_t smem = (size_t)2 * 32 * p.W * 5 * 4 + (size_t)G * p.A * 8;
    if (smem > 200 * 1024) return -100;
    const int n_groups = (p.E + G - 1) / G;
    const int grid = n_groups < slots ? n_groups : slots;
    const bool hasc = p.commission > 0.0f;
#define FAST_CASE(N) return hasc ? launch_fast_w<N, true>(p, smem, grid, s) : launch_fast_w<N, false>(p, smem, grid, s)
    switch (npl) {
        case 1: FAST_CASE(1);
        case 2: FAST_CASE(2);
        default: FAST_CASE(4);
    }
#undef FAST_CASE
}
