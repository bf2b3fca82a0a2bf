// env_step_tma.cu — fused step + observation kernel as a warp-specialised TMA pipeline (sm_100a).
//
// One persistent CTA per SM, 13 warps in three roles, all hand-offs through mbarriers:
//
//   steppers  (4 warps)  advance the envs of group g+1 (env_step_warp: the whole reference transition in
//                        registers) while group g is being streamed, and publish w', the ring pointer
//                        and the window row of every env of the group in a double-buffered shared block;
//   producer  (1 warp)   issues the feature-window loads of the next tiles as TMA bulk copies
//                        (cp.async.bulk global→shared, one 16·W-byte run per asset-row, L2 evict_last)
//                        into an NS-deep ring of staging buffers, completion counted on full[s];
//   consumers (8 warps)  turn a staged tile into the exact [32, W, 5] byte image of the reference obs
//                        layout (16-byte shared loads, stride-5 conflict-free shared stores, weight channel
//                        from register-prefetched ring rows, the row written by this step from shared
//                        memory) and hand it to the TMA store engine as ONE bulk store (L2 evict_first).
//
// Global-memory latency is therefore never on the consumers' critical path: NS tiles of feature loads are
// always in flight per SM, and the stepping latency of a group overlaps the streaming of the previous one.
// Reference semantics: trading_env.py:44-105, weight_buffer.py:13-44 (see env_step.cuh / obs_tile.cuh).
#include <cuda_runtime.h>
#include <stdint.h>
#include "pmrl_b200.h"
#include "pmrl_device.cuh"
#include "env_step.cuh"
#include "env_launch.h"
#include "host_util.h"

namespace pmrl {

constexpr int kConsWarps = 8;
constexpr int kConsThreads = kConsWarps * 32;
constexpr int kStepperWarps = 4;
constexpr int kProducerWarp = kConsWarps;
constexpr int kFirstStepper = kConsWarps + 1;
constexpr int kTmaWarps = kConsWarps + 1 + kStepperWarps;
constexpr int kTmaThreads = kTmaWarps * 32;              // 416
constexpr int kMaxStages = 6;
constexpr int kMaxGroupT = 8;

struct GroupEnvT { int row0, shift, fresh_slot, pad; };

struct Cursor {           // (env-in-group, asset) of a running asset-row index, advanced without divisions
    int el, a;
    __device__ __forceinline__ void init(int gar, int A) { el = gar / A; a = gar - el * A; }
    __device__ __forceinline__ void advance(int d, int A) { a += d; while (a >= A) { a -= A; ++el; } }
};

template <int NPL, bool HASC>
__global__ void __launch_bounds__(kTmaThreads, 1) k_env_step_obs_tma(const StepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kMaxStages], s_empty[kMaxStages], s_gfull[2], s_gempty[2];
    __shared__ GroupEnvT s_env[2][kMaxGroupT];
    __shared__ double s_stats[kStepperWarps * PMRL_STATS_LEN];

    const int A = p.A, W = p.W, T = p.T, G = p.group_envs, NS = p.tma_stages;
    const int tile_floats = 32 * W * 5;                  // one obs tile: 32 asset-rows × W × 5
    const int stage_floats = 32 * W * 4;                 // one staged feature tile: 32 asset-rows × W × 4
    float* const tile0 = reinterpret_cast<float*>(smem_raw);
    float* const tile1 = tile0 + tile_floats;
    float* const stage0 = tile1 + tile_floats;           // NS staging buffers
    float* const s_wnew = stage0 + NS * stage_floats;    // [2][G*A]  w' per asset-row of the group
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_groups = (p.E + G - 1) / G;
    const int n_it = ((int)blockIdx.x < n_groups) ? (n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const size_t row_floats = (size_t)W * 5;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kConsWarps); }
        for (int b = 0; b < 2; ++b) { mbar_init(&s_gfull[b], kStepperWarps); mbar_init(&s_gempty[b], kConsWarps + 1); }
        mbar_fence_init();
    }
    for (int q = tid; q < kStepperWarps * PMRL_STATS_LEN; q += kTmaThreads)
        s_stats[q] = (q % PMRL_STATS_LEN >= PMRL_STAT_MAX_V) ? -INFINITY : 0.0;
    __syncthreads();

    if (warp >= kFirstStepper) {
        // ================================ steppers ================================
        const int sw = warp - kFirstStepper;
        for (int it = 0; it < n_it; ++it) {
            const int grp = blockIdx.x + it * gridDim.x;
            const int e0 = grp * G, ne = min(G, p.E - e0), b = it & 1;
            mbar_wait(&s_gempty[b], ((it >> 1) & 1) ^ 1);             // the block of group it-2 has been consumed
            float* const wnew_b = s_wnew + b * G * A;
            for (int el = sw; el < ne; el += kStepperWarps) {
                const int e = e0 + el;
                EnvVectors<NPL, HASC> ev;
                StepOut so;
                env_step_warp<NPL, HASC>(p, e, lane, ev, so, s_stats + sw * PMRL_STATS_LEN);
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    const int a = lane + 32 * j;
                    if (a < A) wnew_b[el * A + a] = ev.a[j];
                }
                if (lane == 0) {
                    GroupEnvT ge;
                    ge.row0 = p.t0[e] + so.k;
                    ge.shift = so.is_full ? 0 : (W - so.idx_new);      // weight_buffer.py:38-42
                    ge.fresh_slot = so.did_reset ? 0 : so.slot_written; // the row written by this launch comes from smem
                    ge.pad = 0;
                    s_env[b][el] = ge;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_gfull[b]);
        }
        if (p.stats) {                                                 // 4 stepper warps → 10 atomics per CTA
            named_bar_sync(2, kStepperWarps * 32);
            const int q = tid - kFirstStepper * 32;
            if (q < PMRL_STATS_LEN) {
                double v = s_stats[q];
                for (int wi = 1; wi < kStepperWarps; ++wi) {
                    const double o = s_stats[wi * PMRL_STATS_LEN + q];
                    v = (q >= PMRL_STAT_MAX_V) ? fmax(v, o) : v + o;
                }
                if (q >= PMRL_STAT_MAX_V) { if (v > -INFINITY) atomic_max_double(p.stats + q, v); }
                else if (v != 0.0) atomicAdd(p.stats + q, v);
            }
        }
    } else if (warp == kProducerWarp) {
        // ================================ producer ================================
        const float4* __restrict__ tbl = reinterpret_cast<const float4*>(p.feat_am);
        int s = 0, ph = 0;
        for (int it = 0; it < n_it; ++it) {
            const int grp = blockIdx.x + it * gridDim.x;
            const int e0 = grp * G, ne = min(G, p.E - e0), b = it & 1;
            mbar_wait(&s_gfull[b], (it >> 1) & 1);
            const int R = ne * A, ntiles = (R + 31) >> 5;
            const int row0_l = s_env[b][lane & (kMaxGroupT - 1)].row0;  // lane l < 8 holds the window row of env l
            Cursor cur;
            cur.init(lane, A);                                          // this lane's asset-row of tile 0
            for (int ti = 0; ti < ntiles; ++ti) {
                mbar_wait(&s_empty[s], ph ^ 1);                         // consumers are done with this staging buffer
                const int nr = min(32, R - ti * 32);
                if (lane == 0) mbar_arrive_expect_tx(&s_full[s], (uint32_t)(nr * W * 16));
                __syncwarp();
                const int row0 = __shfl_sync(PMRL_FULL_MASK, row0_l, cur.el & (kMaxGroupT - 1));
                if (lane < nr)
                    bulk_load_g2s(stage0 + (size_t)s * stage_floats + lane * W * 4, tbl + ((size_t)cur.a * T + row0),
                                  (uint32_t)(W * 16), &s_full[s], kPolicyEvictLast);
                cur.advance(32, A);
                if (++s == NS) { s = 0; ph ^= 1; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_gempty[b]);
        }
    } else {
        // ================================ consumers ================================
        // thread-invariant shared-memory offsets: features rows warp+8i (i<4) × window rows lane, lane+32;
        // weights: asset-row = lane, window columns warp+8j (j<8)
        const int fbase = (warp * W + lane) * 5, fstep = 8 * W * 5;
        const int sbase = warp * W + lane, sstep = 8 * W;               // float4 index into the staging buffer
        const bool w0 = lane < W, w1 = lane + 32 < W;
        const int wbase = (lane * W + warp) * 5 + 4;
        const int nj = min(8, max(0, (W - warp + 7) >> 3));
        const int WA = W * A;
        int s = 0, ph = 0, buf = 0;
        for (int it = 0; it < n_it; ++it) {
            const int grp = blockIdx.x + it * gridDim.x;
            const int e0 = grp * G, ne = min(G, p.E - e0), b = it & 1;
            mbar_wait(&s_gfull[b], (it >> 1) & 1);
            const float* const wnew_b = s_wnew + b * G * A;
            const GroupEnvT* const env_b = s_env[b];
            const int R = ne * A, ntiles = (R + 31) >> 5;
            const float* __restrict__ hist_g = p.hist + (size_t)e0 * WA;
            float* const obs_grp = p.obs + (size_t)e0 * A * row_floats;
            Cursor cw;
            cw.init(lane, A);
            float wv[8], fresh = 0.0f;
            int wf = -1;                                                // window column that shows the fresh ring row
            int shift_r = 0;                                            // ring shift of the env whose rows sit in wv[]
            auto load_ring = [&](int r0) {                              // ring rows of tile r0 → registers; nothing here
                if (r0 + lane < R) {                                    // consumes the loaded values (they are used one
                    const GroupEnvT ge = env_b[cw.el];                  // tile later), so the DRAM latency stays hidden
                    const float* __restrict__ base = hist_g + (cw.el * WA + cw.a);
                    fresh = wnew_b[r0 + lane];
                    wf = ge.fresh_slot + ge.shift;
                    shift_r = ge.shift;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                       // never load the row this launch has just written
                        const int cs = max(warp + 8 * j - ge.shift, 0);   // (patched from smem; a load hitting the in-flight
                        if (j < nj && cs != ge.fresh_slot) wv[j] = ld_once_c(base + cs * A);   // store stalls the L1 queue)
                    }
                }
                cw.advance(32, A);
            };
            load_ring(0);
            for (int ti = 0; ti < ntiles; ++ti) {
                float* const tile = buf ? tile1 : tile0;
                if (tid == 0) bulk_wait_read<1>();                      // the store that last used this buffer has drained
                named_bar_sync(1, kConsThreads);
                const int r0 = ti * 32;
                const int nr = min(32, R - r0);
                // weight channel from the prefetched registers; the row written by this step from shared memory
                if (lane < nr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)                         // zero front padding while the ring is not full
                        if (j < nj) tile[wbase + 40 * j] = (warp + 8 * j >= shift_r) ? wv[j] : 0.0f;
                    if ((wf & 7) == warp) tile[(lane * W + wf) * 5 + 4] = fresh;
                }
                if (ti + 1 < ntiles) load_ring(r0 + 32);                // flies during the feature copy, the barrier and the store
                // feature channels: staged by TMA → interleave into the 5-float rows
                mbar_wait(&s_full[s], ph);
                const float4* const st = reinterpret_cast<const float4*>(stage0 + (size_t)s * stage_floats);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (warp + 8 * i < nr) {
                        float* d = tile + fbase + i * fstep;
                        if (w0) { const float4 v = st[sbase + i * sstep]; d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
                        if (w1) { const float4 v = st[sbase + i * sstep + 32]; d[160] = v.x; d[161] = v.y; d[162] = v.z; d[163] = v.w; }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[s]);                // staging buffer may be refilled
                if (++s == NS) { s = 0; ph ^= 1; }
                fence_proxy_async_smem();
                named_bar_sync(1, kConsThreads);
                float* const gdst = obs_grp + (size_t)r0 * row_floats;
                const int n = nr * W * 5;
                if (((((uintptr_t)gdst) | ((size_t)n * 4)) & 15) == 0) {
                    if (tid == 0) { bulk_store_s2g(gdst, tile, (uint32_t)n * 4u, kPolicyEvictFirst); bulk_commit(); }
                } else {
                    for (int q = tid; q < n; q += kConsThreads) gdst[q] = tile[q];
                }
                buf ^= 1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_gempty[b]);
        }
        if (tid == 0) bulk_wait_read<0>();
    }
}

}  // namespace pmrl

using namespace pmrl;

template <int NPL, bool HASC>
static int launch_tma_t(StepParams& p, size_t smem, int grid, cudaStream_t s) {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_env_step_obs_tma<NPL, HASC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return pmrl_fail((int)e, "cudaFuncSetAttribute(k_env_step_obs_tma) failed");
        attr_done[dev] = true;
    }
    k_env_step_obs_tma<NPL, HASC><<<grid, kTmaThreads, smem, s>>>(p);
    return pmrl_check_launch("k_env_step_obs_tma");
}

int pmrl_launch_step_obs_tma(StepParams& p, int npl, int stages, int group, cudaStream_t s) {
    if (p.F != 5 || p.W > 64 || npl > 16) return -100;
    if ((size_t)p.A * p.T >= (1u << 31)) return -100;
    const int sms = pmrl_sm_count();
    int G = group > 0 ? group : kMaxGroupT;
    if (G > kMaxGroupT) G = kMaxGroupT;
    while (G > 1 && (p.E + G - 1) / G < 4 * sms) G >>= 1;               // small batches: more, smaller groups
    const size_t tile_b = (size_t)32 * p.W * 5 * 4, stage_b = (size_t)32 * p.W * 4 * 4;
    const size_t fixed = 2 * tile_b + (size_t)2 * G * p.A * 4;
    const size_t budget = 224 * 1024;
    int NS = stages > 0 ? stages : 4;
    if (NS > kMaxStages) NS = kMaxStages;
    while (NS > 2 && fixed + NS * stage_b > budget) --NS;
    if (fixed + NS * stage_b > budget) return -100;
    p.group_envs = G;
    p.tma_stages = NS;
    p.tile_assets = 32;
    const size_t smem = fixed + NS * stage_b;
    const int n_groups = (p.E + G - 1) / G;
    const int grid = n_groups < sms ? n_groups : sms;
    const bool hasc = p.commission > 0.0f;
#define TMA_CASE(N) return hasc ? launch_tma_t<N, true>(p, smem, grid, s) : launch_tma_t<N, false>(p, smem, grid, s)
    switch (npl) {
        case 1: TMA_CASE(1);
        case 2: TMA_CASE(2);
        case 4: TMA_CASE(4);
        case 8: TMA_CASE(8);
        default: TMA_CASE(16);
    }
#undef TMA_CASE
}
