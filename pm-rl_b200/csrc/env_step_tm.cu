// env_step_tm.cu — fused step + observation kernel, variant "TM": every byte enters and leaves the SM through TMA.
//
//   feature windows  one cp.async.bulk.tensor.3d per tile: box {4 floats, W rows, TR assets} of the asset-major table
//                    viewed as a rank-3 tensor {4, T, A} lands in shared memory as the dense [TR, W, 4] staging tile
//                    (two stages, issued two tiles ahead by one thread, completion on an mbarrier, L2 evict_last);
//   weight ring      one 1-D bulk load of the env's whole ring (W·A·4 B), double-buffered over consecutive envs
//                    (as in env_step_rt.cu);
//   observation      the [TR, W, 5] tile (byte image of obs[e, a0:a0+TR]) leaves as one 1-D bulk store.
// The SM only does the shared→shared interleave (16-byte loads, stride-5 conflict-free stores) and the stepping
// (one warp per env, env_step.cuh).  No global load goes through registers or the L1 load queue.
// Tiles are cut per env (TR | A, 16 <= TR <= 32, TR·W·5 % 4 == 0) so that one TMA box never crosses an env boundary.
// Weight channel semantics: ActionBuffer.get_all (weight_buffer.py:32-44); window rows: data/instrument.py:351-356.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <mutex>
#include "pmrl_b200.h"
#include "pmrl_device.cuh"
#include "env_step.cuh"
#include "env_launch.h"
#include "host_util.h"

namespace pmrl {

constexpr int kTmConsWarps = 8;
constexpr int kTmConsThreads = kTmConsWarps * 32;          // 256 streaming threads
constexpr int kTmStepWarps = 4;                            // + 4 warps that advance the NEXT group meanwhile
constexpr int kTmThreads = kTmConsThreads + kTmStepWarps * 32;   // 384
constexpr int kTmGroup = 8;

struct TmEnv { int row0, shift, fresh_slot, pad; };

__device__ __forceinline__ void tma_load_3d(void* sdst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 :: "r"(smem_u32(sdst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

template <int NPL, bool HASC>
__global__ void __launch_bounds__(kTmThreads, 2) k_env_step_obs_tm(const StepParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_stats[kTmStepWarps * PMRL_STATS_LEN];
    __shared__ TmEnv s_env[2][kTmGroup];
    __shared__ __align__(8) uint64_t s_rbar[2], s_fbar[2], s_gfull[2], s_gempty[2];
    const int A = p.A, W = p.W, G = p.group_envs, TR = p.tile_assets;
    const int WA = W * A;
    const int TPE = A / TR;                                          // tiles per env
    const int tile_floats = TR * W * 5, stage_floats = TR * W * 4;
    float* const stage0 = reinterpret_cast<float*>(smem_raw);        // [2][TR*W*4]  TMA destinations first: 128-byte aligned
    float* const tile0 = stage0 + 2 * stage_floats;
    float* const tile1 = tile0 + tile_floats;
    float* const s_ring = tile1 + tile_floats;                       // [2][W*A]
    float* const s_wnew = s_ring + 2 * WA;                           // [2][G*A]  w' of the group being streamed / being stepped
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_groups = (p.E + G - 1) / G;
    const int n_it = ((int)blockIdx.x < n_groups) ? (n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_rbar[b], 1); mbar_init(&s_fbar[b], 1);
            mbar_init(&s_gfull[b], kTmStepWarps); mbar_init(&s_gempty[b], kTmConsWarps);
        }
        mbar_fence_init();
    }
    for (int q = tid; q < kTmStepWarps * PMRL_STATS_LEN; q += kTmThreads)
        s_stats[q] = (q % PMRL_STATS_LEN >= PMRL_STAT_MAX_V) ? -INFINITY : 0.0;
    __syncthreads();

    if (warp >= kTmConsWarps) {
        // ================================ steppers: group it+1 while group it streams ================================
        const int sw = warp - kTmConsWarps;
        for (int it = 0; it < n_it; ++it) {
            const int grp = blockIdx.x + it * gridDim.x;
            const int e0 = grp * G, ne = min(G, p.E - e0), b = it & 1;
            mbar_wait(&s_gempty[b], ((it >> 1) & 1) ^ 1);               // the block of group it-2 has been consumed
            float* const wnew_b = s_wnew + b * G * A;
            for (int el = sw; el < ne; el += kTmStepWarps) {
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
                    TmEnv ge;
                    ge.row0 = p.t0[e] + so.k;
                    ge.shift = so.is_full ? 0 : (W - so.idx_new);        // weight_buffer.py:38-42
                    ge.fresh_slot = so.did_reset ? 0 : so.slot_written;   // the row written by this launch comes from smem
                    ge.pad = 0;
                    s_env[b][el] = ge;
                }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_gfull[b]);
        }
        if (p.stats) {                                                   // 4 stepper warps → 10 atomics per CTA
            named_bar_sync(2, kTmStepWarps * 32);
            const int q = tid - kTmConsThreads;
            if (q < PMRL_STATS_LEN) {
                double v = s_stats[q];
                for (int wi = 1; wi < kTmStepWarps; ++wi) {
                    const double o = s_stats[wi * PMRL_STATS_LEN + q];
                    v = (q >= PMRL_STAT_MAX_V) ? fmax(v, o) : v + o;
                }
                if (q >= PMRL_STAT_MAX_V) { if (v > -INFINITY) atomic_max_double(p.stats + q, v); }
                else if (v != 0.0) atomicAdd(p.stats + q, v);
            }
        }
        return;
    }

    // ================================ consumers (256 threads) ================================
    // thread-invariant offsets: features rows warp+8i (i<4) × window rows lane, lane+32; weights: asset-row = lane, columns warp+8j
    const int fbase = (warp * W + lane) * 5, fstep = 8 * W * 5;
    const int sbase = warp * W + lane, sstep = 8 * W;                // float4 index into the staging tile
    const bool w0 = lane < W, w1 = lane + 32 < W;
    const int wbase = (lane * W + warp) * 5 + 4;
    const int nj = min(8, max(0, (W - warp + 7) >> 3));
    const uint32_t stage_bytes = (uint32_t)stage_floats * 4u;

    // issue cursors of thread 0: the TMA loads form ONE stream across group boundaries (two tiles / two envs ahead)
    int f_it = 0, f_tile = 0, f_n = 0;                               // next feature tile to issue: group iter, tile in group, running index
    int r_it = 0, r_env = 0, r_n = 0;                                // next ring to issue: group iter, env in group, running index
    auto issue_feats = [&](int upto_n) {                             // thread 0 only
        while (f_n < upto_n && f_it < n_it) {
            const int grp = blockIdx.x + f_it * gridDim.x;
            const int ne = min(G, p.E - grp * G);
            if (f_tile == 0) mbar_wait(&s_gfull[f_it & 1], (uint32_t)((f_it >> 1) & 1));   // steppers are (normally) a group ahead
            const int el = f_tile / TPE, k = f_tile - el * TPE;
            mbar_arrive_expect_tx(&s_fbar[f_n & 1], stage_bytes);
            tma_load_3d(stage0 + (f_n & 1) * stage_floats, &tmap, 0, s_env[f_it & 1][el].row0, k * TR, &s_fbar[f_n & 1], kPolicyEvictLast);
            ++f_n;
            if (++f_tile == ne * TPE) { f_tile = 0; ++f_it; }
        }
    };
    auto issue_rings = [&](int upto_n) {                             // thread 0 only
        while (r_n < upto_n && r_it < n_it) {
            const int grp = blockIdx.x + r_it * gridDim.x;
            const int e0 = grp * G, ne = min(G, p.E - e0);
            mbar_arrive_expect_tx(&s_rbar[r_n & 1], (uint32_t)WA * 4u);
            bulk_load_g2s(s_ring + (r_n & 1) * WA, p.hist + (size_t)(e0 + r_env) * WA, (uint32_t)WA * 4u, &s_rbar[r_n & 1], kPolicyEvictFirst);
            ++r_n;
            if (++r_env == ne) { r_env = 0; ++r_it; }
        }
    };
    if (tid == 0) { issue_rings(2); issue_feats(2); }

    int buf = 0;
    int c_tile = 0, c_env = 0;                                       // running indices of the tile / env being consumed
    for (int it = 0; it < n_it; ++it) {
        const int grp = blockIdx.x + it * gridDim.x;
        const int e0 = grp * G, ne = min(G, p.E - e0), b = it & 1;
        mbar_wait(&s_gfull[b], (uint32_t)((it >> 1) & 1));
        const float* const wnew_b = s_wnew + b * G * A;
        const TmEnv* const env_b = s_env[b];
        float* gdst = p.obs + (size_t)e0 * A * (W * 5);
        for (int el = 0; el < ne; ++el, ++c_env) {
            const TmEnv ge = env_b[el];
            for (int k = 0; k < TPE; ++k, ++c_tile) {
                const int a0 = k * TR;
                float* const tile = buf ? tile1 : tile0;
                if (tid == 0) bulk_wait_read<1>();                    // the store that last used this tile buffer has drained
                named_bar_sync(1, kTmConsThreads);
                const float4* const st = reinterpret_cast<const float4*>(stage0 + (c_tile & 1) * stage_floats);
                // weight channel from the staged ring (lane = asset-row of the tile)
                if (lane < TR) {
                    mbar_wait(&s_rbar[c_env & 1], (uint32_t)((c_env >> 1) & 1));
                    const float* __restrict__ rs = s_ring + (c_env & 1) * WA + a0 + lane;
                    const float fresh = wnew_b[el * A + a0 + lane];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < nj) {
                            const int slot = warp + 8 * j - ge.shift;
                            float v = 0.0f;                           // zero front padding while the ring is not full
                            if (slot >= 0) v = (slot == ge.fresh_slot) ? fresh : rs[slot * A];
                            tile[wbase + 40 * j] = v;
                        }
                    }
                }
                // feature channels from the TMA-staged [TR, W, 4] tile
                mbar_wait(&s_fbar[c_tile & 1], (uint32_t)((c_tile >> 1) & 1));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (warp + 8 * i < TR) {
                        float* d = tile + fbase + i * fstep;
                        if (w0) { const float4 v = st[sbase + i * sstep]; d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
                        if (w1) { const float4 v = st[sbase + i * sstep + 32]; d[160] = v.x; d[161] = v.y; d[162] = v.z; d[163] = v.w; }
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(1, kTmConsThreads);                    // tile complete; its stage (and maybe a ring buffer) released
                if (tid == 0) {
                    bulk_store_s2g(gdst, tile, (uint32_t)tile_floats * 4u, kPolicyEvictFirst);
                    bulk_commit();
                    issue_feats(c_tile + 3);                          // refill the stage just consumed (tile index c_tile + 2)
                    issue_rings(c_env + (k == TPE - 1 ? 1 : 0) + 2);  // an env's ring buffer is free after its last tile
                }
                gdst += tile_floats;
                buf ^= 1;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_gempty[b]);                     // steppers may reuse this group block
    }
    if (tid == 0) bulk_wait_read<0>();
}

}  // namespace pmrl

using namespace pmrl;

// ---- host: tensor-map cache (encoding costs microseconds; the table pointer / shape rarely change) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapKey { const void* ptr; int A, T, W, TR, dev; };
static TmapKey g_keys[8];
static CUtensorMap g_maps[8];
static int g_nmaps = 0, g_next = 0;
static std::mutex g_tmap_mutex;

static int get_tmap(const StepParams& p, int TR, CUtensorMap* out) {
    int dev = 0;
    cudaGetDevice(&dev);
    TmapKey key{p.feat_am, p.A, p.T, p.W, TR, dev};
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    for (int i = 0; i < g_nmaps; ++i)
        if (memcmp(&g_keys[i], &key, sizeof(key)) == 0) { *out = g_maps[i]; return 0; }
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -100;
        encode = (EncodeTiledFn)fn;
    }
    // feat_am [A, T, 4] fp32 as a rank-3 tensor {4, T, A} (innermost first)
    const cuuint64_t gdim[3] = {4, (cuuint64_t)p.T, (cuuint64_t)p.A};
    const cuuint64_t gstride[2] = {16, (cuuint64_t)p.T * 16};         // bytes, dims 1..2
    const cuuint32_t box[3] = {4, (cuuint32_t)p.W, (cuuint32_t)TR};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.feat_am), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -100;
    const int slot = g_nmaps < 8 ? g_nmaps++ : (g_next++ & 7);
    g_keys[slot] = key; g_maps[slot] = m;
    *out = m;
    return 0;
}

template <int NPL, bool HASC>
static int launch_tm_t(StepParams& p, const CUtensorMap& tmap, size_t smem, int grid, cudaStream_t s) {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_env_step_obs_tm<NPL, HASC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return pmrl_fail((int)e, "cudaFuncSetAttribute(k_env_step_obs_tm) failed");
        attr_done[dev] = true;
    }
    k_env_step_obs_tm<NPL, HASC><<<grid, kTmThreads, smem, s>>>(p, tmap);
    return pmrl_check_launch("k_env_step_obs_tm");
}

int pmrl_launch_step_obs_tm(StepParams& p, int npl, int group, int ctas_per_sm, cudaStream_t s) {
    if (p.F != 5 || p.W > 64 || p.W < 2 || npl > 4 || p.A < 16) return -100;
    if (((size_t)p.W * p.A) % 4 != 0 || ((uintptr_t)p.hist) % 16 != 0 || ((uintptr_t)p.feat_am) % 16 != 0 || ((uintptr_t)p.obs) % 16 != 0) return -100;
    // tile height: the largest TR in [16, 32] that divides A and keeps every tile 16-byte aligned in obs
    int TR = 0;
    for (int t = 32; t >= 16; --t)
        if (p.A % t == 0 && ((size_t)t * p.W * 5) % 4 == 0) { TR = t; break; }
    if (!TR) return -100;
    const int per_sm = ctas_per_sm > 0 ? ctas_per_sm : 2;
    const int slots = pmrl_sm_count() * per_sm;
    int G = group > 0 ? group : kTmGroup;
    if (G > kTmGroup) G = kTmGroup;
    while (G > 1 && (p.E + G - 1) / G < 2 * slots) G >>= 1;
    p.group_envs = G;
    p.tile_assets = TR;
    const size_t smem = (size_t)2 * TR * p.W * 5 * 4 + (size_t)2 * TR * p.W * 4 * 4 + (size_t)2 * p.W * p.A * 4 + (size_t)2 * G * p.A * 4;
    if (smem > (size_t)(226 * 1024) / per_sm - 1024) return -100;
    if (((size_t)TR * p.W * 16) % 128 != 0) return -100;              // both TMA staging tiles must be 128-byte aligned
    CUtensorMap tmap;
    if (get_tmap(p, TR, &tmap) != 0) return -100;
    const int n_groups = (p.E + G - 1) / G;
    const int grid = n_groups < slots ? n_groups : slots;
    const bool hasc = p.commission > 0.0f;
#define TM_CASE(N) return hasc ? launch_tm_t<N, true>(p, tmap, smem, grid, s) : launch_tm_t<N, false>(p, tmap, smem, grid, s)
    switch (npl) {
        case 1: TM_CASE(1);
        case 2: TM_CASE(2);
        default: TM_CASE(4);
    }
#undef TM_CASE
}
