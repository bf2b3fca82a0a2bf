// ffd.cu — feature path: fixed-width-window fractional differencing, per-series scaling, table packing.
//
//   k_ffd_weights    FixedFracDiff._objective weights/width      data/ffd.py:38-47
//   k_ffd_conv       FixedFracDiff.transform (valid conv, tail-aligned)  data/ffd.py:80-89
//   k_scale_series   Instrument.scale (sklearn MinMaxScaler / StandardScaler)  data/instrument.py:318-336
//   k_pack_*         series-major planes → [A,T,C] asset-major feature table + [T,A] close plane
//                    (the window layout of Instrument.window, data/instrument.py:339-356)
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "pmrl_b200.h"
#include "host_util.h"

namespace pmrl {

// One block per series.  torch.cumprod on the CPU multiplies left to right with a *double* accumulator (ATen
// acc_type<float>) and rounds every prefix to fp32, so the chain is reproduced bit for bit the same way: the
// factors 1 - (d+1)/k are independent (all threads, fp32 like torch.div(...)+1), only the running product is a
// sequential scan (one thread, one DMUL per element, factors staged through shared memory).
constexpr int kWThreads = 128;
constexpr int kWChunk = 2048;

__global__ void __launch_bounds__(kWThreads) k_ffd_weights(const double* __restrict__ d, int T, float thres,
                                                           float* __restrict__ weights, int32_t* __restrict__ widths) {
    __shared__ float s_proto[kWChunk];
    const int n = blockIdx.x, tid = threadIdx.x;
    const float factor = (float)(-(d[n] + 1.0));           // zeros(len-1) - (d+1)        (ffd.py:38)
    float* __restrict__ w = weights + (size_t)n * T;
    double acc = 1.0;
    int width = 0;
    for (int k0 = 0; k0 < T; k0 += kWChunk) {
        const int kc = min(kWChunk, T - k0);
        __syncthreads();
        for (int q = tid; q < kc; q += kWThreads) {
            const int k = k0 + q;
            s_proto[q] = k == 0 ? 1.0f : __fadd_rn(__fdiv_rn(factor, (float)k), 1.0f);    // div(factor, k) + 1  (ffd.py:39)
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll 8
            for (int q = 0; q < kc; ++q) {
                acc *= (double)s_proto[q];                   // cumprod                      (ffd.py:40)
                const float wk = (float)acc;
                s_proto[q] = wk;
                if (fabsf(wk) > thres) width = k0 + q;       // where(|w|>thres).max()       (ffd.py:43)
            }
        }
        __syncthreads();
        for (int q = tid; q < kc; q += kWThreads) w[k0 + q] = s_proto[q];
    }
    if (tid == 0) widths[n] = width;
}

// Valid convolution with taps w[0:width] (tap `width` itself is dropped, ffd.py:47), tail-aligned:
//   out[n, j] = Σ_{k<width} w[n,k] · x[n, max_width + j − k]
// Block = (series n, tile of 1024 outputs), 128 threads × 8 CONSECUTIVE outputs.  Taps are consumed in chunks of 256
// staged in shared memory with the matching x segment.  Per group of 8 taps a thread keeps a 16-float register window
// of x (two 16-byte shared loads refill it, the rest is reused from the previous group) and reads the 8 taps with two
// broadcast 16-byte loads: 64 FMAs per 4 shared loads (the first version did 4 FMAs per 5 loads and was LDS-bound).
constexpr int kConvThreads = 128;
constexpr int kPer = 8;
constexpr int kTile = kConvThreads * kPer;     // 1024 outputs per block
constexpr int kChunk = 256;                    // taps per stage (multiple of 8)

__global__ void __launch_bounds__(kConvThreads) k_ffd_conv(const float* __restrict__ x, const double* __restrict__ d,
                                                           const float* __restrict__ weights,
                                                           const int32_t* __restrict__ widths,
                                                           int T, int max_width, float* __restrict__ out) {
    __shared__ __align__(16) float s_w[kChunk];
    __shared__ __align__(16) float s_x[kTile + kChunk];
    const int n = blockIdx.y;
    const int j0 = blockIdx.x * kTile;
    const int Tout = T - max_width;
    const float* __restrict__ xn = x + (size_t)n * T;
    float* __restrict__ on = out + (size_t)n * Tout;
    const int tid = threadIdx.x;
    const int width = widths[n];
    if (!(d[n] > 0.0) || width <= 0) {                       // `if self.d_opt[f] > 0` (ffd.py:84): untouched series
        for (int q = tid; q < kTile; q += kConvThreads) {
            const int j = j0 + q;
            if (j < Tout) on[j] = xn[max_width + j];
        }
        return;
    }
    const float* __restrict__ wn = weights + (size_t)n * T;
    float acc[kPer];
#pragma unroll
    for (int r = 0; r < kPer; ++r) acc[r] = 0.0f;
    for (int k0 = 0; k0 < width; k0 += kChunk) {
        // s_w[kk] = w[k0+kk] (0 beyond `width`); s_x[i] = x[lo + i] with lo = max_width + j0 − (k0 + kChunk − 1), so that the
        // operand of output j0 + 8t + r and tap k0 + kk sits at s_x[8t + r + (kChunk − 1) − kk]
        const int lo = max_width + j0 - (k0 + kChunk - 1);
        __syncthreads();
        for (int q = tid; q < kChunk; q += kConvThreads) s_w[q] = (k0 + q < width) ? wn[k0 + q] : 0.0f;
        for (int q = tid; q < kTile + kChunk; q += kConvThreads) {
            const int xi = lo + q;
            s_x[q] = (xi >= 0 && xi < T) ? xn[xi] : 0.0f;
        }
        __syncthreads();
        const int ngroups = (min(kChunk, width - k0) + 7) >> 3;
        int wb = 8 * tid + kChunk - 8;                       // window base of group 0 (a multiple of 4 floats)
        float xw[16];
        {
            const float4 a0 = *reinterpret_cast<const float4*>(s_x + wb + 8), a1 = *reinterpret_cast<const float4*>(s_x + wb + 12);
            xw[0] = a0.x; xw[1] = a0.y; xw[2] = a0.z; xw[3] = a0.w; xw[4] = a1.x; xw[5] = a1.y; xw[6] = a1.z; xw[7] = a1.w;
        }
        for (int g = 0; g < ngroups; ++g, wb -= 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) xw[8 + i] = xw[i];   // the upper half of the window is the previous lower half
            const float4 b0 = *reinterpret_cast<const float4*>(s_x + wb), b1 = *reinterpret_cast<const float4*>(s_x + wb + 4);
            xw[0] = b0.x; xw[1] = b0.y; xw[2] = b0.z; xw[3] = b0.w; xw[4] = b1.x; xw[5] = b1.y; xw[6] = b1.z; xw[7] = b1.w;
            const float4 t0 = *reinterpret_cast<const float4*>(s_w + 8 * g), t1 = *reinterpret_cast<const float4*>(s_w + 8 * g + 4);
            const float tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int r = 0; r < kPer; ++r) acc[r] = fmaf(tw[u], xw[7 - u + r], acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
        const int j = j0 + 8 * tid + r;
        if (j < Tout) on[j] = acc[r];
    }
}

// ---- per-series scaling (one block per series) ----------------------------------------------------------
constexpr int kScaleThreads = 256;

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    T r = sm[0];
    for (int i = 1; i < kScaleThreads / 32; ++i) r = op(r, sm[i]);
    return r;
}

__global__ void __launch_bounds__(kScaleThreads) k_scale_series(const float* __restrict__ x, int L, int method,
                                                                float* __restrict__ out) {
    __shared__ double sm_d[kScaleThreads / 32];
    __shared__ float sm_f[kScaleThreads / 32];
    const int n = blockIdx.x, tid = threadIdx.x;
    const float* __restrict__ xn = x + (size_t)n * L;
    float* __restrict__ on = out + (size_t)n * L;
    if (method == 0) {
        // sklearn MinMaxScaler in fp32: scale_ = 1/(max−min) (1 where the range is 0), min_ = 0 − min·scale_,
        // X *= scale_; X += min_   (two roundings, no FMA)
        float mn = INFINITY, mx = -INFINITY;
        for (int i = tid; i < L; i += kScaleThreads) { const float v = xn[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
        mn = block_reduce(mn, [](float a, float b) { return fminf(a, b); }, sm_f);
        mx = block_reduce(mx, [](float a, float b) { return fmaxf(a, b); }, sm_f);
        float range = __fsub_rn(mx, mn);
        if (range < 10.0f * 1.1920929e-07f) range = 1.0f;          // _handle_zeros_in_scale
        const float sc = __fdiv_rn(1.0f, range);
        const float off = __fsub_rn(0.0f, __fmul_rn(mn, sc));
        for (int i = tid; i < L; i += kScaleThreads) on[i] = __fadd_rn(__fmul_rn(xn[i], sc), off);
    } else {
        // sklearn StandardScaler: mean/var accumulated in fp64 (population variance), X −= mean; X /= sqrt(var)
        double s = 0.0;
        for (int i = tid; i < L; i += kScaleThreads) s += (double)xn[i];
        s = block_reduce(s, [](double a, double b) { return a + b; }, sm_d);
        const double mean = s / (double)L;
        double q = 0.0;
        for (int i = tid; i < L; i += kScaleThreads) { const double dlt = (double)xn[i] - mean; q += dlt * dlt; }
        q = block_reduce(q, [](double a, double b) { return a + b; }, sm_d);
        double sd = sqrt(q / (double)L);
        if (sd < 10.0 * 2.220446049250313e-16) sd = 1.0;
        for (int i = tid; i < L; i += kScaleThreads) {
            const float c = (float)((double)xn[i] - mean);          // in-place fp32 array −= fp64 mean
            on[i] = (float)((double)c / sd);
        }
    }
}

// ---- table packing ----------------------------------------------------------------------------------------
// series [A*C, L] → feat_am [A, L, C]: per asset a C×L → L×C transpose; writes coalesced, reads L2-resident.
__global__ void k_pack_feat(const float* __restrict__ series, int A, int C, int L, float* __restrict__ feat_am) {
    const size_t total = (size_t)A * L * C;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(q % C);
        const size_t r = q / C;
        const int t = (int)(r % L);
        const int a = (int)(r / L);
        feat_am[q] = series[((size_t)a * C + c) * L + t];
    }
}
// close [A, L] → close_tm [L, A] through a 32×33 shared tile.
__global__ void k_pack_close(const float* __restrict__ close, int A, int L, float* __restrict__ close_tm) {
    __shared__ float tile[32][33];
    const int t0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int a = a0 + r, t = t0 + threadIdx.x;
        if (a < A && t < L) tile[r][threadIdx.x] = close[(size_t)a * L + t];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int t = t0 + r, a = a0 + threadIdx.x;
        if (a < A && t < L) close_tm[(size_t)t * A + a] = tile[threadIdx.x][r];
    }
}

}  // namespace pmrl

using namespace pmrl;

extern "C" int pmrl_ffd_weights(const double* d, int32_t N, int32_t T, float thres,
                                float* weights, int32_t* widths, void* stream) {
    if (!d || !weights || !widths) return pmrl_fail(PMRL_E_ARG, "ffd_weights: NULL pointer");
    if (N < 0 || T < 2) return pmrl_fail(PMRL_E_SHAPE, "ffd_weights: N >= 0 and T >= 2 required");
    if (T > (1 << 24)) return pmrl_fail(PMRL_E_SHAPE, "ffd_weights: T must stay exactly representable in fp32");
    if (N == 0) return 0;
    k_ffd_weights<<<N, kWThreads, 0, (cudaStream_t)stream>>>(d, T, thres, weights, widths);
    return pmrl_check_launch("k_ffd_weights");
}

extern "C" int pmrl_ffd_transform(const float* x, const double* d, const float* weights, const int32_t* widths,
                                  int32_t N, int32_t T, int32_t max_width, float* out, void* stream) {
    if (!x || !d || !weights || !widths || !out) return pmrl_fail(PMRL_E_ARG, "ffd_transform: NULL pointer");
    if (N < 0 || T < 2 || max_width < 0 || max_width >= T) return pmrl_fail(PMRL_E_SHAPE, "ffd_transform: need 0 <= max_width < T");
    if (N == 0) return 0;
    const int Tout = T - max_width;
    dim3 grid((Tout + kTile - 1) / kTile, N);
    k_ffd_conv<<<grid, kConvThreads, 0, (cudaStream_t)stream>>>(x, d, weights, widths, T, max_width, out);
    return pmrl_check_launch("k_ffd_conv");
}

extern "C" int pmrl_scale_series(const float* x, int32_t N, int32_t L, int32_t method, float* out, void* stream) {
    if (!x || !out) return pmrl_fail(PMRL_E_ARG, "scale_series: NULL pointer");
    if (method != 0 && method != 1) return pmrl_fail(PMRL_E_ARG, "scale_series: method must be 0 (minmax) or 1 (standard)");
    if (N < 0 || L < 1) return pmrl_fail(PMRL_E_SHAPE, "scale_series: N >= 0, L >= 1 required");
    if (N == 0) return 0;
    k_scale_series<<<N, kScaleThreads, 0, (cudaStream_t)stream>>>(x, L, method, out);
    return pmrl_check_launch("k_scale_series");
}

extern "C" int pmrl_pack_features(const float* series, const float* close, int32_t A, int32_t C, int32_t L,
                                  float* feat_am, float* close_tm, void* stream) {
    if (A < 1 || C < 1 || L < 1) return pmrl_fail(PMRL_E_SHAPE, "pack_features: A, C, L >= 1 required");
    if ((series == nullptr) != (feat_am == nullptr) || (close == nullptr) != (close_tm == nullptr))
        return pmrl_fail(PMRL_E_ARG, "pack_features: series/feat_am and close/close_tm come in pairs");
    cudaStream_t s = (cudaStream_t)stream;
    if (series) {
        const size_t total = (size_t)A * L * C;
        const int blocks = (int)((total + 255) / 256 < (size_t)pmrl_sm_count() * 16 ? (total + 255) / 256 : (size_t)pmrl_sm_count() * 16);
        k_pack_feat<<<blocks, 256, 0, s>>>(series, A, C, L, feat_am);
        if (int rc = pmrl_check_launch("k_pack_feat")) return rc;
    }
    if (close) {
        dim3 grid((L + 31) / 32, (A + 31) / 32), block(32, 8);
        k_pack_close<<<grid, block, 0, s>>>(close, A, L, close_tm);
        if (int rc = pmrl_check_launch("k_pack_close")) return rc;
    }
    return 0;
}
