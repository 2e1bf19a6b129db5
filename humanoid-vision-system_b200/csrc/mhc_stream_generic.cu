// K1 forward for every stream shape the tuned kernel (mhc_stream_fwd.cu: n = 4, C = 512) does not take:
// n in {2, 4} residual streams, C a multiple of 8 up to 1024 channels, and the fp32-accurate projection operand
// (HVS_MHC_SPLIT_PHI) for all shapes.  Same arithmetic contract: x exact bf16; RMS statistics, gates,
// Sinkhorn, accumulation fp32; the operand scale*phi rounded to bf16 unless HVS_MHC_SPLIT_PHI; one rounding of y to bf16.
// Reference primitives: RMSNorm src/models/manifold_layers.py:449-456, gates :213/:216, SinkhornKnoppProjection.forward
// :56-77 (batched branch).
//
// A warp per token, CUDA-core FMAs (no tensor-core tiling: this is the general path, not the roofline one): lane l owns
// the 8-channel groups l, l + 32, ... of EVERY stream, so the mixing is lane-local; the n*n + 2n logit partials and
// the sum of squares are butterfly-reduced; the coefficient math is done redundantly by all lanes (no divergence); the
// projection rows are read through L1 (the operand is <= 393 KB, L2-resident).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace hvs {
namespace {

constexpr int kMaxVec = 4;                 // 8-channel groups per lane per stream: C <= 1024

struct GenParams {
    const __nv_bfloat16* x;
    const float* phi;
    const float* bias;
    const float* alpha;
    const float* scale;
    __nv_bfloat16* y;
    __nv_bfloat16* u;
    float* coeffs;
    int64_t T;
    int C, iters, split;
    float eps_rms, eps_sk;
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoid_acc(float v) { return __fdiv_rn(1.0f, 1.0f + expf(-v)); }

template <int N>
__global__ void __launch_bounds__(128) mhc_stream_generic_fwd_kernel(const GenParams p) {
    constexpr int K = N * N + 2 * N;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int C = p.C, nvec = (C + 255) / 256;
    float bias[K];
#pragma unroll
    for (int l = 0; l < K; ++l) bias[l] = p.bias[l];
    const float a_pre = p.alpha[0], a_post = p.alpha[1], a_res = p.alpha[2];
    for (int64_t tok = warp; tok < p.T; tok += nwarps) {
        const __nv_bfloat16* xt = p.x + tok * N * C;
        uint4 xr[N][kMaxVec];
        float acc[K];
#pragma unroll
        for (int l = 0; l < K; ++l) acc[l] = 0.f;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < N; ++j) {
#pragma unroll
            for (int iv = 0; iv < kMaxVec; ++iv) {
                const int c0 = 8 * (lane + 32 * iv);
                xr[j][iv] = make_uint4(0u, 0u, 0u, 0u);
                if (iv < nvec && c0 < C) {
                    xr[j][iv] = *reinterpret_cast<const uint4*>(xt + j * C + c0);
                    const uint32_t w4[4] = {xr[j][iv].x, xr[j][iv].y, xr[j][iv].z, xr[j][iv].w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float xv = (e & 1) ? bf16hi(w4[e >> 1]) : bf16lo(w4[e >> 1]);
                        ss = fmaf(xv, xv, ss);
                        const int k = j * C + c0 + e;
                        const float sc = __ldg(p.scale + k);
                        const float4* row = reinterpret_cast<const float4*>(p.phi + (size_t)k * K);
#pragma unroll
                        for (int q = 0; q < K / 4; ++q) {
                            const float4 f = __ldg(row + q);
                            float w0 = sc * f.x, w1 = sc * f.y, w2 = sc * f.z, w3 = sc * f.w;
                            if (!p.split) {          // operand rounded to bf16 (the reference's autocast convention)
                                w0 = __bfloat162float(__float2bfloat16_rn(w0)); w1 = __bfloat162float(__float2bfloat16_rn(w1));
                                w2 = __bfloat162float(__float2bfloat16_rn(w2)); w3 = __bfloat162float(__float2bfloat16_rn(w3));
                            }
                            acc[4 * q] = fmaf(xv, w0, acc[4 * q]); acc[4 * q + 1] = fmaf(xv, w1, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(xv, w2, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv, w3, acc[4 * q + 3]);
                        }
                    }
                }
            }
        }
        ss = wsum(ss);
#pragma unroll
        for (int l = 0; l < K; ++l) acc[l] = wsum(acc[l]);
        const float inv_rms = 1.0f / sqrtf(ss / (float)(N * C) + p.eps_rms);          // :451
        float hpre[N], hpost[N], pm[N * N];
#pragma unroll
        for (int j = 0; j < N; ++j) hpre[j] = sigmoid_acc(fmaf(a_pre, acc[j] * inv_rms, bias[j]));                   // :213
#pragma unroll
        for (int i = 0; i < N; ++i) hpost[i] = 2.0f * sigmoid_acc(fmaf(a_post, acc[N + i] * inv_rms, bias[N + i]));   // :216
#pragma unroll
        for (int i = 0; i < N; ++i) {                                                  // softmax(row) * m  (:56-57)
            float lg[N], mx = -INFINITY, sum = 0.f;
#pragma unroll
            for (int j = 0; j < N; ++j) { lg[j] = fmaf(a_res, acc[2 * N + i * N + j] * inv_rms, bias[2 * N + i * N + j]); mx = fmaxf(mx, lg[j]); }
#pragma unroll
            for (int j = 0; j < N; ++j) { lg[j] = expf(lg[j] - mx); sum += lg[j]; }
#pragma unroll
            for (int j = 0; j < N; ++j) pm[i * N + j] = __fdiv_rn(lg[j], sum) * (float)N;
        }
        for (int it = 0; it < p.iters; ++it) {                                         // :64-72
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float rs = 0.f;
#pragma unroll
                for (int j = 0; j < N; ++j) rs += pm[i * N + j];
                rs += p.eps_sk;
#pragma unroll
                for (int j = 0; j < N; ++j) pm[i * N + j] = __fdiv_rn(pm[i * N + j], rs);
            }
#pragma unroll
            for (int j = 0; j < N; ++j) {
                float cs = 0.f;
#pragma unroll
                for (int i = 0; i < N; ++i) cs += pm[i * N + j];
                cs += p.eps_sk;
#pragma unroll
                for (int i = 0; i < N; ++i) pm[i * N + j] = __fdiv_rn(pm[i * N + j], cs);
            }
        }
        if (p.coeffs != nullptr && lane == 0) {
            float* co = p.coeffs + tok * K;
#pragma unroll
            for (int j = 0; j < N; ++j) { co[j] = hpre[j]; co[N + j] = hpost[j]; }
#pragma unroll
            for (int e = 0; e < N * N; ++e) co[2 * N + e] = pm[e];
        }
        // mixing, lane-local: u = H_pre^T x;  y_i = sum_j H_res[i][j] x_j + H_post[i] u   (fp32, one rounding)
#pragma unroll
        for (int iv = 0; iv < kMaxVec; ++iv) {
            const int c0 = 8 * (lane + 32 * iv);
            if (iv >= nvec || c0 >= C) continue;
            float xf[N][8], uf[8];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const uint32_t w4[4] = {xr[j][iv].x, xr[j][iv].y, xr[j][iv].z, xr[j][iv].w};
#pragma unroll
                for (int e = 0; e < 8; ++e) xf[j][e] = (e & 1) ? bf16hi(w4[e >> 1]) : bf16lo(w4[e >> 1]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < N; ++j) a = fmaf(hpre[j], xf[j][e], a);
                uf[e] = a;
            }
            if (p.u != nullptr)
                *reinterpret_cast<uint4*>(p.u + tok * C + c0) = make_uint4(pack_bf16(uf[0], uf[1]), pack_bf16(uf[2], uf[3]),
                                                                           pack_bf16(uf[4], uf[5]), pack_bf16(uf[6], uf[7]));
            if (p.y != nullptr) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float a = 0.f;
#pragma unroll
                        for (int j = 0; j < N; ++j) a = fmaf(pm[i * N + j], xf[j][e], a);
                        o[e] = fmaf(hpost[i], uf[e], a);
                    }
                    *reinterpret_cast<uint4*>(p.y + (tok * N + i) * C + c0) =
                        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
                }
            }
        }
    }
}

// y_i = sum_j H_res[i][j] x_j + H_post[i] fu  for coefficients produced by the forward (wrapped-layer path), any shape
template <int N>
__global__ void __launch_bounds__(256) mhc_stream_generic_post_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ coeffs,
                                                                       const __nv_bfloat16* __restrict__ fu, __nv_bfloat16* __restrict__ y,
                                                                       int64_t T, int C) {
    constexpr int K = N * N + 2 * N;
    const int groups = C / 8;
    const int64_t total = T * groups;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tok = idx / groups;
        const int c0 = (int)(idx % groups) * 8;
        const float* co = coeffs + tok * K;
        float xf[N][8], ff[8];
        const uint4 fv = *reinterpret_cast<const uint4*>(fu + tok * C + c0);
        const uint32_t f4[4] = {fv.x, fv.y, fv.z, fv.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) ff[e] = (e & 1) ? bf16hi(f4[e >> 1]) : bf16lo(f4[e >> 1]);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const uint4 v = *reinterpret_cast<const uint4*>(x + (tok * N + j) * C + c0);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) xf[j][e] = (e & 1) ? bf16hi(w4[e >> 1]) : bf16lo(w4[e >> 1]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < N; ++j) a = fmaf(co[2 * N + i * N + j], xf[j][e], a);
                o[e] = fmaf(co[N + i], ff[e], a);
            }
            *reinterpret_cast<uint4*>(y + (tok * N + i) * C + c0) =
                make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
    }
}

}  // namespace

bool generic_stream_shape_ok(int n, int C) { return (n == 2 || n == 4) && C >= 8 && C % 8 == 0 && C <= 1024; }

int launch_generic_stream_fwd(const void* x, const float* phi, const float* bias, const float* alpha, const float* scale, void* y,
                              void* u, float* coeffs, int64_t T, int n, int C, int iters, float eps_rms, float eps_sk, int split,
                              cudaStream_t stream) {
    GenParams p{reinterpret_cast<const __nv_bfloat16*>(x), phi, bias, alpha, scale, reinterpret_cast<__nv_bfloat16*>(y),
                reinterpret_cast<__nv_bfloat16*>(u), coeffs, T, C, iters, split, eps_rms, eps_sk};
    int64_t blocks = (T + 3) / 4;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (n == 2) mhc_stream_generic_fwd_kernel<2><<<(int)blocks, 128, 0, stream>>>(p);
    else mhc_stream_generic_fwd_kernel<4><<<(int)blocks, 128, 0, stream>>>(p);
    count_launch();
    return launch_status();
}

int launch_generic_stream_post(const void* x, const float* coeffs, const void* fu, void* y, int64_t T, int n, int C, cudaStream_t stream) {
    int64_t blocks = (T * (C / 8) + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (n == 2)
        mhc_stream_generic_post_kernel<2><<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), coeffs,
            reinterpret_cast<const __nv_bfloat16*>(fu), reinterpret_cast<__nv_bfloat16*>(y), T, C);
    else
        mhc_stream_generic_post_kernel<4><<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), coeffs,
            reinterpret_cast<const __nv_bfloat16*>(fu), reinterpret_cast<__nv_bfloat16*>(y), T, C);
    count_launch();
    return launch_status();
}

}  // namespace hvs
