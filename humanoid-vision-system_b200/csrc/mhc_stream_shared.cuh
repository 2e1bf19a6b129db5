// Shapes and the per-token coefficient math shared by the stream-mHC forward and backward kernels.
#pragma once
#include <cuda_runtime.h>

#include "ptx_sm100.cuh"

namespace hvs {

constexpr int kN = 4;                 // residual streams
constexpr int kC = 512;               // channels per stream
constexpr int kRow = kN * kC;         // flattened row the RMSNorm and the projection see
constexpr int kL = kN * kN + 2 * kN;  // logits per token: H_pre | H_post | H_res
constexpr int kTileTok = 16;          // tokens per tile (= M of the warp MMA)
constexpr int kRowBytes = kRow * 2;

// exp through the hardware ex2 unit: relative error ~2^-22 plus |v| * 2^-24 from the argument scaling;
// the coefficient tests bound the end-to-end deviation from the fp32 oracle at 1e-5 relative.
__device__ __forceinline__ float fast_exp(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v * 1.4426950408889634f));
    return r;
}
__device__ __forceinline__ float sigmoid_f32(float v) { return rcp_approx(1.0f + fast_exp(-v)); }

// Sinkhorn-Knopp on a 4x4 block held in registers (SinkhornKnoppProjection.forward,
// src/models/manifold_layers.py:56-77): softmax over each row times m (:57), then `iters` x
// { row / (row_sum + eps) ; column / (col_sum + eps) }.  Normalisations multiply by the hardware
// reciprocal (<= 1 ulp) instead of dividing: the iteration renormalises, so the deviation from the
// reference's IEEE division stays at the 1e-7 level (tests bound it at 1e-5 relative).
__device__ __forceinline__ void sinkhorn4x4(float (&pm)[16], int iters, float eps) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float mx = fmaxf(fmaxf(pm[4 * i], pm[4 * i + 1]), fmaxf(pm[4 * i + 2], pm[4 * i + 3]));
        const float e0 = fast_exp(pm[4 * i] - mx), e1 = fast_exp(pm[4 * i + 1] - mx);
        const float e2 = fast_exp(pm[4 * i + 2] - mx), e3 = fast_exp(pm[4 * i + 3] - mx);
        const float r4 = 4.0f * rcp_approx((e0 + e1) + (e2 + e3));
        pm[4 * i] = e0 * r4;
        pm[4 * i + 1] = e1 * r4;
        pm[4 * i + 2] = e2 * r4;
        pm[4 * i + 3] = e3 * r4;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float r = rcp_approx(((pm[4 * i] + pm[4 * i + 1]) + (pm[4 * i + 2] + pm[4 * i + 3])) + eps);
            pm[4 * i] *= r; pm[4 * i + 1] *= r; pm[4 * i + 2] *= r; pm[4 * i + 3] *= r;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float r = rcp_approx(((pm[j] + pm[4 + j]) + (pm[8 + j] + pm[12 + j])) + eps);
            pm[j] *= r; pm[4 + j] *= r; pm[8 + j] *= r; pm[12 + j] *= r;
        }
    }
}

// raw[24] = x . (scale*phi) un-normalised;  logits = alpha_g * (inv_rms * raw) + bias.
__device__ __forceinline__ void coefficients_from_raw(const float (&raw)[kL], float inv_rms, const float (&bias)[kL],
                                                      float a_pre, float a_post, float a_res, int iters, float eps,
                                                      float (&hpre)[kN], float (&hpost)[kN], float (&pm)[kN * kN]) {
#pragma unroll
    for (int j = 0; j < kN; ++j) hpre[j] = sigmoid_f32(fmaf(a_pre, raw[j] * inv_rms, bias[j]));
#pragma unroll
    for (int i = 0; i < kN; ++i) hpost[i] = 2.0f * sigmoid_f32(fmaf(a_post, raw[kN + i] * inv_rms, bias[kN + i]));
#pragma unroll
    for (int e = 0; e < kN * kN; ++e) pm[e] = fmaf(a_res, raw[2 * kN + e] * inv_rms, bias[2 * kN + e]);
    sinkhorn4x4(pm, iters, eps);
}

}  // namespace hvs
