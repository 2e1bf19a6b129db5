// Row normalisations of the mHC path, one warp per row (rows are 32...4096 elements, i.e. L1-resident):
//   RMSNorm.forward                         src/models/manifold_layers.py:449-456   x / sqrt(mean(x^2) + eps) * scale
//   nn.LayerNorm around the module's token path (norm_pre :250, norm_post :267)
// Statistics and arithmetic are fp32 whatever the storage type; the mean and the variance are two separate
// passes over the (cached) row, so no E[x^2] - mean^2 cancellation.  Cross-row reductions (dscale) are two-stage
// with a fixed order: bitwise reproducible.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace hvs {
namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T> __device__ __forceinline__ float ldf(const T* p, int64_t i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T* p, int64_t i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, int64_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------- RMSNorm
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
rmsnorm_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ scale, TO* __restrict__ out, int64_t rows,
                   int dim, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const TI* xr = x + r * dim;
        float ss = 0.f;
        for (int j = lane; j < dim; j += 32) { const float v = ldf(xr, j); ss = fmaf(v, v, ss); }
        ss = wsum(ss);
        const float rms = sqrtf(ss / (float)dim + eps);                 // :451
        for (int j = lane; j < dim; j += 32) stf(out + r * dim, j, __fdiv_rn(ldf(xr, j), rms) * scale[j]);   // :454
    }
}

// dx = scale*dy/rms - x * sum_j(dy_j scale_j x_j) / (dim rms^3);  dscale partial per CTA (fixed order inside the CTA)
template <typename TI, typename TG>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const TI* __restrict__ x, const float* __restrict__ scale, const TG* __restrict__ dy,
                   TG* __restrict__ dx, float* __restrict__ dscale_part, int64_t rows, int dim, float eps) {
    extern __shared__ float sm_ds[];                                    // [8 warps][dim]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* mine = sm_ds + (size_t)w * dim;
    for (int j = lane; j < dim; j += 32) mine[j] = 0.f;
    const int64_t warp = (int64_t)blockIdx.x * 8 + w;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const TI* xr = x + r * dim;
        const TG* gr = dy + r * dim;
        float ss = 0.f, dot = 0.f;
        for (int j = lane; j < dim; j += 32) {
            const float v = ldf(xr, j);
            ss = fmaf(v, v, ss);
            dot = fmaf(ldf(gr, j) * scale[j], v, dot);
        }
        ss = wsum(ss);
        dot = wsum(dot);
        const float ms = ss / (float)dim + eps;
        const float inv = rsqrtf(ms);
        const float k = dot * inv / ((float)dim * ms);
        for (int j = lane; j < dim; j += 32) {
            const float v = ldf(xr, j), g = ldf(gr, j);
            stf(dx + r * dim, j, g * scale[j] * inv - v * k);
            mine[j] += g * v * inv;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < dim; j += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += sm_ds[(size_t)k * dim + j];
        dscale_part[(size_t)blockIdx.x * dim + j] = s;
    }
}

__global__ void colsum_finalize_kernel(const float* __restrict__ part, float* __restrict__ out, int parts, int dim) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    float s = 0.f;
    for (int p = 0; p < parts; ++p) s += part[(size_t)p * dim + j];
    out[j] = s;
}

// ---------------------------------------------------------------------------- LayerNorm (forward, K2 prologue / epilogue)
// out = (x - mean) / sqrt(var + eps) * w + b; optionally also a plain bf16 copy of x (the A operand of x @ H_res),
// both written into rows padded to `out_ld` / `copy_ld` elements (pad columns zero-filled so they can sit in a GEMM K).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                     TO* __restrict__ out, __nv_bfloat16* __restrict__ copy, int64_t rows, int dim, int out_ld,
                     int copy_ld, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const TI* xr = x + r * dim;
        float s = 0.f;
        for (int j = lane; j < dim; j += 32) s += ldf(xr, j);
        const float mean = wsum(s) / (float)dim;
        float q = 0.f;
        for (int j = lane; j < dim; j += 32) { const float d = ldf(xr, j) - mean; q = fmaf(d, d, q); }
        const float inv = rsqrtf(wsum(q) / (float)dim + eps);
        TO* o = out + r * out_ld;
        for (int j = lane; j < out_ld; j += 32)
            stf(o, j, j < dim ? (ldf(xr, j) - mean) * inv * w[j] + b[j] : 0.f);
        if (copy != nullptr) {
            __nv_bfloat16* c = copy + r * copy_ld;
            for (int j = lane; j < copy_ld; j += 32) c[j] = __float2bfloat16_rn(j < dim ? ldf(xr, j) : 0.f);
        }
    }
}

// Small rows (D = 32 ... 512, the widths of the backbone's mHC layers, where T is in the millions): D / 8 lanes per row
// (16 for D = 512 with two vectors each), every lane holds 8 (16) consecutive elements from one 16-byte load, statistics
// by xor-shuffles inside the lane group, 16-byte stores.  The warp-per-row kernel above moves 64 bytes per warp
// instruction at D = 32; this one moves 512.
template <typename TI> struct Load8;
template <> struct Load8<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&f)[8]) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
        f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
        f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
    }
};
template <> struct Load8<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&f)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
__device__ __forceinline__ uint4 pack8_bf16(const float (&f)[8]) {
    uint4 o;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(f[1]), "f"(f[0]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(f[3]), "f"(f[2]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.z) : "f"(f[5]), "f"(f[4]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.w) : "f"(f[7]), "f"(f[6]));
    return o;
}

template <typename TO> __device__ __forceinline__ void store8(TO* p, const float (&f)[8]);
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) { *reinterpret_cast<uint4*>(p) = pack8_bf16(f); }
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

template <typename TI, typename TO, int D>
__global__ void __launch_bounds__(256)
layernorm_small_rows_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                            TO* __restrict__ out, __nv_bfloat16* __restrict__ copy, int64_t rows, float eps) {
    constexpr int kVec = D >= 512 ? D / 256 : 1;            // 8-element vectors per lane
    constexpr int kLanes = D / (8 * kVec);                  // lanes per row (4 ... 32)
    constexpr int kRowsPerWarp = 32 / kLanes;
    const int lane = threadIdx.x & 31;
    const int sub = lane / kLanes, l = lane % kLanes;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float wv[kVec][8], bv[kVec][8];
#pragma unroll
    for (int v = 0; v < kVec; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) { wv[v][j] = w[(v * kLanes + l) * 8 + j]; bv[v][j] = b[(v * kLanes + l) * 8 + j]; }
    for (int64_t r0 = warp * kRowsPerWarp; r0 < rows; r0 += nwarps * kRowsPerWarp) {
        const int64_t r = r0 + sub;
        const bool ok = r < rows;
        float f[kVec][8];
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < kVec; ++v) {
            if (ok) Load8<TI>::ld(x + r * D + (v * kLanes + l) * 8, f[v]);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[v][j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s += f[v][j];
        }
#pragma unroll
        for (int o = kLanes / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / (float)D);
        float q = 0.f;
#pragma unroll
        for (int v = 0; v < kVec; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = f[v][j] - mean; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = kLanes / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float inv = rsqrtf(q * (1.0f / (float)D) + eps);
        if (ok) {
#pragma unroll
            for (int v = 0; v < kVec; ++v) {
                if (copy != nullptr) *reinterpret_cast<uint4*>(copy + r * D + (v * kLanes + l) * 8) = pack8_bf16(f[v]);
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = (f[v][j] - mean) * inv * wv[v][j] + bv[v][j];
                store8<TO>(out + r * D + (v * kLanes + l) * 8, y);
            }
        }
    }
}

template <typename TI, typename TO>
bool launch_small_rows(const TI* x, const float* w, const float* b, TO* out, __nv_bfloat16* copy, int64_t rows, int dim,
                       float eps, cudaStream_t stream) {
    int64_t blocks = (rows * (dim >= 512 ? 32 : dim / 8) / 32 + 7) / 8;      // 8 warps per block
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const int g = (int)blocks;
    switch (dim) {
        case 32: layernorm_small_rows_kernel<TI, TO, 32><<<g, 256, 0, stream>>>(x, w, b, out, copy, rows, eps); return true;
        case 64: layernorm_small_rows_kernel<TI, TO, 64><<<g, 256, 0, stream>>>(x, w, b, out, copy, rows, eps); return true;
        case 128: layernorm_small_rows_kernel<TI, TO, 128><<<g, 256, 0, stream>>>(x, w, b, out, copy, rows, eps); return true;
        case 256: layernorm_small_rows_kernel<TI, TO, 256><<<g, 256, 0, stream>>>(x, w, b, out, copy, rows, eps); return true;
        case 512: layernorm_small_rows_kernel<TI, TO, 512><<<g, 256, 0, stream>>>(x, w, b, out, copy, rows, eps); return true;
        default: return false;
    }
}

// LayerNorm backward for the same small rows (training of the module's norm_pre / norm_post, fp32 statistics):
//   xhat = (x - mean) rstd, g = dy w:  dx = rstd (g - mean(g) - xhat mean(g xhat));  dw = sum_rows dy xhat;  db = sum_rows dy.
// A lane owns the same columns for every row it meets, so dw / db accumulate in registers; they are combined across the
// row sub-groups of the warp, the warps of the CTA (shared memory) and the CTAs (workspace + finalize), all in fixed order.
template <typename TI, typename TG, int D>
__global__ void __launch_bounds__(256)
layernorm_small_rows_bwd_kernel(const TI* __restrict__ x, const float* __restrict__ w, const TG* __restrict__ dy,
                                TI* __restrict__ dx, float* __restrict__ part, int64_t rows, float eps) {
    constexpr int kVec = D >= 512 ? D / 256 : 1;
    constexpr int kLanes = D / (8 * kVec);
    constexpr int kRowsPerWarp = 32 / kLanes;
    __shared__ float sm[8][2][D];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int sub = lane / kLanes, l = lane % kLanes;
    const int64_t warp = (int64_t)blockIdx.x * 8 + wid;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    float wv[kVec][8], aw[kVec][8], ab[kVec][8];
#pragma unroll
    for (int v = 0; v < kVec; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) { wv[v][j] = w[(v * kLanes + l) * 8 + j]; aw[v][j] = 0.f; ab[v][j] = 0.f; }
    for (int64_t r0 = warp * kRowsPerWarp; r0 < rows; r0 += nwarps * kRowsPerWarp) {
        const int64_t r = r0 + sub;
        const bool ok = r < rows;
        float f[kVec][8], g[kVec][8];
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < kVec; ++v) {
            if (ok) {
                Load8<TI>::ld(x + r * D + (v * kLanes + l) * 8, f[v]);
                Load8<TG>::ld(dy + r * D + (v * kLanes + l) * 8, g[v]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) { f[v][j] = 0.f; g[v][j] = 0.f; }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s += f[v][j];
        }
#pragma unroll
        for (int o = kLanes / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / (float)D);
        float q = 0.f;
#pragma unroll
        for (int v = 0; v < kVec; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = f[v][j] - mean; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = kLanes / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / (float)D) + eps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int v = 0; v < kVec; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (f[v][j] - mean) * rstd;
                f[v][j] = xh;
                aw[v][j] = fmaf(g[v][j], xh, aw[v][j]);
                ab[v][j] += g[v][j];
                g[v][j] *= wv[v][j];
                sg += g[v][j];
                sgx = fmaf(g[v][j], xh, sgx);
            }
#pragma unroll
        for (int o = kLanes / 2; o > 0; o >>= 1) { sg += __shfl_xor_sync(0xffffffffu, sg, o); sgx += __shfl_xor_sync(0xffffffffu, sgx, o); }
        sg *= (1.0f / (float)D); sgx *= (1.0f / (float)D);
        if (ok) {
#pragma unroll
            for (int v = 0; v < kVec; ++v) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = rstd * (g[v][j] - sg - f[v][j] * sgx);
                store8<TI>(dx + r * D + (v * kLanes + l) * 8, o);
            }
        }
    }
    // combine the row sub-groups of the warp (lanes with the same l), then the warps, then write the CTA partial
#pragma unroll
    for (int v = 0; v < kVec; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int o = kLanes; o < 32; o <<= 1) {
                aw[v][j] += __shfl_xor_sync(0xffffffffu, aw[v][j], o);
                ab[v][j] += __shfl_xor_sync(0xffffffffu, ab[v][j], o);
            }
            if (sub == 0) { sm[wid][0][(v * kLanes + l) * 8 + j] = aw[v][j]; sm[wid][1][(v * kLanes + l) * 8 + j] = ab[v][j]; }
        }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * D; c += 256) {
        const int which = c / D, col = c % D;
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][which][col];
        part[((size_t)blockIdx.x * 2 + which) * D + col] = t;
    }
}

// (a warp per column: lanes stride over the partials, then a fixed-order butterfly -- hundreds of partials summed by one
// thread each was 16 us per launch, 152 launches per training step)
__global__ void ln_bwd_finalize_kernel(const float* __restrict__ part, float* __restrict__ dw, float* __restrict__ db, int parts, int dim) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= dim) return;
    float a = 0.f, b = 0.f;
    for (int p = lane; p < parts; p += 32) { a += part[((size_t)p * 2) * dim + j]; b += part[((size_t)p * 2 + 1) * dim + j]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { dw[j] = a; db[j] = b; }
}

template <typename TI, typename TG>
bool launch_small_rows_bwd(const TI* x, const float* w, const TG* dy, TI* dx, float* part, int blocks, int64_t rows, int dim, float eps,
                           cudaStream_t stream) {
    switch (dim) {
        case 32: layernorm_small_rows_bwd_kernel<TI, TG, 32><<<blocks, 256, 0, stream>>>(x, w, dy, dx, part, rows, eps); return true;
        case 64: layernorm_small_rows_bwd_kernel<TI, TG, 64><<<blocks, 256, 0, stream>>>(x, w, dy, dx, part, rows, eps); return true;
        case 128: layernorm_small_rows_bwd_kernel<TI, TG, 128><<<blocks, 256, 0, stream>>>(x, w, dy, dx, part, rows, eps); return true;
        case 256: layernorm_small_rows_bwd_kernel<TI, TG, 256><<<blocks, 256, 0, stream>>>(x, w, dy, dx, part, rows, eps); return true;
        case 512: layernorm_small_rows_bwd_kernel<TI, TG, 512><<<blocks, 256, 0, stream>>>(x, w, dy, dx, part, rows, eps); return true;
        default: return false;
    }
}

// Wide rows (D = 1024, 1792: a few thousand tokens at most): a warp per row for dx and the row statistics, then column
// sums over row chunks.  Same arithmetic and the same fixed-order partial layout as the small-row kernel.
template <typename TI, typename TG>
__global__ void __launch_bounds__(256)
layernorm_wide_bwd_dx_kernel(const TI* __restrict__ x, const float* __restrict__ w, const TG* __restrict__ dy, TI* __restrict__ dx,
                             float2* __restrict__ stats, int64_t rows, int dim, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    for (int64_t r = warp; r < rows; r += (int64_t)gridDim.x * 8) {
        const TI* xr = x + r * dim;
        const TG* gr = dy + r * dim;
        float s = 0.f;
        for (int j = lane; j < dim; j += 32) s += ldf(xr, j);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)dim;
        float q = 0.f;
        for (int j = lane; j < dim; j += 32) { const float d = ldf(xr, j) - mean; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / (float)dim + eps);
        float sg = 0.f, sgx = 0.f;
        for (int j = lane; j < dim; j += 32) {
            const float g = ldf(gr, j) * w[j], xh = (ldf(xr, j) - mean) * rstd;
            sg += g;
            sgx = fmaf(g, xh, sgx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sg += __shfl_xor_sync(0xffffffffu, sg, o); sgx += __shfl_xor_sync(0xffffffffu, sgx, o); }
        sg /= (float)dim; sgx /= (float)dim;
        for (int j = lane; j < dim; j += 32) {
            const float g = ldf(gr, j) * w[j], xh = (ldf(xr, j) - mean) * rstd;
            stf(dx + r * dim, j, rstd * (g - sg - xh * sgx));
        }
        if (lane == 0) stats[r] = make_float2(mean, rstd);
    }
}
template <typename TI, typename TG>
__global__ void __launch_bounds__(128)
layernorm_wide_bwd_dw_kernel(const TI* __restrict__ x, const TG* __restrict__ dy, const float2* __restrict__ stats, float* __restrict__ part,
                             int64_t rows, int dim, int rows_per_part) {
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col >= dim) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_part;
    const int64_t r1 = r0 + rows_per_part < rows ? r0 + rows_per_part : rows;
    float aw = 0.f, ab = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
        const float2 st = stats[r];
        const float g = ldf(dy, r * dim + col);
        aw = fmaf(g, (ldf(x, r * dim + col) - st.x) * st.y, aw);
        ab += g;
    }
    part[((size_t)blockIdx.y * 2) * dim + col] = aw;
    part[((size_t)blockIdx.y * 2 + 1) * dim + col] = ab;
}
constexpr int kWideRowsPerPart = 64;
template <typename TI, typename TG>
void launch_wide_bwd(const TI* x, const float* w, const TG* dy, TI* dx, float* part, float2* stats, int parts, int64_t rows, int dim,
                     float eps, cudaStream_t stream) {
    int64_t blocks = (rows + 7) / 8;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    layernorm_wide_bwd_dx_kernel<TI, TG><<<(int)blocks, 256, 0, stream>>>(x, w, dy, dx, stats, rows, dim, eps);
    layernorm_wide_bwd_dw_kernel<TI, TG><<<dim3((dim + 127) / 128, parts), 128, 0, stream>>>(x, dy, stats, part, rows, dim, kWideRowsPerPart);
}
// Signal-ratio monitor of the module (manifold_layers.py:295-303): mean_rows ||out_r|| / (mean_rows ||x_r|| + 1e-8).  A warp per
// row, fixed-order partials per CTA, one small second stage: two launches instead of the reference's eight (two casts, two
// norms, two means, a division, an indexed store).
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
row_norm_sums_kernel(const TA* __restrict__ a, const TB* __restrict__ b, int64_t rows, int dim, float* __restrict__ part) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float sa = 0.f, sb = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * 8 + wid; r < rows; r += (int64_t)gridDim.x * 8) {
        float qa = 0.f, qb = 0.f;
        for (int j = lane; j < dim; j += 32) {
            const float va = ldf(a, r * dim + j), vb = ldf(b, r * dim + j);
            qa = fmaf(va, va, qa);
            qb = fmaf(vb, vb, qb);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { qa += __shfl_xor_sync(0xffffffffu, qa, o); qb += __shfl_xor_sync(0xffffffffu, qb, o); }
        sa += sqrtf(qa);
        sb += sqrtf(qb);
    }
    __shared__ float sm[8][2];
    if (lane == 0) { sm[wid][0] = sa; sm[wid][1] = sb; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
        part[(size_t)blockIdx.x * 2 + threadIdx.x] = t;
    }
}
__global__ void signal_ratio_final_kernel(const float* __restrict__ part, int nparts, int64_t rows, float* __restrict__ dst) {
    const int lane = threadIdx.x;
    float sa = 0.f, sb = 0.f;
    for (int i = lane; i < nparts; i += 32) { sa += part[2 * i]; sb += part[2 * i + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
    if (lane == 0) dst[0] = (sa / (float)rows) / (sb / (float)rows + 1e-8f);
}
inline int row_norm_grid(int64_t rows) {
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

inline bool ln_small_dim(int dim) { return dim == 32 || dim == 64 || dim == 128 || dim == 256 || dim == 512; }

inline int grid_for_rows(int64_t rows) {
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace hvs

extern "C" int hvs_rmsnorm_fwd(const void* x, int x_dtype, const float* scale, void* out, int out_dtype, int64_t rows,
                               int dim, float eps, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows < 0 || dim <= 0 || !scale) return HVS_ERR_BAD_ARG;
    if (rows == 0) return HVS_OK;
    if (!x || !out) return HVS_ERR_BAD_ARG;
    const int g = grid_for_rows(rows);
    if (x_dtype == HVS_DTYPE_F32 && out_dtype == HVS_DTYPE_F32)
        rmsnorm_fwd_kernel<float, float><<<g, 256, 0, stream>>>((const float*)x, scale, (float*)out, rows, dim, eps);
    else if (x_dtype == HVS_DTYPE_BF16 && out_dtype == HVS_DTYPE_BF16)
        rmsnorm_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)x, scale, (__nv_bfloat16*)out, rows, dim, eps);
    else if (x_dtype == HVS_DTYPE_BF16 && out_dtype == HVS_DTYPE_F32)
        rmsnorm_fwd_kernel<__nv_bfloat16, float><<<g, 256, 0, stream>>>((const __nv_bfloat16*)x, scale, (float*)out, rows, dim, eps);
    else if (x_dtype == HVS_DTYPE_F32 && out_dtype == HVS_DTYPE_BF16)
        rmsnorm_fwd_kernel<float, __nv_bfloat16><<<g, 256, 0, stream>>>((const float*)x, scale, (__nv_bfloat16*)out, rows, dim, eps);
    else
        return HVS_ERR_UNSUPPORTED;
    count_launch();
    return launch_status();
}

extern "C" size_t hvs_rmsnorm_bwd_workspace(int64_t rows, int dim) {
    if (rows < 0 || dim <= 0) return 0;
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = (int64_t)hvs::sm_count() * 2;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (size_t)blocks * dim * sizeof(float);
}

extern "C" int hvs_rmsnorm_bwd(const void* x, int dtype, const float* scale, const void* dy, void* dx, float* dscale,
                               int64_t rows, int dim, float eps, void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows < 0 || dim <= 0 || !scale || !dscale) return HVS_ERR_BAD_ARG;
    if (dim > 6144) return HVS_ERR_UNSUPPORTED;                       // 8 warps x dim floats of shared memory
    if (rows > 0 && (!x || !dy || !dx)) return HVS_ERR_BAD_ARG;
    const size_t need = hvs_rmsnorm_bwd_workspace(rows, dim);
    if (!workspace || workspace_bytes < need) return HVS_ERR_WORKSPACE;
    const int blocks = (int)(need / ((size_t)dim * sizeof(float)));
    const int smem = 8 * dim * (int)sizeof(float);
    float* part = (float*)workspace;
    if (dtype == HVS_DTYPE_F32) {
        HVS_SET_MAX_SMEM((rmsnorm_bwd_kernel<float, float>), 8 * 6144 * 4);
        rmsnorm_bwd_kernel<float, float><<<blocks, 256, smem, stream>>>((const float*)x, scale, (const float*)dy, (float*)dx, part, rows, dim, eps);
    } else if (dtype == HVS_DTYPE_BF16) {
        HVS_SET_MAX_SMEM((rmsnorm_bwd_kernel<__nv_bfloat16, __nv_bfloat16>), 8 * 6144 * 4);
        rmsnorm_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, smem, stream>>>((const __nv_bfloat16*)x, scale, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, part, rows, dim, eps);
    } else {
        return HVS_ERR_UNSUPPORTED;
    }
    colsum_finalize_kernel<<<(dim + 127) / 128, 128, 0, stream>>>(part, dscale, blocks, dim);
    count_launch(2);
    return launch_status();
}

extern "C" int hvs_layernorm_fwd(const void* x, int x_dtype, const float* weight, const float* bias, void* out, int out_dtype,
                                 void* x_bf16_copy, int64_t rows, int dim, int out_ld, int copy_ld, float eps, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows < 0 || dim <= 0 || !weight || !bias || out_ld < dim || (x_bf16_copy && copy_ld < dim)) return HVS_ERR_BAD_ARG;
    if (rows == 0) return HVS_OK;
    if (!x || !out) return HVS_ERR_BAD_ARG;
    const int g = grid_for_rows(rows);
    __nv_bfloat16* cp = (__nv_bfloat16*)x_bf16_copy;
    const bool dense = out_ld == dim && (cp == nullptr || copy_ld == dim) &&
                       ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(cp)) & 15) == 0;
    if (dense) {
        bool done = false;
        if (x_dtype == HVS_DTYPE_BF16 && out_dtype == HVS_DTYPE_BF16) done = launch_small_rows((const __nv_bfloat16*)x, weight, bias, (__nv_bfloat16*)out, cp, rows, dim, eps, stream);
        else if (x_dtype == HVS_DTYPE_F32 && out_dtype == HVS_DTYPE_BF16) done = launch_small_rows((const float*)x, weight, bias, (__nv_bfloat16*)out, cp, rows, dim, eps, stream);
        else if (x_dtype == HVS_DTYPE_BF16 && out_dtype == HVS_DTYPE_F32) done = launch_small_rows((const __nv_bfloat16*)x, weight, bias, (float*)out, cp, rows, dim, eps, stream);
        else if (x_dtype == HVS_DTYPE_F32 && out_dtype == HVS_DTYPE_F32) done = launch_small_rows((const float*)x, weight, bias, (float*)out, cp, rows, dim, eps, stream);
        if (done) {
            count_launch();
            return launch_status();
        }
    }
    if (x_dtype == HVS_DTYPE_F32 && out_dtype == HVS_DTYPE_BF16)
        layernorm_fwd_kernel<float, __nv_bfloat16><<<g, 256, 0, stream>>>((const float*)x, weight, bias, (__nv_bfloat16*)out, cp, rows, dim, out_ld, copy_ld, eps);
    else if (x_dtype == HVS_DTYPE_BF16 && out_dtype == HVS_DTYPE_BF16)
        layernorm_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)x, weight, bias, (__nv_bfloat16*)out, cp, rows, dim, out_ld, copy_ld, eps);
    else if (x_dtype == HVS_DTYPE_F32 && out_dtype == HVS_DTYPE_F32)
        layernorm_fwd_kernel<float, float><<<g, 256, 0, stream>>>((const float*)x, weight, bias, (float*)out, cp, rows, dim, out_ld, copy_ld, eps);
    else if (x_dtype == HVS_DTYPE_BF16 && out_dtype == HVS_DTYPE_F32)
        layernorm_fwd_kernel<__nv_bfloat16, float><<<g, 256, 0, stream>>>((const __nv_bfloat16*)x, weight, bias, (float*)out, cp, rows, dim, out_ld, copy_ld, eps);
    else
        return HVS_ERR_UNSUPPORTED;
    count_launch();
    return launch_status();
}

extern "C" size_t hvs_layernorm_bwd_workspace(int64_t rows, int dim) {
    if (rows < 0 || dim <= 0) return 0;
    if (!hvs::ln_small_dim(dim)) {                      // wide rows: partials per 64-row chunk + (mean, rstd) per row
        const int64_t parts = rows > 0 ? (rows + hvs::kWideRowsPerPart - 1) / hvs::kWideRowsPerPart : 1;
        return (size_t)parts * 2 * dim * sizeof(float) + (size_t)(rows > 0 ? rows : 1) * sizeof(float2);
    }
    int64_t blocks = (rows + 63) / 64;
    const int64_t cap = (int64_t)hvs::sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (size_t)blocks * 2 * dim * sizeof(float);
}

extern "C" int hvs_layernorm_bwd(const void* x, int x_dtype, const float* weight, const void* dy, int dy_dtype, void* dx, float* dweight,
                                 float* dbias, int64_t rows, int dim, float eps, void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows < 0 || dim <= 0 || !weight || !dweight || !dbias) return HVS_ERR_BAD_ARG;
    if (dim % 8 || dim > 8192) return HVS_ERR_UNSUPPORTED;
    if (rows > 0 && (!x || !dy || !dx)) return HVS_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) return HVS_ERR_ALIGNMENT;
    const size_t need = hvs_layernorm_bwd_workspace(rows, dim);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15)) return HVS_ERR_WORKSPACE;
    float* part = (float*)workspace;
    typedef __nv_bfloat16 bf;
    if (!ln_small_dim(dim)) {
        const int parts = rows > 0 ? (int)((rows + kWideRowsPerPart - 1) / kWideRowsPerPart) : 0;
        if (parts > 65535) return HVS_ERR_UNSUPPORTED;                  // gridDim.y: rows <= 4.19 M (the model has 6400 at this width)
        float2* stats = reinterpret_cast<float2*>(part + (size_t)(parts > 0 ? parts : 1) * 2 * dim);
        if (rows > 0) {
            if (x_dtype == HVS_DTYPE_F32 && dy_dtype == HVS_DTYPE_F32) launch_wide_bwd((const float*)x, weight, (const float*)dy, (float*)dx, part, stats, parts, rows, dim, eps, stream);
            else if (x_dtype == HVS_DTYPE_BF16 && dy_dtype == HVS_DTYPE_F32) launch_wide_bwd((const bf*)x, weight, (const float*)dy, (bf*)dx, part, stats, parts, rows, dim, eps, stream);
            else if (x_dtype == HVS_DTYPE_F32 && dy_dtype == HVS_DTYPE_BF16) launch_wide_bwd((const float*)x, weight, (const bf*)dy, (float*)dx, part, stats, parts, rows, dim, eps, stream);
            else if (x_dtype == HVS_DTYPE_BF16 && dy_dtype == HVS_DTYPE_BF16) launch_wide_bwd((const bf*)x, weight, (const bf*)dy, (bf*)dx, part, stats, parts, rows, dim, eps, stream);
            else return HVS_ERR_UNSUPPORTED;
        }
        ln_bwd_finalize_kernel<<<(dim + 3) / 4, 128, 0, stream>>>(part, dweight, dbias, parts, dim);
        count_launch(3);
        return launch_status();
    }
    const int blocks = (int)(need / ((size_t)2 * dim * sizeof(float)));
    bool ok = false;
    if (x_dtype == HVS_DTYPE_F32 && dy_dtype == HVS_DTYPE_F32) ok = launch_small_rows_bwd((const float*)x, weight, (const float*)dy, (float*)dx, part, blocks, rows, dim, eps, stream);
    else if (x_dtype == HVS_DTYPE_BF16 && dy_dtype == HVS_DTYPE_F32) ok = launch_small_rows_bwd((const bf*)x, weight, (const float*)dy, (bf*)dx, part, blocks, rows, dim, eps, stream);
    else if (x_dtype == HVS_DTYPE_F32 && dy_dtype == HVS_DTYPE_BF16) ok = launch_small_rows_bwd((const float*)x, weight, (const bf*)dy, (float*)dx, part, blocks, rows, dim, eps, stream);
    else if (x_dtype == HVS_DTYPE_BF16 && dy_dtype == HVS_DTYPE_BF16) ok = launch_small_rows_bwd((const bf*)x, weight, (const bf*)dy, (bf*)dx, part, blocks, rows, dim, eps, stream);
    if (!ok) return HVS_ERR_UNSUPPORTED;
    ln_bwd_finalize_kernel<<<(dim + 3) / 4, 128, 0, stream>>>(part, dweight, dbias, blocks, dim);
    count_launch(2);
    return launch_status();
}

extern "C" size_t hvs_signal_ratio_workspace(int64_t rows) {
    return (size_t)hvs::row_norm_grid(rows < 0 ? 0 : rows) * 2 * sizeof(float);
}

extern "C" int hvs_signal_ratio(const void* out, int out_dtype, const void* x, int x_dtype, int64_t rows, int dim, float* dst,
                                void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows <= 0 || dim <= 0 || !out || !x || !dst) return HVS_ERR_BAD_ARG;
    if (!workspace || workspace_bytes < hvs_signal_ratio_workspace(rows)) return HVS_ERR_WORKSPACE;
    const int grid = row_norm_grid(rows);
    float* part = (float*)workspace;
    typedef __nv_bfloat16 bf;
    if (out_dtype == HVS_DTYPE_F32 && x_dtype == HVS_DTYPE_F32) row_norm_sums_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)out, (const float*)x, rows, dim, part);
    else if (out_dtype == HVS_DTYPE_BF16 && x_dtype == HVS_DTYPE_F32) row_norm_sums_kernel<bf, float><<<grid, 256, 0, stream>>>((const bf*)out, (const float*)x, rows, dim, part);
    else if (out_dtype == HVS_DTYPE_F32 && x_dtype == HVS_DTYPE_BF16) row_norm_sums_kernel<float, bf><<<grid, 256, 0, stream>>>((const float*)out, (const bf*)x, rows, dim, part);
    else if (out_dtype == HVS_DTYPE_BF16 && x_dtype == HVS_DTYPE_BF16) row_norm_sums_kernel<bf, bf><<<grid, 256, 0, stream>>>((const bf*)out, (const bf*)x, rows, dim, part);
    else return HVS_ERR_UNSUPPORTED;
    signal_ratio_final_kernel<<<1, 32, 0, stream>>>(part, grid, rows, dst);
    count_launch(2);
    return launch_status();
}
