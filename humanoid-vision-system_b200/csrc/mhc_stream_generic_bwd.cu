// K1 backward for every stream shape the tuned kernels (mhc_stream_bwd*.cu: n = 4, C = 512) do not take: n in {2, 4}
// residual streams, C a multiple of 8 up to 1024.  Three launches:
//   1. a warp per token: recompute the coefficients (same arithmetic as mhc_stream_generic.cu), G = dy x^T, gate gradients,
//      the Sinkhorn iterations differentiated in the reference's own form P / (sum + eps) -- the normalisers of every
//      iteration are kept, the sweep walks back reconstructing P (P_in = P_out * sum) --, softmax / RMSNorm backward,
//      dx = M^T dy + kappa x + W e (one rounding to bf16), and e = d raw as two bf16 terms (fp32-accurate) into a [T, 64]
//      operand tile;
//   2. dW = x^T [E_hi | E_lo] on the tcgen05 GEMM kernel (k2_gemm.cu: both operands MN-major, split-K with a fixed-order
//      reduction): x is read as the [T, n C] matrix it is;
//   3. finalize: dphi = scale (dW_hi + dW_lo), dscale = sum_k phi dW, dbias / dalpha from the per-CTA partials (fixed order).
// Oracle: autograd through oracle/mhc_ref.py::stream_mhc_forward (reference primitives
// src/models/manifold_layers.py:56-77, :213-216, :449-456).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace hvs {
namespace {

constexpr int kMaxVec = 4;                 // 8-channel groups per lane per stream: C <= 1024
constexpr int kMaxIters = 24;
constexpr int kEw = 64;                    // columns of the E operand tile: [hi 0..K) | 0 .. 32) | [lo 32..32+K) | 0 .. 64)
constexpr int kWarps = 4;

struct GenBwdParams {
    const __nv_bfloat16* x;
    const __nv_bfloat16* dy;
    const float* phi;
    const float* bias;
    const float* alpha;
    const float* scale;
    __nv_bfloat16* dx;
    __nv_bfloat16* e_hl;     // [T, 64]
    float* part;             // [grid, K + 3]
    int64_t T;
    int C, iters;
    float eps_rms, eps_sk;
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoid_acc(float v) { return __fdiv_rn(1.0f, 1.0f + expf(-v)); }
__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
    f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}

template <int N>
__global__ void __launch_bounds__(kWarps * 32) mhc_stream_generic_bwd_kernel(const GenBwdParams p) {
    constexpr int K = N * N + 2 * N;
    __shared__ float sm_hist[kWarps][kMaxIters][2 * N];    // row sums | column sums of every iteration
    __shared__ float sm_e[kWarps][K];
    __shared__ float sm_acc[kWarps][K + 3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * kWarps + wid;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    const int C = p.C, nvec = (C + 255) / 256;
    float bias[K];
#pragma unroll
    for (int l = 0; l < K; ++l) bias[l] = p.bias[l];
    const float a_pre = p.alpha[0], a_post = p.alpha[1], a_res = p.alpha[2];
    float acc_b[K], acc_a[3] = {0.f, 0.f, 0.f};            // dbias / dalpha of this warp's tokens (identical in every lane)
#pragma unroll
    for (int l = 0; l < K; ++l) acc_b[l] = 0.f;

    for (int64_t tok = warp; tok < p.T; tok += nwarps) {
        const __nv_bfloat16* xt = p.x + tok * N * C;
        const __nv_bfloat16* dyt = p.dy + tok * N * C;
        // ---- pass 1: raw = x W (W = bf16(scale * phi)), sum x^2, G = dy x^T
        uint4 xr[N][kMaxVec];
        float raw[K], G[N * N];
#pragma unroll
        for (int l = 0; l < K; ++l) raw[l] = 0.f;
#pragma unroll
        for (int l = 0; l < N * N; ++l) G[l] = 0.f;
        float ss = 0.f;
#pragma unroll
        for (int iv = 0; iv < kMaxVec; ++iv) {
            const int c0 = 8 * (lane + 32 * iv);
            const bool on = iv < nvec && c0 < C;
            float xf[N][8];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                xr[j][iv] = on ? *reinterpret_cast<const uint4*>(xt + j * C + c0) : make_uint4(0u, 0u, 0u, 0u);
                unpack8(xr[j][iv], xf[j]);
            }
            if (!on) continue;
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xv = xf[j][e];
                    ss = fmaf(xv, xv, ss);
                    const int k = j * C + c0 + e;
                    const float sc = __ldg(p.scale + k);
                    const float4* row = reinterpret_cast<const float4*>(p.phi + (size_t)k * K);
#pragma unroll
                    for (int q = 0; q < K / 4; ++q) {
                        const float4 f = __ldg(row + q);
                        raw[4 * q] = fmaf(xv, bf16r(sc * f.x), raw[4 * q]); raw[4 * q + 1] = fmaf(xv, bf16r(sc * f.y), raw[4 * q + 1]);
                        raw[4 * q + 2] = fmaf(xv, bf16r(sc * f.z), raw[4 * q + 2]); raw[4 * q + 3] = fmaf(xv, bf16r(sc * f.w), raw[4 * q + 3]);
                    }
                }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float df[8];
                unpack8(*reinterpret_cast<const uint4*>(dyt + i * C + c0), df);
#pragma unroll
                for (int j = 0; j < N; ++j)
#pragma unroll
                    for (int e = 0; e < 8; ++e) G[i * N + j] = fmaf(df[e], xf[j][e], G[i * N + j]);
            }
        }
        ss = wsum(ss);
#pragma unroll
        for (int l = 0; l < K; ++l) raw[l] = wsum(raw[l]);
#pragma unroll
        for (int l = 0; l < N * N; ++l) G[l] = wsum(G[l]);
        // ---- coefficients, forward (every lane the same; the normalisers of each iteration go to shared memory)
        const float inv_rms = 1.0f / sqrtf(ss / (float)(N * C) + p.eps_rms);
        float hpre[N], hpost[N], P[N * N];
#pragma unroll
        for (int j = 0; j < N; ++j) hpre[j] = sigmoid_acc(fmaf(a_pre, raw[j] * inv_rms, bias[j]));
#pragma unroll
        for (int i = 0; i < N; ++i) hpost[i] = 2.0f * sigmoid_acc(fmaf(a_post, raw[N + i] * inv_rms, bias[N + i]));
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float lg[N], mx = -INFINITY, sum = 0.f;
#pragma unroll
            for (int j = 0; j < N; ++j) { lg[j] = fmaf(a_res, raw[2 * N + i * N + j] * inv_rms, bias[2 * N + i * N + j]); mx = fmaxf(mx, lg[j]); }
#pragma unroll
            for (int j = 0; j < N; ++j) { lg[j] = expf(lg[j] - mx); sum += lg[j]; }
#pragma unroll
            for (int j = 0; j < N; ++j) P[i * N + j] = __fdiv_rn(lg[j], sum) * (float)N;
        }
        __syncwarp();
        for (int it = 0; it < p.iters; ++it) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float rs = 0.f;
#pragma unroll
                for (int j = 0; j < N; ++j) rs += P[i * N + j];
                rs += p.eps_sk;
                if (lane == 0) sm_hist[wid][it][i] = rs;
#pragma unroll
                for (int j = 0; j < N; ++j) P[i * N + j] = __fdiv_rn(P[i * N + j], rs);
            }
#pragma unroll
            for (int j = 0; j < N; ++j) {
                float cs = 0.f;
#pragma unroll
                for (int i = 0; i < N; ++i) cs += P[i * N + j];
                cs += p.eps_sk;
                if (lane == 0) sm_hist[wid][it][N + j] = cs;
#pragma unroll
                for (int i = 0; i < N; ++i) P[i * N + j] = __fdiv_rn(P[i * N + j], cs);
            }
        }
        __syncwarp();
        // M = H_res + H_post (x) H_pre  (y_i = sum_j M_ij x_j)
        float M[N * N];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) M[i * N + j] = fmaf(hpost[i], hpre[j], P[i * N + j]);
        // ---- gate gradients
        float dl[K];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < N; ++i) d = fmaf(hpost[i], G[i * N + j], d);
            dl[j] = d * hpre[j] * (1.0f - hpre[j]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float d = 0.f;
#pragma unroll
            for (int j = 0; j < N; ++j) d = fmaf(hpre[j], G[i * N + j], d);
            dl[N + i] = d * hpost[i] * (1.0f - 0.5f * hpost[i]);
        }
        // ---- reverse sweep through the iterations: dP starts as G (dH_res), P is walked back to the softmax start
        float dP[N * N];
#pragma unroll
        for (int l = 0; l < N * N; ++l) dP[l] = G[l];
        for (int it = p.iters - 1; it >= 0; --it) {
#pragma unroll
            for (int j = 0; j < N; ++j) {                      // P_out = P_in / cs_j
                const float cs = sm_hist[wid][it][N + j];
                float t = 0.f;
#pragma unroll
                for (int i = 0; i < N; ++i) t = fmaf(dP[i * N + j], P[i * N + j], t);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    dP[i * N + j] = __fdiv_rn(dP[i * N + j] - t, cs);
                    P[i * N + j] *= cs;
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {                      // P_out = P_in / rs_i
                const float rs = sm_hist[wid][it][i];
                float t = 0.f;
#pragma unroll
                for (int j = 0; j < N; ++j) t = fmaf(dP[i * N + j], P[i * N + j], t);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    dP[i * N + j] = __fdiv_rn(dP[i * N + j] - t, rs);
                    P[i * N + j] *= rs;
                }
            }
        }
        // softmax * N backward: dl = P0 (dP0 - sum_j(dP0 P0) / N)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < N; ++j) t = fmaf(dP[i * N + j], P[i * N + j], t);
            t *= 1.0f / (float)N;
#pragma unroll
            for (int j = 0; j < N; ++j) dl[2 * N + i * N + j] = P[i * N + j] * (dP[i * N + j] - t);
        }
        // ---- e = d raw, kappa (RMSNorm backward), dbias / dalpha
        float e[K];
        float d_inv = 0.f;
#pragma unroll
        for (int l = 0; l < K; ++l) {
            const float ag = l < N ? a_pre : l < 2 * N ? a_post : a_res;
            e[l] = dl[l] * ag * inv_rms;
            d_inv = fmaf(dl[l] * ag, raw[l], d_inv);
            acc_b[l] += dl[l];
            acc_a[l < N ? 0 : l < 2 * N ? 1 : 2] = fmaf(dl[l] * raw[l], inv_rms, acc_a[l < N ? 0 : l < 2 * N ? 1 : 2]);
        }
        const float kappa = -d_inv * inv_rms * inv_rms * inv_rms / (float)(N * C);
        // E operand tile of the dW GEMM: hi | lo bf16 terms of e
        if (lane == 0) {
#pragma unroll
            for (int l = 0; l < K; ++l) sm_e[wid][l] = e[l];
        }
        __syncwarp();
        {
            float v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = 2 * lane + h;
                const int k = col < 32 ? col : col - 32;
                float val = 0.f;
                if (k < K) {
                    const float ev = sm_e[wid][k];
                    const float hi = bf16r(ev);
                    val = col < 32 ? hi : ev - hi;
                }
                v[h] = val;
            }
            reinterpret_cast<uint32_t*>(p.e_hl + tok * kEw)[lane] = pack_bf16(v[0], v[1]);
        }
        // ---- pass 2: dx = M^T dy + kappa x + W e
#pragma unroll
        for (int iv = 0; iv < kMaxVec; ++iv) {
            const int c0 = 8 * (lane + 32 * iv);
            if (iv >= nvec || c0 >= C) continue;
            float df[N][8];
#pragma unroll
            for (int i = 0; i < N; ++i) unpack8(*reinterpret_cast<const uint4*>(dyt + i * C + c0), df[i]);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                float xf[8], o[8];
                unpack8(xr[j][iv], xf);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float a = kappa * xf[q];
#pragma unroll
                    for (int i = 0; i < N; ++i) a = fmaf(M[i * N + j], df[i][q], a);
                    const int k = j * C + c0 + q;
                    const float sc = __ldg(p.scale + k);
                    const float4* row = reinterpret_cast<const float4*>(p.phi + (size_t)k * K);
#pragma unroll
                    for (int r = 0; r < K / 4; ++r) {
                        const float4 f = __ldg(row + r);
                        a = fmaf(bf16r(sc * f.x), e[4 * r], a); a = fmaf(bf16r(sc * f.y), e[4 * r + 1], a);
                        a = fmaf(bf16r(sc * f.z), e[4 * r + 2], a); a = fmaf(bf16r(sc * f.w), e[4 * r + 3], a);
                    }
                    o[q] = a;
                }
                *reinterpret_cast<uint4*>(p.dx + (tok * N + j) * C + c0) =
                    make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
            }
        }
        __syncwarp();                                          // sm_e / sm_hist are reused by the next token
    }
    // ---- dbias / dalpha of the CTA: the four warps in a fixed order
    if (lane == 0) {
#pragma unroll
        for (int l = 0; l < K; ++l) sm_acc[wid][l] = acc_b[l];
#pragma unroll
        for (int g = 0; g < 3; ++g) sm_acc[wid][K + g] = acc_a[g];
    }
    __syncthreads();
    if (threadIdx.x < K + 3) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += sm_acc[w][threadIdx.x];
        p.part[(size_t)blockIdx.x * (K + 3) + threadIdx.x] = t;
    }
}

// dphi = scale (dW_hi + dW_lo);  dscale = sum_k phi dW;  dbias / dalpha = sum of the CTA partials (warp per output, fixed order)
__global__ void __launch_bounds__(256) generic_bwd_finalize_kernel(const float* __restrict__ dwhl, const float* __restrict__ part, int nparts,
                                                                   const float* __restrict__ phi, const float* __restrict__ scale,
                                                                   float* __restrict__ dphi, float* __restrict__ dscale, float* __restrict__ dbias,
                                                                   float* __restrict__ dalpha, int rows, int K) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w < rows) {
        float dw = 0.f, ds = 0.f;
        if (lane < K) {
            dw = dwhl[(size_t)w * kEw + lane] + dwhl[(size_t)w * kEw + 32 + lane];
            dphi[(size_t)w * K + lane] = scale[w] * dw;
            ds = phi[(size_t)w * K + lane] * dw;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
        if (lane == 0) dscale[w] = ds;
    } else if (w < rows + K + 3) {
        const int c = w - rows;
        float a = 0.f;
        for (int i = lane; i < nparts; i += 32) a += part[(size_t)i * (K + 3) + c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            if (c < K) dbias[c] = a;
            else dalpha[c - K] = a;
        }
    }
}

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
inline int generic_grid(int64_t T) {
    int64_t blocks = (T + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}
struct GenWs { size_t off_e, off_dw, off_part, off_split, total; int splits; };
GenWs gen_carve(int64_t T, int n, int C) {
    GenWs w{};
    const int K = n * n + 2 * n;
    const int64_t rows = (int64_t)n * C;
    w.splits = T > 0 ? hvs_gemm_choose_split(rows, kEw, T) : 1;
    size_t off = 0;
    w.off_e = off; off += up256((size_t)(T > 0 ? T : 1) * kEw * 2);
    w.off_dw = off; off += up256((size_t)rows * kEw * 4);
    w.off_part = off; off += up256((size_t)generic_grid(T) * (K + 3) * 4);
    w.off_split = off; off += w.splits > 1 ? up256((size_t)w.splits * rows * kEw * 4) : 0;
    w.total = off;
    return w;
}

}  // namespace

bool generic_stream_shape_ok(int n, int C);

size_t generic_stream_bwd_workspace(int64_t T, int n, int C) {
    if (T < 0 || !generic_stream_shape_ok(n, C)) return 0;
    return gen_carve(T, n, C).total;
}

int launch_generic_stream_bwd(const void* x, const void* dy, const float* phi, const float* bias, const float* alpha, const float* scale,
                              void* dx, float* dphi, float* dbias, float* dalpha, float* dscale, int64_t T, int n, int C, int iters,
                              float eps_rms, float eps_sk, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (iters < 0 || iters > kMaxIters) return HVS_ERR_UNSUPPORTED;
    const GenWs w = gen_carve(T, n, C);
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return HVS_ERR_ALIGNMENT;
    if (workspace_bytes < w.total) return HVS_ERR_WORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    const int K = n * n + 2 * n;
    const int rows = n * C;
    __nv_bfloat16* e_hl = reinterpret_cast<__nv_bfloat16*>(ws + w.off_e);
    float* dwhl = reinterpret_cast<float*>(ws + w.off_dw);
    float* part = reinterpret_cast<float*>(ws + w.off_part);
    int grid = 0;
    if (T > 0) {
        GenBwdParams p{reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), phi, bias, alpha, scale,
                       reinterpret_cast<__nv_bfloat16*>(dx), e_hl, part, T, C, iters, eps_rms, eps_sk};
        grid = generic_grid(T);
        if (n == 2) mhc_stream_generic_bwd_kernel<2><<<grid, kWarps * 32, 0, stream>>>(p);
        else mhc_stream_generic_bwd_kernel<4><<<grid, kWarps * 32, 0, stream>>>(p);
        count_launch();
        int rc = launch_status();
        if (rc) return rc;
        // dW[(j, c)][hi | lo] = x^T E over the tokens: x is the [T, n C] matrix it is (MN-major A), E [T, 64] (MN-major B)
        hvs_gemm_args g{};
        g.a0 = x; g.lda0 = rows; g.b0 = e_hl; g.ldb0 = kEw; g.K0 = (int)T;
        g.a_mn_major = 1; g.b_mn_major = 1;
        g.out = w.splits > 1 ? (void*)(ws + w.off_split) : (void*)dwhl;
        g.out_dtype = HVS_DTYPE_F32; g.ldo = kEw; g.M = rows; g.N = kEw; g.epilogue = HVS_GEMM_EPI_NONE;
        g.split_k = w.splits; g.split_stride = (int64_t)rows * kEw;
        rc = hvs_gemm_bf16_ex(&g, stream);
        if (rc) return rc;
        if (w.splits > 1) {
            rc = hvs_reduce_partials(reinterpret_cast<const float*>(ws + w.off_split), w.splits, (int64_t)rows * kEw, (int64_t)rows * kEw, dwhl, stream);
            if (rc) return rc;
        }
    } else {
        HVS_CUDA_TRY(cudaMemsetAsync(dwhl, 0, (size_t)rows * kEw * 4, stream));
    }
    generic_bwd_finalize_kernel<<<(rows + K + 3 + 7) / 8, 256, 0, stream>>>(dwhl, part, grid, phi, scale, dphi, dscale, dbias, dalpha, rows, K);
    count_launch();
    return launch_status();
}

}  // namespace hvs
