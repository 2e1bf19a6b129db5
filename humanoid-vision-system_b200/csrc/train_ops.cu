// Training-side and input-side helpers on the edges of the hot path.
//
// hvs_grad_clip_dual   ManifoldConstrainedTrainer._apply_manifold_gradient_clipping
//                      (src/training/mhc_trainer.py:342-383): the gradients of the mHC parameters are clipped to
//                      mhc_max_norm (0.5), all others to max_grad_norm (1.0), each by torch's clip_grad_norm_ rule
//                      (coef = max_norm / (norm + 1e-6), applied only when < 1).  The reference does this with two
//                      foreach norms, two .item() host syncs and two foreach multiplies; here it is three launches over
//                      a table of tensors, no host synchronisation, fixed-order (bitwise reproducible) reductions.
// hvs_preprocess_u8    ImagePreprocessor "accurate" path (src/inference/preprocessing.py:252-273): HWC uint8 frame ->
//                      bilinear resize (cv2.INTER_LINEAR sampling: half-pixel centres, edge clamp) -> optional BGR->RGB
//                      (:199-203) -> / 255 -> (x - mean) / std -> CHW tensor, one kernel, for the streaming config.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace hvs {
namespace {

constexpr int kClipThreads = 256;
constexpr int kChunk = 1 << 16;          // elements per work item

struct ClipTable {
    const hvs_grad_tensor* tensors;      // device copy
    const int* chunk_tensor;             // [nchunks] tensor index of each chunk
    const int* chunk_first;              // [ntensors] first chunk of each tensor
    float* partial;                      // [nchunks]
    float* result;                       // [4]: norm group 0, norm group 1, coef 0, coef 1
    int ntensors, nchunks;
    float max_norm0, max_norm1;
};

__device__ __forceinline__ float block_sum(float v, float* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < kClipThreads / 32 ? sm[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;                            // valid in thread 0
}

__global__ void __launch_bounds__(kClipThreads) clip_sumsq_kernel(const ClipTable tb) {
    __shared__ float sm[kClipThreads / 32];
    for (int c = blockIdx.x; c < tb.nchunks; c += gridDim.x) {
        const int t = tb.chunk_tensor[c];
        const hvs_grad_tensor g = tb.tensors[t];
        const int64_t lo = (int64_t)(c - tb.chunk_first[t]) * kChunk;
        const int64_t hi = lo + kChunk < g.numel ? lo + kChunk : g.numel;
        float s = 0.f;
        for (int64_t i = lo + threadIdx.x; i < hi; i += kClipThreads) { const float v = g.grad[i]; s = fmaf(v, v, s); }
        s = block_sum(s, sm);
        if (threadIdx.x == 0) tb.partial[c] = s;
    }
}

__global__ void __launch_bounds__(kClipThreads) clip_finalize_kernel(const ClipTable tb) {
    __shared__ float sm[kClipThreads / 32];
    float s0 = 0.f, s1 = 0.f;
    for (int c = threadIdx.x; c < tb.nchunks; c += kClipThreads) {      // fixed assignment, fixed order
        const float v = tb.partial[c];
        if (tb.tensors[tb.chunk_tensor[c]].group == 0) s0 += v; else s1 += v;
    }
    s0 = block_sum(s0, sm);
    s1 = block_sum(s1, sm);
    if (threadIdx.x == 0) {
        const float n0 = sqrtf(s0), n1 = sqrtf(s1);
        const float c0 = tb.max_norm0 / (n0 + 1e-6f), c1 = tb.max_norm1 / (n1 + 1e-6f);
        tb.result[0] = n0; tb.result[1] = n1;
        tb.result[2] = c0 < 1.f ? c0 : 1.f;
        tb.result[3] = c1 < 1.f ? c1 : 1.f;
    }
}

__global__ void __launch_bounds__(kClipThreads) clip_scale_kernel(const ClipTable tb) {
    const float c0 = tb.result[2], c1 = tb.result[3];
    if (c0 == 1.f && c1 == 1.f) return;
    for (int c = blockIdx.x; c < tb.nchunks; c += gridDim.x) {
        const int t = tb.chunk_tensor[c];
        const hvs_grad_tensor g = tb.tensors[t];
        const float k = g.group == 0 ? c0 : c1;
        if (k == 1.f) continue;
        const int64_t lo = (int64_t)(c - tb.chunk_first[t]) * kChunk;
        const int64_t hi = lo + kChunk < g.numel ? lo + kChunk : g.numel;
        for (int64_t i = lo + threadIdx.x; i < hi; i += kClipThreads) g.grad[i] *= k;
    }
}

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

struct ClipLayout { size_t off_t, off_ct, off_cf, off_part, table_bytes, total; int nchunks; };

ClipLayout clip_layout(const hvs_grad_tensor* t, int n) {
    ClipLayout L{};
    int64_t chunks = 0;
    for (int i = 0; i < n; ++i) chunks += (t[i].numel + kChunk - 1) / kChunk;
    L.nchunks = (int)chunks;
    size_t off = 0;
    L.off_t = off; off += up256((size_t)n * sizeof(hvs_grad_tensor));
    L.off_ct = off; off += up256((size_t)chunks * 4);
    L.off_cf = off; off += up256((size_t)n * 4);
    L.table_bytes = off;
    L.off_part = off; off += up256((size_t)chunks * 4);
    L.total = off;
    return L;
}

// ---------------------------------------------------------------------------- preprocessing
template <typename TO> __device__ __forceinline__ TO cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct PreParams {
    const uint8_t* src; int sh, sw, sc; int64_t src_pitch;       // HWC, channels 1 or 3
    int dh, dw; int swap_rb; float mean[3], inv_std[3]; float scale;
};

template <typename TO>
__global__ void preprocess_kernel(const PreParams p, TO* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= p.dw) return;
    const float fy = ((float)y + 0.5f) * ((float)p.sh / (float)p.dh) - 0.5f;
    const float fx = ((float)x + 0.5f) * ((float)p.sw / (float)p.dw) - 0.5f;
    int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
    const float wy = fy - (float)y0, wx = fx - (float)x0;
    const int y1 = min(max(y0 + 1, 0), p.sh - 1), x1 = min(max(x0 + 1, 0), p.sw - 1);
    y0 = min(max(y0, 0), p.sh - 1); x0 = min(max(x0, 0), p.sw - 1);
    const uint8_t* r0 = p.src + (int64_t)y0 * p.src_pitch;
    const uint8_t* r1 = p.src + (int64_t)y1 * p.src_pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int sc = p.sc == 1 ? 0 : (p.swap_rb ? 2 - c : c);
        const float a = r0[x0 * p.sc + sc], b = r0[x1 * p.sc + sc], d = r1[x0 * p.sc + sc], e = r1[x1 * p.sc + sc];
        const float top = a + (b - a) * wx, bot = d + (e - d) * wx;
        const float v = (top + (bot - top) * wy) * p.scale;
        dst[((int64_t)c * p.dh + y) * p.dw + x] = cvt_out<TO>((v - p.mean[c]) * p.inv_std[c]);
    }
}

}  // namespace
}  // namespace hvs

extern "C" size_t hvs_grad_clip_dual_workspace(const hvs_grad_tensor* tensors_host, int num_tensors) {
    if (!tensors_host || num_tensors <= 0) return 256;
    return hvs::clip_layout(tensors_host, num_tensors).total;
}

extern "C" int hvs_grad_clip_dual(const hvs_grad_tensor* tensors_host, int num_tensors, float max_norm_group0,
                                  float max_norm_group1, float* result4, void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_tensors < 0 || !result4 || (num_tensors > 0 && !tensors_host)) return HVS_ERR_BAD_ARG;
    if (num_tensors == 0) return (int)cudaMemsetAsync(result4, 0, 16, stream);
    for (int i = 0; i < num_tensors; ++i)
        if (!tensors_host[i].grad || tensors_host[i].numel < 0 || (tensors_host[i].group != 0 && tensors_host[i].group != 1)) return HVS_ERR_BAD_ARG;
    const ClipLayout L = clip_layout(tensors_host, num_tensors);
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return HVS_ERR_ALIGNMENT;
    if (workspace_bytes < L.total) return HVS_ERR_WORKSPACE;
    std::vector<uint8_t> tab(L.table_bytes, 0);
    memcpy(tab.data() + L.off_t, tensors_host, (size_t)num_tensors * sizeof(hvs_grad_tensor));
    int* ct = reinterpret_cast<int*>(tab.data() + L.off_ct);
    int* cf = reinterpret_cast<int*>(tab.data() + L.off_cf);
    int c = 0;
    for (int i = 0; i < num_tensors; ++i) {
        cf[i] = c;
        const int n = (int)((tensors_host[i].numel + kChunk - 1) / kChunk);
        for (int k = 0; k < n; ++k) ct[c++] = i;
    }
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    {
        const int rcu = upload_table(ws, std::move(tab), stream);
        if (rcu) return rcu;
    }
    ClipTable tb{};
    tb.tensors = reinterpret_cast<const hvs_grad_tensor*>(ws + L.off_t);
    tb.chunk_tensor = reinterpret_cast<const int*>(ws + L.off_ct);
    tb.chunk_first = reinterpret_cast<const int*>(ws + L.off_cf);
    tb.partial = reinterpret_cast<float*>(ws + L.off_part);
    tb.result = result4;
    tb.ntensors = num_tensors; tb.nchunks = L.nchunks;
    tb.max_norm0 = max_norm_group0; tb.max_norm1 = max_norm_group1;
    if (L.nchunks == 0) return (int)cudaMemsetAsync(result4, 0, 16, stream);
    int grid = sm_count() * 8;
    if (grid > L.nchunks) grid = L.nchunks;
    clip_sumsq_kernel<<<grid, kClipThreads, 0, stream>>>(tb);
    clip_finalize_kernel<<<1, kClipThreads, 0, stream>>>(tb);
    clip_scale_kernel<<<grid, kClipThreads, 0, stream>>>(tb);
    count_launch(3);
    return launch_status();
}

extern "C" int hvs_preprocess_u8(const void* src, int src_h, int src_w, int src_channels, int64_t src_pitch_bytes, void* dst,
                                 int dst_dtype, int dst_h, int dst_w, int swap_rb, const float* mean3_host,
                                 const float* std3_host, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!src || !dst || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return HVS_ERR_BAD_ARG;
    if (src_channels != 1 && src_channels != 3) return HVS_ERR_UNSUPPORTED;
    if (src_pitch_bytes < (int64_t)src_w * src_channels) return HVS_ERR_BAD_ARG;
    PreParams p{};
    p.src = (const uint8_t*)src; p.sh = src_h; p.sw = src_w; p.sc = src_channels; p.src_pitch = src_pitch_bytes;
    p.dh = dst_h; p.dw = dst_w; p.swap_rb = swap_rb; p.scale = 1.0f / 255.0f;
    for (int c = 0; c < 3; ++c) {
        p.mean[c] = mean3_host ? mean3_host[c] : 0.f;
        p.inv_std[c] = std3_host ? 1.0f / std3_host[c] : 1.f;
    }
    dim3 grid((dst_w + 127) / 128, dst_h);
    if (dst_dtype == HVS_DTYPE_F32) preprocess_kernel<float><<<grid, 128, 0, stream>>>(p, (float*)dst);
    else if (dst_dtype == HVS_DTYPE_F16) preprocess_kernel<__half><<<grid, 128, 0, stream>>>(p, (__half*)dst);
    else if (dst_dtype == HVS_DTYPE_BF16) preprocess_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(p, (__nv_bfloat16*)dst);
    else return HVS_ERR_UNSUPPORTED;
    count_launch();
    return launch_status();
}

// ---------------------------------------------------------------------------------------------- squeeze-excite gate + residual
// The two elementwise passes that follow every mHC hop of the backbone's ConvMHCLayer (src/models/vision_backbone.py:125-133:
// x = x * channel_attention(x), then + identity) as one pass over the channels-last token view: out[t, c] = y[t, c] * g[b(t), c] (+ r[t, c]).
namespace hvs {
namespace {
__global__ void __launch_bounds__(256) gate_residual_kernel(const uint4* __restrict__ y, const __nv_bfloat16* __restrict__ gate,
                                                            const uint4* __restrict__ res, uint4* __restrict__ out, int64_t rows,
                                                            int64_t rows_per_image, int c8) {
    const int64_t total = rows * c8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / c8;
        const int g8 = (int)(i - row * c8);
        const int64_t b = row / rows_per_image;
        const uint4 yv = y[i];
        const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gate + (b * c8 + g8) * 8));
        uint4 rv = make_uint4(0u, 0u, 0u, 0u);
        if (res != nullptr) rv = res[i];
        const uint32_t ya[4] = {yv.x, yv.y, yv.z, yv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w}, ra[4] = {rv.x, rv.y, rv.z, rv.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = pack_bf16(fmaf(bf16lo(ya[k]), bf16lo(ga[k]), bf16lo(ra[k])), fmaf(bf16hi(ya[k]), bf16hi(ga[k]), bf16hi(ra[k])));
        out[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
}  // namespace
}  // namespace hvs

extern "C" int hvs_gate_residual_bf16(const void* y, const void* gate, const void* residual, void* out, int64_t images,
                                      int64_t rows_per_image, int channels, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (images < 0 || rows_per_image <= 0 || channels <= 0) return HVS_ERR_BAD_ARG;
    if (images == 0) return HVS_OK;
    if (!y || !gate || !out) return HVS_ERR_BAD_ARG;
    if (channels % 8) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gate) | reinterpret_cast<uintptr_t>(residual) |
         reinterpret_cast<uintptr_t>(out)) & 15)
        return HVS_ERR_ALIGNMENT;
    const int64_t rows = images * rows_per_image, total = rows * (channels / 8);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    gate_residual_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(y), reinterpret_cast<const __nv_bfloat16*>(gate),
                                                          reinterpret_cast<const uint4*>(residual), reinterpret_cast<uint4*>(out), rows,
                                                          rows_per_image, channels / 8);
    count_launch();
    return launch_status();
}

// ---------------------------------------------------------------------------------------------- folded-BatchNorm bias + activation
// After eval-mode BatchNorm is folded into the convolution (y = conv_w'(x) + b'), ATen adds the bias in a broadcast pass of its
// own and the activation in another; here both are one vectorised pass over the channels-last map:
//   out[t, c] = act(y[t, c] + bias[c]),  act in {identity, SiLU (vision_backbone.py:36-45 default), ReLU}.
namespace hvs {
namespace {
template <int ACT>
__global__ void __launch_bounds__(256) bias_act_kernel(const uint4* __restrict__ y, const float* __restrict__ bias, uint4* __restrict__ out,
                                                       int64_t rows, int c8) {
    const int64_t total = rows * c8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int g8 = (int)(i % c8);
        const uint4 yv = y[i];
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * g8), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * g8 + 1);
        float v[8] = {bf16lo(yv.x) + b0.x, bf16hi(yv.x) + b0.y, bf16lo(yv.y) + b0.z, bf16hi(yv.y) + b0.w,
                      bf16lo(yv.z) + b1.x, bf16hi(yv.z) + b1.y, bf16lo(yv.w) + b1.z, bf16hi(yv.w) + b1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (ACT == 1) v[k] = __fdividef(v[k], 1.0f + __expf(-v[k]));
            else if (ACT == 2) v[k] = fmaxf(v[k], 0.f);
        }
        out[i] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}
}  // namespace
}  // namespace hvs

extern "C" int hvs_bias_act_bf16(const void* y, const float* bias, void* out, int64_t rows, int channels, int activation, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows < 0 || channels <= 0) return HVS_ERR_BAD_ARG;
    if (rows == 0) return HVS_OK;
    if (!y || !bias || !out) return HVS_ERR_BAD_ARG;
    if (channels % 8 || activation < 0 || activation > 2) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(out)) & 15) return HVS_ERR_ALIGNMENT;
    const int64_t total = rows * (channels / 8);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const uint4* yp = reinterpret_cast<const uint4*>(y);
    uint4* op = reinterpret_cast<uint4*>(out);
    if (activation == 0) bias_act_kernel<0><<<(int)blocks, 256, 0, stream>>>(yp, bias, op, rows, channels / 8);
    else if (activation == 1) bias_act_kernel<1><<<(int)blocks, 256, 0, stream>>>(yp, bias, op, rows, channels / 8);
    else bias_act_kernel<2><<<(int)blocks, 256, 0, stream>>>(yp, bias, op, rows, channels / 8);
    count_launch();
    return launch_status();
}
