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

// ---------------------------------------------------------------------------------------------- squeeze-excite gate
// ConvMHCLayer.channel_attention (src/models/vision_backbone.py:77-83, applied :126-127): AdaptiveAvgPool2d(1) -> 1x1 conv
// C -> C/4 -> activation -> 1x1 conv C/4 -> C -> sigmoid, as ONE launch over the channels-last map.  As torch ops it is a
// reduce kernel, two tiny GEMMs with their bias passes, an activation and a sigmoid: 6-7 launches per layer, 28 layers --
// a fifth of the launches of a batch-1 frame.  Here: every CTA sums a slice of an image's rows per channel (16-byte loads,
// fp32), the slices' partial sums go to the workspace, and the LAST CTA of an image to arrive (one counter per image) adds
// them in slice order and runs the two small matrix-vector products.  Rounding points are the bf16 autocast path's: the
// pooled mean, both 1x1 outputs (+ bias), the activation and the sigmoid are each rounded to bf16; accumulation in fp32.
namespace hvs {
namespace {
constexpr int kSeThreads = 256;
constexpr int kSeMaxC = 2048;

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <int ACT>
__global__ void __launch_bounds__(kSeThreads) se_gate_kernel(const uint4* __restrict__ y, const __nv_bfloat16* __restrict__ w1,
                                                             const __nv_bfloat16* __restrict__ b1, const __nv_bfloat16* __restrict__ w2,
                                                             const __nv_bfloat16* __restrict__ b2, __nv_bfloat16* __restrict__ gate,
                                                             float* __restrict__ part, int* __restrict__ counters,
                                                             int64_t rows_per_image, int c8, int hidden, int rows_per_cta) {
    __shared__ float s_red[kSeThreads][9];                 // (+1: the row lanes of a column group land in different banks)
    __shared__ float s_pool[kSeMaxC];
    __shared__ float s_hid[kSeMaxC / 2];
    __shared__ int s_last;
    const int C = c8 * 8;
    const int b = blockIdx.y, slice = blockIdx.x, slices = gridDim.x;
    const int tid = threadIdx.x;
    // ---- 1. partial sums of rows [r0, r1) of image b: thread = (row lane, column group of 8 channels)
    const int lanes = kSeThreads / c8 > 0 ? kSeThreads / c8 : 1;     // row lanes per column group (c8 <= 256)
    const int cg = tid % c8, rl = tid / c8;
    const int64_t r0 = (int64_t)slice * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < rows_per_image ? r0 + rows_per_cta : rows_per_image;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (rl < lanes) {
        const uint4* base = y + ((int64_t)b * rows_per_image) * c8 + cg;
        int64_t r = r0 + rl;
        for (; r + 7 * lanes < r1; r += 8 * lanes) {       // eight independent loads in flight
            uint4 vs[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) vs[k] = base[(r + k * lanes) * c8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[0] += bf16lo(vs[k].x); acc[1] += bf16hi(vs[k].x); acc[2] += bf16lo(vs[k].y); acc[3] += bf16hi(vs[k].y);
                acc[4] += bf16lo(vs[k].z); acc[5] += bf16hi(vs[k].z); acc[6] += bf16lo(vs[k].w); acc[7] += bf16hi(vs[k].w);
            }
        }
        for (; r < r1; r += lanes) {
            const uint4 v = base[r * c8];
            acc[0] += bf16lo(v.x); acc[1] += bf16hi(v.x); acc[2] += bf16lo(v.y); acc[3] += bf16hi(v.y);
            acc[4] += bf16lo(v.z); acc[5] += bf16hi(v.z); acc[6] += bf16lo(v.w); acc[7] += bf16hi(v.w);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s_red[tid][k] = acc[k];
    __syncthreads();
    float* my_part = part + ((int64_t)b * slices + slice) * C;
    for (int c = tid; c < C; c += kSeThreads) {            // channel c = column group c / 8, element c % 8: add the row lanes in order
        const int g = c >> 3, e = c & 7;
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += s_red[l * c8 + g][e];
        my_part[c] = t;
    }
    // ---- 2. the last CTA of the image to get here finishes the gate
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&counters[b], 1) == slices - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float inv = 1.0f / (float)rows_per_image;
    {
        // the slices' partial sums: up to eight threads per channel, each over every P-th slice (loads independent of the adds, eight
        // in flight), then the P pieces in order -- a fixed order, whatever the arrival order of the CTAs was
        const int P = C >= kSeThreads ? 1 : (kSeThreads / C < 8 ? kSeThreads / C : 8);
        float* piece = &s_red[0][0];                        // (the row-lane sums are dead) [P][C]
        __syncthreads();
        for (int i = tid; i < P * C; i += kSeThreads) {
            const int c = i % C, pi = i / C;
            const float* pp = part + (int64_t)b * slices * C + c;
            float t = 0.f;
#pragma unroll 8
            for (int sl = pi; sl < slices; sl += P) t += __ldcg(pp + (int64_t)sl * C);
            piece[pi * C + c] = t;
        }
        __syncthreads();
        for (int c = tid; c < C; c += kSeThreads) {
            float t = 0.f;
            for (int pi = 0; pi < P; ++pi) t += piece[pi * C + c];
            s_pool[c] = bf16_round(t * inv);              // AdaptiveAvgPool2d(1) output in bf16
        }
    }
    __syncthreads();
    // Both products: a warp per output, EIGHT outputs per pass, lanes along the inputs in 16-byte pieces -- the weights come
    // from L2 once, so what counts is loads in flight (one output at a time with 2-byte loads was 50-120 us at C = 512).
    const int warp = tid >> 5, lane = tid & 31;
    auto dot8 = [](const uint4& v, const float* x) {
        return fmaf(bf16lo(v.x), x[0], fmaf(bf16hi(v.x), x[1], fmaf(bf16lo(v.y), x[2], fmaf(bf16hi(v.y), x[3],
               fmaf(bf16lo(v.z), x[4], fmaf(bf16hi(v.z), x[5], fmaf(bf16lo(v.w), x[6], bf16hi(v.w) * x[7])))))));
    };
    // hidden = act(W1 pool + b1)
    {
        const uint4* wv = reinterpret_cast<const uint4*>(w1);
        for (int j0 = warp * 8; j0 < hidden; j0 += (kSeThreads / 32) * 8) {
            float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int g = lane; g < c8; g += 32) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(wv + (int64_t)min(j0 + u, hidden - 1) * c8 + g);
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] += dot8(v[u], s_pool + g * 8);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t[u] += __shfl_xor_sync(0xffffffffu, t[u], o);
            }
            if (lane < 8 && j0 + lane < hidden) {
                float tv = t[0];
#pragma unroll
                for (int u = 1; u < 8; ++u) tv = lane == u ? t[u] : tv;
                float v = bf16_round(tv + __bfloat162float(b1[j0 + lane]));
                if (ACT == 1) v = __fdividef(v, 1.0f + __expf(-v));
                else if (ACT == 2) v = fmaxf(v, 0.f);
                s_hid[j0 + lane] = bf16_round(v);
            }
        }
    }
    __syncthreads();
    // gate = sigmoid(W2 hidden + b2)
    {
        const uint4* wv = reinterpret_cast<const uint4*>(w2);
        const int h8 = hidden >> 3;
        for (int c0 = warp * 8; c0 < C; c0 += (kSeThreads / 32) * 8) {
            float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int g = lane; g < h8; g += 32) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(wv + (int64_t)(c0 + u) * h8 + g);       // C % 8 == 0: all eight rows exist
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] += dot8(v[u], s_hid + g * 8);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t[u] += __shfl_xor_sync(0xffffffffu, t[u], o);
            }
            if (lane < 8) {
                float tv = t[0];
#pragma unroll
                for (int u = 1; u < 8; ++u) tv = lane == u ? t[u] : tv;
                const float z = bf16_round(tv + __bfloat162float(b2[c0 + lane]));
                gate[(int64_t)b * C + c0 + lane] = __float2bfloat16_rn(__fdividef(1.0f, 1.0f + __expf(-z)));
            }
        }
    }
    if (tid == 0) counters[b] = 0;                         // ready for the next launch (also when this one is a graph replay)
}

struct SePlan { int slices, rows_per_cta; size_t part_bytes, total; };
SePlan se_plan(int64_t images, int64_t rows_per_image, int channels) {
    SePlan p;
    int64_t want = (2 * (int64_t)sm_count() + images - 1) / images;               // about two CTAs per SM over the batch
    const int lanes = kSeThreads / (channels / 8) > 0 ? kSeThreads / (channels / 8) : 1;
    const int64_t most = (rows_per_image + 4 * lanes - 1) / (4 * lanes);           // at least four rows per row lane
    if (want > most) want = most;
    if (want < 1) want = 1;
    p.rows_per_cta = (int)((rows_per_image + want - 1) / want);
    p.slices = (int)((rows_per_image + p.rows_per_cta - 1) / p.rows_per_cta);
    p.part_bytes = (((size_t)images * p.slices * channels * 4) + 255) & ~(size_t)255;
    p.total = p.part_bytes + ((((size_t)images * 4) + 255) & ~(size_t)255);
    return p;
}
}  // namespace
}  // namespace hvs

extern "C" size_t hvs_se_gate_workspace(int64_t images, int64_t rows_per_image, int channels) {
    if (images <= 0 || rows_per_image <= 0 || channels <= 0 || channels % 8 || channels > hvs::kSeMaxC) return 0;
    return hvs::se_plan(images, rows_per_image, channels).total;
}

extern "C" int hvs_se_gate_bf16(const void* y, const void* w1, const void* b1, const void* w2, const void* b2, void* gate,
                                int64_t images, int64_t rows_per_image, int channels, int hidden, int activation,
                                void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (images < 0 || rows_per_image <= 0 || channels <= 0 || hidden <= 0) return HVS_ERR_BAD_ARG;
    if (images == 0) return HVS_OK;
    if (!y || !w1 || !b1 || !w2 || !b2 || !gate || !workspace) return HVS_ERR_BAD_ARG;
    if (channels % 8 || hidden % 8 || channels > kSeMaxC || channels / 8 > kSeThreads || hidden > kSeMaxC / 2 || activation < 0 ||
        activation > 2 || images > 65535)
        return HVS_ERR_UNSUPPORTED;
    if (((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(w2)) & 15) ||
        (reinterpret_cast<uintptr_t>(workspace) & 255))
        return HVS_ERR_ALIGNMENT;
    const SePlan pl = se_plan(images, rows_per_image, channels);
    if (workspace_bytes < pl.total) return HVS_ERR_WORKSPACE;
    float* part = reinterpret_cast<float*>(workspace);
    int* counters = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + pl.part_bytes);
    const dim3 grid((unsigned)pl.slices, (unsigned)images);
    const uint4* yp = reinterpret_cast<const uint4*>(y);
    const __nv_bfloat16 *w1p = reinterpret_cast<const __nv_bfloat16*>(w1), *b1p = reinterpret_cast<const __nv_bfloat16*>(b1);
    const __nv_bfloat16 *w2p = reinterpret_cast<const __nv_bfloat16*>(w2), *b2p = reinterpret_cast<const __nv_bfloat16*>(b2);
    __nv_bfloat16* gp = reinterpret_cast<__nv_bfloat16*>(gate);
    if (activation == 0) se_gate_kernel<0><<<grid, kSeThreads, 0, stream>>>(yp, w1p, b1p, w2p, b2p, gp, part, counters, rows_per_image, channels / 8, hidden, pl.rows_per_cta);
    else if (activation == 1) se_gate_kernel<1><<<grid, kSeThreads, 0, stream>>>(yp, w1p, b1p, w2p, b2p, gp, part, counters, rows_per_image, channels / 8, hidden, pl.rows_per_cta);
    else se_gate_kernel<2><<<grid, kSeThreads, 0, stream>>>(yp, w1p, b1p, w2p, b2p, gp, part, counters, rows_per_image, channels / 8, hidden, pl.rows_per_cta);
    count_launch();
    return launch_status();
}

// ---------------------------------------------------------------------------------------------- folded-BatchNorm bias + activation
// After eval-mode BatchNorm is folded into the convolution (y = conv_w'(x) + b'), ATen adds the bias in a broadcast pass of its
// own and the activation in another; here both are one vectorised pass over the channels-last map:
//   out[t, c] = act(y[t, c] + bias[c]),  act in {identity, SiLU (vision_backbone.py:36-45 default), ReLU (feature_fusion.py
//   refinement stacks), LeakyReLU(0.1) (yolo_head.py:99-112 conv_layers)}.
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
            else if (ACT == 3) v[k] = v[k] > 0.f ? v[k] : 0.1f * v[k];
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
    if (channels % 8 || activation < 0 || activation > 3) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(out)) & 15) return HVS_ERR_ALIGNMENT;
    const int64_t total = rows * (channels / 8);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const uint4* yp = reinterpret_cast<const uint4*>(y);
    uint4* op = reinterpret_cast<uint4*>(out);
    if (activation == 0) bias_act_kernel<0><<<(int)blocks, 256, 0, stream>>>(yp, bias, op, rows, channels / 8);
    else if (activation == 1) bias_act_kernel<1><<<(int)blocks, 256, 0, stream>>>(yp, bias, op, rows, channels / 8);
    else if (activation == 2) bias_act_kernel<2><<<(int)blocks, 256, 0, stream>>>(yp, bias, op, rows, channels / 8);
    else bias_act_kernel<3><<<(int)blocks, 256, 0, stream>>>(yp, bias, op, rows, channels / 8);
    count_launch();
    return launch_status();
}
