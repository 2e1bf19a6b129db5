// YOLODecoder.forward (src/models/yolo_head.py:220-294) for one scale, reading the prediction
// head's output in place through element strides.
//   box centre  (gx + sigmoid(tx)) / W, (gy + sigmoid(ty)) / H            (:262-263, repair R7)
//   box size    anchor_w * exp(tw), anchor_h * exp(th), anchors / 416      (:269-270, :50-51)
//   corners     c -/+ size / 2                                             (:273-276)
//   scores      sigmoid(obj) * sigmoid(cls), max / first argmax over classes (:282-285)
// HBM-bound streaming kernels: (5+C) values in, 7 values out per cell.  Mappings:
//   * channel-strided input (the permuted NCHW conv output), unit W stride, W % 4 == 0: four adjacent cells per thread,
//     every channel plane read with 8 / 16-byte loads -- decode_four_cells_first_max when the [cells, C] score tensor is
//     not wanted (the class search runs on the logits, one sigmoid per cell; hvs_yolo_decode_scales runs all scales of
//     the head in one launch), decode_four_cells_per_thread when it is;
//   * any other strided input: a thread per cell, lanes along W;
//   * channel-contiguous input: a warp per cell, lanes along the channel axis.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace hvs {
namespace {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ float sigmoidf_rn(float v) { return __fdiv_rn(1.0f, 1.0f + expf(-v)); }

struct DecodeParams {
    const void* pred;
    int64_t s[5];
    const float* anchor_wh;
    float* boxes;
    float* class_scores;
    int64_t* class_idx;
    float* objectness;
    float* scores;
    int B, A, H, W, C;
};

__device__ __forceinline__ void write_box(const DecodeParams& p, int64_t cell, int a, int h, int w, float tx, float ty,
                                          float tw, float th) {
    const float bx = __fdiv_rn((float)w + sigmoidf_rn(tx), (float)p.W);
    const float by = __fdiv_rn((float)h + sigmoidf_rn(ty), (float)p.H);
    const float bw = __ldg(p.anchor_wh + 2 * a) * expf(tw);
    const float bh = __ldg(p.anchor_wh + 2 * a + 1) * expf(th);
    const float hw = bw * 0.5f, hh = bh * 0.5f;      // x / 2 == x * 0.5 exactly
    reinterpret_cast<float4*>(p.boxes)[cell] = make_float4(bx - hw, by - hh, bx + hw, by + hh);
}

template <typename T>
__global__ void __launch_bounds__(256) decode_cell_per_thread(const DecodeParams p) {
    const int64_t ncell = (int64_t)p.B * p.A * p.H * p.W;
    const T* base = reinterpret_cast<const T*>(p.pred);
    for (int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < ncell;
         cell += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(cell % p.W);
        int64_t r = cell / p.W;
        const int h = (int)(r % p.H); r /= p.H;
        const int a = (int)(r % p.A);
        const int b = (int)(r / p.A);
        const T* q = base + b * p.s[0] + a * p.s[1] + h * p.s[2] + w * p.s[3];
        const float tx = to_f32(q[0]), ty = to_f32(q[p.s[4]]), tw = to_f32(q[2 * p.s[4]]), th = to_f32(q[3 * p.s[4]]);
        write_box(p, cell, a, h, w, tx, ty, tw, th);
        const float obj = sigmoidf_rn(to_f32(q[4 * p.s[4]]));
        float best = -INFINITY;
        int besti = 0;
        for (int c = 0; c < p.C; ++c) {
            const float sc = obj * sigmoidf_rn(to_f32(q[(5 + c) * p.s[4]]));
            if (p.scores != nullptr) p.scores[cell * p.C + c] = sc;
            if (c == 0 || sc > best) { best = sc; besti = c; }
        }
        p.class_scores[cell] = best;
        p.class_idx[cell] = besti;
        if (p.objectness != nullptr) p.objectness[cell] = obj;
    }
}

// Channel-strided input with unit W stride, W % 4 == 0: a thread takes FOUR adjacent cells of a grid row and reads every
// channel plane with one 8-byte (16-bit types) / 16-byte (fp32) load -- a warp covers 256 / 512 contiguous bytes per load
// and issues a quarter of the load instructions of the one-cell mapping (which left the kernel latency-bound at 13 % of
// the HBM roofline); the class loop is unrolled so eight independent loads are in flight per thread.  Same arithmetic,
// same order, same results as decode_cell_per_thread.
template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<__half> { typedef uint2 type; };
template <> struct Vec4<__nv_bfloat16> { typedef uint2 type; };

template <typename T> __device__ __forceinline__ void unpack4(const typename Vec4<T>::type& v, float (&f)[4]);
template <> __device__ __forceinline__ void unpack4<float>(const float4& v, float (&f)[4]) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
template <> __device__ __forceinline__ void unpack4<__nv_bfloat16>(const uint2& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack4<__half>(const uint2& v, float (&f)[4]) {
    const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
}

template <typename T>
__global__ void __launch_bounds__(256) decode_four_cells_per_thread(const DecodeParams p) {
    typedef typename Vec4<T>::type V;
    const int wq = p.W >> 2;
    const int64_t nquad = (int64_t)p.B * p.A * p.H * wq;
    const T* base = reinterpret_cast<const T*>(p.pred);
    for (int64_t quad = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; quad < nquad; quad += (int64_t)gridDim.x * blockDim.x) {
        const int w0 = (int)(quad % wq) << 2;
        int64_t r = quad / wq;
        const int h = (int)(r % p.H); r /= p.H;
        const int a = (int)(r % p.A);
        const int b = (int)(r / p.A);
        const T* q = base + b * p.s[0] + a * p.s[1] + h * p.s[2] + w0;
        const int64_t cell0 = (((int64_t)b * p.A + a) * p.H + h) * p.W + w0;
        float t[5][4];
#pragma unroll
        for (int k = 0; k < 5; ++k) unpack4<T>(*reinterpret_cast<const V*>(q + k * p.s[4]), t[k]);
        float obj[4], best[4];
        int besti[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            write_box(p, cell0 + j, a, h, w0 + j, t[0][j], t[1][j], t[2][j], t[3][j]);
            obj[j] = sigmoidf_rn(t[4][j]);
            best[j] = -INFINITY;
            besti[j] = 0;
        }
        const T* qc = q + 5 * p.s[4];
        int c = 0;
        for (; c + 8 <= p.C; c += 8) {
            V v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const V*>(qc + (int64_t)(c + u) * p.s[4]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float f[4];
                unpack4<T>(v[u], f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float sc = obj[j] * sigmoidf_rn(f[j]);
                    if (p.scores != nullptr) p.scores[(cell0 + j) * p.C + c + u] = sc;
                    if (c + u == 0 || sc > best[j]) { best[j] = sc; besti[j] = c + u; }
                }
            }
        }
        for (; c < p.C; ++c) {
            float f[4];
            unpack4<T>(*reinterpret_cast<const V*>(qc + (int64_t)c * p.s[4]), f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float sc = obj[j] * sigmoidf_rn(f[j]);
                if (p.scores != nullptr) p.scores[(cell0 + j) * p.C + c] = sc;
                if (c == 0 || sc > best[j]) { best[j] = sc; besti[j] = c; }
            }
        }
        *reinterpret_cast<float4*>(p.class_scores + cell0) = make_float4(best[0], best[1], best[2], best[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) p.class_idx[cell0 + j] = besti[j];
        if (p.objectness != nullptr) *reinterpret_cast<float4*>(p.objectness + cell0) = make_float4(obj[0], obj[1], obj[2], obj[3]);
    }
}

// The same mapping when the caller does not ask for the [cells, C] score tensor (the detection path never does): the
// decoder's outputs then need ONE class score per cell, the first maximum of obj * sigmoid(logit).  decode_four_cells_per_thread
// evaluates expf + an IEEE division for all C classes (30 instructions per logit: the kernel ran at 60 % issue and
// 21 % of the HBM roofline).  sigmoid is increasing, so the winner is found on the logits (max, its first index, the
// largest DIFFERENT logit below it: ~10 instructions per logit) and only the winner's score is evaluated -- with the
// same sigmoidf_rn on the same input, so the score is bit-identical.  The index is identical too whenever no other
// class can tie the winner after rounding:
//     (1 - sigmoid(m)) (m - z) >= 2^-19  =>  sigmoid(z) <= sigmoid(m) (1 - 2^-19)    [d ln sigmoid = (1 - sigmoid) dz, decreasing]
// against 3 ulp of error in each computed sigmoid and 1/2 ulp in each product: strictly ordered scores.  Cells where
// that margin fails (saturated sigmoids: logits > 12), with a NaN anywhere, or with a winner score below 1e-30
// (products rounding into the denormals) take the reference loop of decode_four_cells_per_thread for that cell.
// Equal logits give equal scores and the first index in both forms.
template <typename T>
__device__ __noinline__ void decode_cell_reference_loop(const T* class_logits, int64_t class_stride, int C, float obj,
                                                        float* score_out, int64_t* idx_out) {
    float best = -INFINITY;
    int besti = 0;
    for (int c = 0; c < C; ++c) {
        const float sc = obj * sigmoidf_rn(to_f32(class_logits[(int64_t)c * class_stride]));
        if (c == 0 || sc > best) { best = sc; besti = c; }
    }
    *score_out = best;
    *idx_out = besti;
}

// 16-bit inputs keep the search in the packed domain: two cells per instruction (HMNMX2 / HSET2 + LOP3), no unpacking --
// 3.5 instructions per logit.  m is carried with the NaN-propagating maximum, so a NaN logit marks its own cell.
template <typename T> struct Packed2;
template <> struct Packed2<__nv_bfloat16> { typedef __nv_bfloat162 type; static constexpr uint32_t kNegInf = 0xFF80FF80u; };
template <> struct Packed2<__half> { typedef __half2 type; static constexpr uint32_t kNegInf = 0xFC00FC00u; };
template <> struct Packed2<float> { typedef float type; static constexpr uint32_t kNegInf = 0u; };     // (unused)
template <typename H2> __device__ __forceinline__ H2 as_h2(uint32_t v) { return *reinterpret_cast<H2*>(&v); }
template <typename H2> __device__ __forceinline__ uint32_t as_u32(H2 v) { return *reinterpret_cast<uint32_t*>(&v); }
template <typename H2>
__device__ __forceinline__ void first_max_step2(uint32_t& m, uint32_t& m2, uint32_t& bi, uint32_t v, uint32_t cidx2) {
    const H2 hv = as_h2<H2>(v), hm = as_h2<H2>(m);
    const uint32_t ne = __hne2_mask(hv, hm), gt = __hgt2_mask(hv, hm);
    const uint32_t m2c = as_u32(__hmax2(as_h2<H2>(m2), __hmin2(hm, hv)));
    m2 = (m2c & ne) | (m2 & ~ne);
    bi = (cidx2 & gt) | (bi & ~gt);
    m = as_u32(__hmax2_nan(hm, hv));
}
template <typename T> __device__ __forceinline__ float half_to_f32(uint32_t bits16);
template <> __device__ __forceinline__ float half_to_f32<__nv_bfloat16>(uint32_t b) { return __uint_as_float(b << 16); }
template <> __device__ __forceinline__ float half_to_f32<__half>(uint32_t b) { return __half2float(__ushort_as_half((unsigned short)b)); }
template <> __device__ __forceinline__ float half_to_f32<float>(uint32_t b) { return 0.f; }            // (unused)

template <typename T>
__device__ __forceinline__ void first_max_quad(const DecodeParams& p, int64_t quad) {
    typedef typename Vec4<T>::type V;
    constexpr int kU = 16;                                          // independent loads in flight per thread
    const int wq = p.W >> 2;
    const T* base = reinterpret_cast<const T*>(p.pred);
    {
        const int w0 = (int)(quad % wq) << 2;
        int64_t r = quad / wq;
        const int h = (int)(r % p.H); r /= p.H;
        const int a = (int)(r % p.A);
        const int b = (int)(r / p.A);
        const T* q = base + b * p.s[0] + a * p.s[1] + h * p.s[2] + w0;
        const int64_t cell0 = (((int64_t)b * p.A + a) * p.H + h) * p.W + w0;
        V head[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) head[k] = *reinterpret_cast<const V*>(q + k * p.s[4]);
        const T* qc = q + 5 * p.s[4];
        float m[4], m2[4];
        int bi[4];
        bool bad = false;                                           // fp32 inputs: a NaN logit in any of the four cells
        if constexpr (sizeof(T) == 2) {
            typedef typename Packed2<T>::type H2;
            const V v0 = *reinterpret_cast<const V*>(qc);
            uint32_t pm[2] = {v0.x, v0.y}, pm2[2] = {Packed2<T>::kNegInf, Packed2<T>::kNegInf}, pbi[2] = {0u, 0u};
            // always whole groups of kU independent loads: past the last class the index is clamped, and seeing a class twice
            // changes nothing (it is neither above the maximum nor a new value below it)
            for (int c = 1; c < p.C; c += kU) {
                V v[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) v[u] = *reinterpret_cast<const V*>(qc + (int64_t)min(c + u, p.C - 1) * p.s[4]);
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const uint32_t cidx2 = (uint32_t)min(c + u, p.C - 1) * 0x00010001u;
                    first_max_step2<H2>(pm[0], pm2[0], pbi[0], v[u].x, cidx2);
                    first_max_step2<H2>(pm[1], pm2[1], pbi[1], v[u].y, cidx2);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int sh = 16 * (j & 1);
                m[j] = half_to_f32<T>((pm[j >> 1] >> sh) & 0xffffu);      // NaN if the cell had a NaN logit: fails every test below
                m2[j] = half_to_f32<T>((pm2[j >> 1] >> sh) & 0xffffu);
                bi[j] = (int)((pbi[j >> 1] >> sh) & 0xffffu);
            }
        } else {
            {
                float f[4];
                unpack4<T>(*reinterpret_cast<const V*>(qc), f);
#pragma unroll
                for (int j = 0; j < 4; ++j) { m[j] = f[j]; m2[j] = -INFINITY; bi[j] = 0; bad |= f[j] != f[j]; }
            }
            for (int c = 1; c < p.C; c += kU) {
                V v[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) v[u] = *reinterpret_cast<const V*>(qc + (int64_t)min(c + u, p.C - 1) * p.s[4]);
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    float f[4];
                    unpack4<T>(v[u], f);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float z = f[j];
                        bad |= z != z;
                        m2[j] = z != m[j] ? fmaxf(m2[j], fminf(m[j], z)) : m2[j];
                        bi[j] = z > m[j] ? min(c + u, p.C - 1) : bi[j];
                        m[j] = fmaxf(m[j], z);
                    }
                }
            }
        }
        float t[5][4];
#pragma unroll
        for (int k = 0; k < 5; ++k) unpack4<T>(head[k], t[k]);
        float obj[4], best[4];
        bool slow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            write_box(p, cell0 + j, a, h, w0 + j, t[0][j], t[1][j], t[2][j], t[3][j]);
            obj[j] = sigmoidf_rn(t[4][j]);
            best[j] = obj[j] * sigmoidf_rn(m[j]);
            // (1 - sigmoid(m)) (m - m2) >= 2^-19 with 1 / (1 - sigmoid(m)) = 1 + e^m; every comparison is false for a NaN
            const float margin = 1.9073486328125e-06f * (1.0f + expf(m[j]));
            slow[j] = bad || !(margin <= 3.0e38f) || !(m[j] - m2[j] >= margin) || !(best[j] >= 1e-30f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (slow[j]) {
                decode_cell_reference_loop<T>(qc + j, p.s[4], p.C, obj[j], p.class_scores + cell0 + j, p.class_idx + cell0 + j);
            } else {
                p.class_scores[cell0 + j] = best[j];
                p.class_idx[cell0 + j] = bi[j];
            }
        }
        if (p.objectness != nullptr) *reinterpret_cast<float4*>(p.objectness + cell0) = make_float4(obj[0], obj[1], obj[2], obj[3]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) decode_four_cells_first_max(const __grid_constant__ DecodeParams p) {
    const int64_t nquad = (int64_t)p.B * p.A * p.H * (p.W >> 2);
    for (int64_t quad = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; quad < nquad; quad += (int64_t)gridDim.x * blockDim.x)
        first_max_quad<T>(p, quad);
}

// Every scale of the head in ONE launch (hvs_yolo_decode_scales): the 40x40 and 20x20 grids of a batch are 0.5 and 0.13
// waves of their own -- latency-bound launches of 20 us each next to 50 us for the 80x80 grid; as the tail blocks of one
// grid they overlap with it.  Blocks [first_block[k], first_block[k+1]) belong to scale k.
constexpr int kMaxScales = 4;
struct MultiDecodeParams {
    DecodeParams s[kMaxScales];
    unsigned first_block[kMaxScales + 1];
    int n;
};
template <typename T>
__global__ void __launch_bounds__(256) decode_scales_first_max(const __grid_constant__ MultiDecodeParams mp) {
    int k = 0;
#pragma unroll
    for (int i = 1; i < kMaxScales; ++i) k += (i < mp.n && blockIdx.x >= mp.first_block[i]) ? 1 : 0;
    const DecodeParams& p = mp.s[k];
    const int64_t nquad = (int64_t)p.B * p.A * p.H * (p.W >> 2);
    const int64_t quad = (int64_t)(blockIdx.x - mp.first_block[k]) * blockDim.x + threadIdx.x;
    if (quad < nquad) first_max_quad<T>(p, quad);
}

template <typename T>
__global__ void __launch_bounds__(256) decode_cell_per_warp(const DecodeParams p) {
    const int64_t ncell = (int64_t)p.B * p.A * p.H * p.W;
    const T* base = reinterpret_cast<const T*>(p.pred);
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t cell = wid; cell < ncell; cell += nw) {
        const int w = (int)(cell % p.W);
        int64_t r = cell / p.W;
        const int h = (int)(r % p.H); r /= p.H;
        const int a = (int)(r % p.A);
        const int b = (int)(r / p.A);
        const T* q = base + b * p.s[0] + a * p.s[1] + h * p.s[2] + w * p.s[3];
        // lanes 0..4 hold tx,ty,tw,th,obj
        const float head = lane < 5 ? to_f32(q[lane]) : 0.f;
        const float tx = __shfl_sync(0xffffffffu, head, 0), ty = __shfl_sync(0xffffffffu, head, 1);
        const float tw = __shfl_sync(0xffffffffu, head, 2), th = __shfl_sync(0xffffffffu, head, 3);
        const float obj = sigmoidf_rn(__shfl_sync(0xffffffffu, head, 4));
        float best = -INFINITY;
        int besti = 0x7fffffff;
        for (int c = lane; c < p.C; c += 32) {
            const float sc = obj * sigmoidf_rn(to_f32(q[5 + c]));
            if (p.scores != nullptr) p.scores[cell * p.C + c] = sc;
            if (sc > best || besti == 0x7fffffff) { best = sc; besti = c; }
        }
        // max with first-index tie-break; a NaN score never wins over a number (matches `>`)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (oi != 0x7fffffff && (besti == 0x7fffffff || ob > best || (ob == best && oi < besti))) { best = ob; besti = oi; }
        }
        if (lane == 0) {
            write_box(p, cell, a, h, w, tx, ty, tw, th);
            p.class_scores[cell] = best;
            p.class_idx[cell] = besti;
            if (p.objectness != nullptr) p.objectness[cell] = obj;
        }
    }
}

template <typename T>
bool four_cell_mapping_ok(const DecodeParams& p) {
    const uintptr_t vec_bytes = 4 * sizeof(T);
    return p.s[4] != 1 && p.s[3] == 1 && p.W % 4 == 0 && p.s[0] % 4 == 0 && p.s[1] % 4 == 0 && p.s[2] % 4 == 0 && p.s[4] % 4 == 0 &&
           reinterpret_cast<uintptr_t>(p.pred) % vec_bytes == 0 && reinterpret_cast<uintptr_t>(p.class_scores) % 16 == 0 &&
           (p.objectness == nullptr || reinterpret_cast<uintptr_t>(p.objectness) % 16 == 0);
}

template <typename T>
int launch_decode(const DecodeParams& p, cudaStream_t stream) {
    const int64_t ncell = (int64_t)p.B * p.A * p.H * p.W;
    const int cap = sm_count() * 8;
    const bool vec_ok = four_cell_mapping_ok<T>(p);
    timer_begin(5, stream);
    if (p.s[4] == 1) {
        int64_t blocks = (ncell + 7) / 8;
        if (blocks > cap) blocks = cap;
        decode_cell_per_warp<T><<<(int)blocks, 256, 0, stream>>>(p);
    } else if (vec_ok) {
        int64_t blocks = (ncell / 4 + 255) / 256;
        if (blocks > cap) blocks = cap;
        // (first-maximum kernel: one block per 1024 cells, no grid cap -- with the cap at 8 blocks per SM and 4 resident the
        //  batch-64 80x80 grid ran as two waves plus a 16-block third)
        const int64_t all_blocks = (ncell / 4 + 255) / 256;      // (the kernel strides over the grid if this is ever clipped)
        if (p.scores == nullptr) decode_four_cells_first_max<T><<<(unsigned)(all_blocks < 0x7fffffff ? all_blocks : 0x7fffffff), 256, 0, stream>>>(p);
        else decode_four_cells_per_thread<T><<<(int)blocks, 256, 0, stream>>>(p);
    } else {
        int64_t blocks = (ncell + 255) / 256;
        if (blocks > cap) blocks = cap;
        decode_cell_per_thread<T><<<(int)blocks, 256, 0, stream>>>(p);
    }
    timer_end(5, stream);
    count_launch();
    return launch_status();
}

// all scales in one launch when every one of them takes the first-maximum mapping; otherwise scale by scale
template <typename T>
int launch_decode_scales(const DecodeParams* ps, int n, cudaStream_t stream) {
    bool one_launch = n > 1 && n <= kMaxScales;
    uint64_t blocks = 0;
    for (int k = 0; k < n && one_launch; ++k) {
        one_launch = ps[k].scores == nullptr && four_cell_mapping_ok<T>(ps[k]);
        blocks += (uint64_t)(((int64_t)ps[k].B * ps[k].A * ps[k].H * ps[k].W / 4 + 255) / 256);
    }
    if (!one_launch || blocks >= 0x7fffffffull) {
        for (int k = 0; k < n; ++k) {
            const int rc = launch_decode<T>(ps[k], stream);
            if (rc) return rc;
        }
        return HVS_OK;
    }
    MultiDecodeParams mp;
    mp.n = n;
    unsigned at = 0;
    for (int k = 0; k < kMaxScales; ++k) {
        mp.first_block[k] = at;
        if (k < n) {
            mp.s[k] = ps[k];
            at += (unsigned)(((int64_t)ps[k].B * ps[k].A * ps[k].H * ps[k].W / 4 + 255) / 256);
        } else {
            mp.s[k] = ps[0];
        }
    }
    mp.first_block[kMaxScales] = at;
    timer_begin(5, stream);
    decode_scales_first_max<T><<<at, 256, 0, stream>>>(mp);
    timer_end(5, stream);
    count_launch();
    return launch_status();
}

}  // namespace
}  // namespace hvs

extern "C" int hvs_yolo_decode_scales(const hvs_decode_scale* scales_host, int n_scales, int pred_dtype, int B, int C, void* stream) {
    using namespace hvs;
    if (!scales_host || n_scales <= 0 || n_scales > 16) return HVS_ERR_BAD_ARG;
    if (B < 0 || C <= 0) return HVS_ERR_BAD_ARG;
    DecodeParams ps[16];
    for (int k = 0; k < n_scales; ++k) {
        const hvs_decode_scale& q = scales_host[k];
        if (!q.pred || !q.anchor_wh || !q.boxes || !q.class_scores || !q.class_idx) return HVS_ERR_BAD_ARG;
        if (q.A <= 0 || q.H <= 0 || q.W <= 0) return HVS_ERR_BAD_ARG;
        if (reinterpret_cast<uintptr_t>(q.boxes) & 15) return HVS_ERR_ALIGNMENT;
        DecodeParams& p = ps[k];
        p.pred = q.pred;
        for (int i = 0; i < 5; ++i) p.s[i] = q.pred_stride[i];
        p.anchor_wh = q.anchor_wh; p.boxes = q.boxes; p.class_scores = q.class_scores; p.class_idx = q.class_idx;
        p.objectness = q.objectness; p.scores = nullptr;
        p.B = B; p.A = q.A; p.H = q.H; p.W = q.W; p.C = C;
    }
    if (B == 0) return HVS_OK;
    switch (pred_dtype) {
        case HVS_DTYPE_F32: return launch_decode_scales<float>(ps, n_scales, (cudaStream_t)stream);
        case HVS_DTYPE_F16: return launch_decode_scales<__half>(ps, n_scales, (cudaStream_t)stream);
        case HVS_DTYPE_BF16: return launch_decode_scales<__nv_bfloat16>(ps, n_scales, (cudaStream_t)stream);
        default: return HVS_ERR_UNSUPPORTED;
    }
}

extern "C" int hvs_yolo_decode(const void* pred, int pred_dtype, const int64_t* pred_stride_host, const float* anchor_wh,
                               float* boxes, float* class_scores, int64_t* class_idx, float* objectness, float* scores,
                               int B, int A, int H, int W, int C, void* stream) {
    using namespace hvs;
    if (!pred || !pred_stride_host || !anchor_wh || !boxes || !class_scores || !class_idx) return HVS_ERR_BAD_ARG;
    if (B < 0 || A <= 0 || H <= 0 || W <= 0 || C <= 0) return HVS_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(boxes) & 15) return HVS_ERR_ALIGNMENT;
    if (B == 0) return HVS_OK;
    DecodeParams p;
    p.pred = pred;
    for (int i = 0; i < 5; ++i) p.s[i] = pred_stride_host[i];
    p.anchor_wh = anchor_wh; p.boxes = boxes; p.class_scores = class_scores; p.class_idx = class_idx;
    p.objectness = objectness; p.scores = scores;
    p.B = B; p.A = A; p.H = H; p.W = W; p.C = C;
    switch (pred_dtype) {
        case HVS_DTYPE_F32: return launch_decode<float>(p, (cudaStream_t)stream);
        case HVS_DTYPE_F16: return launch_decode<__half>(p, (cudaStream_t)stream);
        case HVS_DTYPE_BF16: return launch_decode<__nv_bfloat16>(p, (cudaStream_t)stream);
        default: return HVS_ERR_UNSUPPORTED;
    }
}
