// Token-path GEMMs of the reference-literal module (K2), ManifoldHyperConnection.forward
// (src/models/manifold_layers.py:248-270):
//     z  = LN_pre(x) @ H_pre                                       EPI_NONE       (bf16 out)
//     z  = GELU(z @ W1^T + b1),  z = GELU(z @ W2^T + b2)           EPI_BIAS_GELU  (bf16 out)
//     y  = LN_post(z @ H_post + x @ H_res)                         EPI_LAYERNORM  (two operand pairs, one accumulator)
// One persistent warp-specialised kernel for sm_100a: a TMA producer thread, a tcgen05.mma issuer thread, eight
// epilogue warps.  D[128 x BN] fp32 accumulators live in tensor memory, double buffered (2 x 256 columns), so the
// epilogue of tile i (tcgen05.ld -> bias / GELU / LayerNorm in registers -> global) overlaps the MMAs of tile i+1.
// Operands: A [M, K] bf16 row-major (tokens x features), B [N, K] bf16 row-major (= nn.Linear's weight layout; the
// static coefficient kernel emits H_pre^T / H_post^T / H_res^T in it), both staged by TMA as 128-byte-swizzled
// K-major tiles of 64 K-elements, 4-6 stages.
// LayerNorm needs the whole output row: with N <= 256 the tile spans it; with 256 < N <= 512 the two halves of the
// row go to the two accumulator buffers and the epilogue normalises across both (no overlap for that GEMM).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "umma_sm100.cuh"

namespace hvs {
namespace {

constexpr int kBM = 128;                 // tokens per tile (UMMA M)
constexpr int kBK = 64;                  // K elements per stage (128 bytes: one swizzle row)
constexpr int kMaxBN = 256;              // UMMA N limit
constexpr int kMaxStages = 8;
constexpr int kABytes = kBM * kBK * 2;   // 16 KB
constexpr int kStageBudget = 196608;     // shared memory for the operand ring
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kSmemBytes = kStageBudget + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int kTmemCols = 512;

struct GemmParams {
    const float* bias;      // [N] or null
    const float* ln_w;      // [N]
    const float* ln_b;      // [N]
    void* out;
    int64_t ldo;            // elements
    int64_t M;
    int N;
    int BN;                 // columns per accumulator (divides N, multiple of 32, <= 256)
    int n_sub;              // accumulators per tile (1, or N / BN <= 2 in LayerNorm mode)
    int kb0, kb1;           // 64-element K blocks from operand pair 0 / pair 1
    int stages;
    int epilogue;
    int out_f32;
    float ln_eps;
    int num_tiles;          // m_tiles * n_outer
    int n_outer;            // N / (BN * n_sub)
    // HVS_GEMM_EPI_YOLO_DECODE (fused prediction conv + decode): outputs in hvs_yolo_decode's layout
    const float* anchor_wh; // [3, 2]
    float* dec_boxes;       // [B, 3, H, W, 4]
    float* dec_scores;      // [B, 3, H, W]
    int64_t* dec_cls;       // [B, 3, H, W]
    float* dec_obj;         // [B, 3, H, W] or null
    int grid_h, grid_w;
};

constexpr int kEpiYoloDecode = 100;      // internal epilogue id (entered through hvs_head_decode_fused)
__device__ __forceinline__ float sigmoidf_rn(float v) { return __fdiv_rn(1.0f, 1.0f + expf(-v)); }

// ---- packed fp32x2 arithmetic (one instruction for two accumulator columns): the GELU epilogue is ISSUE-bound -- at
// K <= 256 a tile's 32768 activations cost more issue slots than its MMAs take cycles -- so instructions per element
// are what counts
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)), the erf form of nn.GELU() (manifold_layers.py:165,168), for two values at once.
// erf by Abramowitz & Stegun 7.1.28: 1 - (1 + a1 z + ... + a6 z^6)^-16 on |z|, sign restored; absolute error of the
// result <= 6e-7 for |x| <= 8 (measured against fp64), i.e. 3 decimal orders below the bf16 rounding of the output.
// ~12 issue slots per element against ~40 for erff().
__device__ __forceinline__ uint32_t gelu2_bf16(float x0, float x1) {
    const u64 x = pk2(x0, x1);
    const u64 z = mul2(x, pk2(0.70710678118654752440f, 0.70710678118654752440f));
    float z0, z1;
    upk2(z, z0, z1);
    const u64 az = pk2(fabsf(z0), fabsf(z1));
    u64 t = fma2(az, pk2(0.0000430638f, 0.0000430638f), pk2(0.0002765672f, 0.0002765672f));
    t = fma2(az, t, pk2(0.0001520143f, 0.0001520143f));
    t = fma2(az, t, pk2(0.0092705272f, 0.0092705272f));
    t = fma2(az, t, pk2(0.0422820123f, 0.0422820123f));
    t = fma2(az, t, pk2(0.0705230784f, 0.0705230784f));
    t = fma2(az, t, pk2(1.0f, 1.0f));
    t = mul2(t, t); t = mul2(t, t); t = mul2(t, t); t = mul2(t, t);
    float t0, t1;
    upk2(t, t0, t1);
    const float e0 = copysignf(1.0f - rcp_approx(t0), z0), e1 = copysignf(1.0f - rcp_approx(t1), z1);
    const u64 hx = mul2(x, pk2(0.5f, 0.5f));
    float g0, g1;
    upk2(fma2(hx, pk2(e0, e1), hx), g0, g1);
    return pack_bf16(g0, g1);
}
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
k2_gemm_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_b0,
               const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_b1, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStageBudget);
    uint64_t* full = bars;                       // [kMaxStages]  TMA -> MMA
    uint64_t* empty = bars + kMaxStages;         // [kMaxStages]  MMA -> TMA
    uint64_t* acc_full = bars + 2 * kMaxStages;  // [2]           MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;          // [2]           epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stage_bytes = kABytes + p.BN * 128;
    const int num_kb = p.kb0 + p.kb1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a0);
        tma_prefetch_desc(&tm_b0);
        if (p.kb1 > 0) { tma_prefetch_desc(&tm_a1); tma_prefetch_desc(&tm_b1); }
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (one thread)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int m_blk = tile / p.n_outer, n_out = tile % p.n_outer;
                for (int sub = 0; sub < p.n_sub; ++sub) {
                    const int n_row0 = (n_out * p.n_sub + sub) * p.BN;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
                        if (kb < p.kb0) {
                            tma_load_2d(sa, &tm_a0, &full[stage], kb * kBK, m_blk * kBM);
                            tma_load_2d(sa + kABytes, &tm_b0, &full[stage], kb * kBK, n_row0);
                        } else {
                            tma_load_2d(sa, &tm_a1, &full[stage], (kb - p.kb0) * kBK, m_blk * kBM);
                            tma_load_2d(sa + kABytes, &tm_b1, &full[stage], (kb - p.kb0) * kBK, n_row0);
                        }
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kBM, p.BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t acc_it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int sub = 0; sub < p.n_sub; ++sub, ++acc_it) {
                    const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
                    mbar_wait(&acc_empty[buf], aph ^ 1u);            // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kMaxBN;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = base + (uint32_t)(stage * stage_bytes);
                        const uint64_t adesc = umma_smem_desc(sa, 16, 1024, kUmmaLayoutSw128);
                        const uint64_t bdesc = umma_smem_desc(sa + kABytes, 16, 1024, kUmmaLayoutSw128);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)           // +32 bytes (16 bf16) along K per step
                            umma_bf16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                        umma_commit(&empty[stage]);                  // stage reusable when these MMAs have read it
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit(&acc_full[buf]);                     // accumulator complete
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int q = warp & 3;                                      // tensor-memory lane quadrant this warp may read
        const int half = (warp - 2) >> 2;                            // which alternate 32-column chunks it takes
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t acc_it = 0;
        const int chunks = p.BN >> 5;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.n_outer, n_out = tile % p.n_outer;
            const int64_t row = (int64_t)m_blk * kBM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t it0 = acc_it;
            for (int sub = 0; sub < p.n_sub; ++sub) {
                const uint32_t it = it0 + (uint32_t)sub;
                mbar_wait(&acc_full[it & 1u], (it >> 1) & 1u);
            }
            tc_fence_after();
            if (p.epilogue == kEpiYoloDecode) {
                // ---- YOLODecoder.forward (yolo_head.py:241-285) on the accumulator row: a row is one pixel, its 255 columns are
                // 3 anchors x (tx, ty, tw, th, obj, 80 classes).  Same arithmetic and order as yolo_decode.cu; the raw
                // predictions never reach HBM.  Half 0 of the warp pair takes anchors 0 and 1, half 1 anchor 2.
                const uint32_t t0 = tmem_base + lane_base + (it0 & 1u) * kMaxBN;
                const int hw = p.grid_h * p.grid_w;
                const int b = (int)(row / hw), rem = (int)(row % hw);
                const int gy = rem / p.grid_w, gx = rem % p.grid_w;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    if ((a == 2) != (half == 1)) continue;
                    const int lo = 85 * a;
                    float tx = 0.f, ty = 0.f, tw = 0.f, th = 0.f, obj = 0.f, best = -INFINITY;
                    int besti = 0;
#pragma unroll
                    for (int c = lo / 32; c <= (lo + 84) / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld32(t0 + (uint32_t)(c * 32), v);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = 32 * c + j, slot = col - lo;
                            if (slot < 0 || slot >= 85) continue;
                            const float x = __uint_as_float(v[j]) + __ldg(p.bias + col);
                            if (slot == 0) tx = x;
                            else if (slot == 1) ty = x;
                            else if (slot == 2) tw = x;
                            else if (slot == 3) th = x;
                            else if (slot == 4) obj = sigmoidf_rn(x);
                            else {
                                const float sc = obj * sigmoidf_rn(x);
                                if (slot == 5 || sc > best) { best = sc; besti = slot - 5; }
                            }
                        }
                    }
                    if (row_ok) {
                        const int64_t cell = (((int64_t)b * 3 + a) * p.grid_h + gy) * p.grid_w + gx;
                        const float bx = __fdiv_rn((float)gx + sigmoidf_rn(tx), (float)p.grid_w);
                        const float by = __fdiv_rn((float)gy + sigmoidf_rn(ty), (float)p.grid_h);
                        const float bw = __ldg(p.anchor_wh + 2 * a) * expf(tw), bh = __ldg(p.anchor_wh + 2 * a + 1) * expf(th);
                        const float hx = bw * 0.5f, hy = bh * 0.5f;
                        reinterpret_cast<float4*>(p.dec_boxes)[cell] = make_float4(bx - hx, by - hy, bx + hx, by + hy);
                        p.dec_scores[cell] = best;
                        p.dec_cls[cell] = besti;
                        if (p.dec_obj != nullptr) p.dec_obj[cell] = obj;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[it0 & 1u]);
                acc_it += 1u;
                continue;
            }
            float mean = 0.f, inv = 0.f;
            if (p.epilogue == HVS_GEMM_EPI_LAYERNORM) {
                // statistics over the whole row (both accumulators), shifted by the row's first element
                float shift = 0.f, s1 = 0.f, s2 = 0.f;
                for (int sub = 0; sub < p.n_sub; ++sub) {
                    const uint32_t t0 = tmem_base + lane_base + ((it0 + (uint32_t)sub) & 1u) * kMaxBN;
                    for (int c = 0; c < chunks; ++c) {
                        uint32_t v[32];
                        tmem_ld32(t0 + (uint32_t)(c * 32), v);
                        tmem_wait_ld();
                        if (sub == 0 && c == 0) shift = __uint_as_float(v[0]);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float d = __uint_as_float(v[j]) - shift;
                            s1 += d;
                            s2 = fmaf(d, d, s2);
                        }
                    }
                }
                const float rn = 1.0f / (float)p.N;
                const float m1 = s1 * rn;
                mean = shift + m1;
                inv = rsqrtf(fmaxf(s2 * rn - m1 * m1, 0.f) + p.ln_eps);
            }
            for (int sub = 0; sub < p.n_sub; ++sub) {
                const uint32_t it = it0 + (uint32_t)sub;
                const uint32_t t0 = tmem_base + lane_base + (it & 1u) * kMaxBN;
                const int n0 = (n_out * p.n_sub + sub) * p.BN;
                for (int c = half; c < chunks; c += 2) {
                    uint32_t v[32];
                    tmem_ld32(t0 + (uint32_t)(c * 32), v);
                    tmem_wait_ld();
                    const int nc = n0 + c * 32;
                    if (p.epilogue == HVS_GEMM_EPI_BIAS_GELU && !p.out_f32) {
                        // the hot epilogue (two of the module's four GEMMs, 88 % of its FLOPs): bias + GELU + bf16 pack, packed
                        uint32_t o[16];
                        const float4* b4 = reinterpret_cast<const float4*>(p.bias + nc);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bb = __ldg(b4 + j);
                            o[2 * j] = gelu2_bf16(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
                            o[2 * j + 1] = gelu2_bf16(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
                        }
                        if (row_ok) {
                            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + nc;
#pragma unroll
                            for (int j = 0; j < 16; j += 4) st_global_v4(op + 2 * j, o[j], o[j + 1], o[j + 2], o[j + 3]);
                        }
                        continue;
                    }
                    if (p.epilogue == HVS_GEMM_EPI_BIAS_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(gelu_erf(__uint_as_float(v[j]) + __ldg(p.bias + nc + j)));
                    } else if (p.epilogue == HVS_GEMM_EPI_LAYERNORM) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            v[j] = __float_as_uint((__uint_as_float(v[j]) - mean) * inv * __ldg(p.ln_w + nc + j) + __ldg(p.ln_b + nc + j));
                    } else if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + nc + j));
                    }
                    if (row_ok) {
                        if (p.out_f32) {
                            float* o = reinterpret_cast<float*>(p.out) + row * p.ldo + nc;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) st_global_v4(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
                            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + nc;
#pragma unroll
                            for (int j = 0; j < 32; j += 8)
                                st_global_v4(o + j, pack_bf16(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                             pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])),
                                             pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])),
                                             pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                for (int sub = 0; sub < p.n_sub; ++sub) mbar_arrive(&acc_empty[(it0 + (uint32_t)sub) & 1u]);
            acc_it += (uint32_t)p.n_sub;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace
}  // namespace hvs

extern "C" int hvs_gemm_bf16(const void* a0, int64_t lda0, const void* b0, int K0, const void* a1, int64_t lda1,
                             const void* b1, int K1, const float* bias, const float* ln_w, const float* ln_b, float ln_eps,
                             void* out, int out_dtype, int64_t ldo, int64_t M, int N, int epilogue, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (M < 0 || N <= 0 || K0 <= 0 || K1 < 0) return HVS_ERR_BAD_ARG;
    if (M == 0) return HVS_OK;
    if (!a0 || !b0 || !out || (K1 > 0 && (!a1 || !b1))) return HVS_ERR_BAD_ARG;
    // K need not be a multiple of the 64-element stage: the tensor maps carry the true extents and TMA zero-fills the
    // rest of the box on both operands (so a D = 32 layer needs no padded copies); rows must be 16-byte multiples
    if (K0 % 8 || K1 % 8 || N % 32 || lda0 < K0 || (K1 > 0 && lda1 < K1) || ldo < N) return HVS_ERR_UNSUPPORTED;
    if (lda0 % 8 || (K1 > 0 && lda1 % 8) || ldo % 8) return HVS_ERR_ALIGNMENT;
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return HVS_ERR_ALIGNMENT;
    if (M >= ((int64_t)1 << 31)) return HVS_ERR_UNSUPPORTED;
    if (out_dtype != HVS_DTYPE_F32 && out_dtype != HVS_DTYPE_BF16) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(a0) | reinterpret_cast<uintptr_t>(b0) | reinterpret_cast<uintptr_t>(a1) |
         reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(out)) & 15)
        return HVS_ERR_ALIGNMENT;
    GemmParams p{};
    p.bias = bias; p.ln_w = ln_w; p.ln_b = ln_b; p.out = out; p.ldo = ldo; p.M = M; p.N = N;
    p.BN = N >= kMaxBN ? kMaxBN : N;
    if (N % p.BN) return HVS_ERR_UNSUPPORTED;
    p.n_sub = 1;
    p.epilogue = epilogue;
    if (epilogue == HVS_GEMM_EPI_LAYERNORM) {
        if (!ln_w || !ln_b || N > 2 * kMaxBN) return N > 2 * kMaxBN ? HVS_ERR_UNSUPPORTED : HVS_ERR_BAD_ARG;
        p.n_sub = N / p.BN;
    } else if (epilogue == HVS_GEMM_EPI_BIAS_GELU) {
        if (!bias) return HVS_ERR_BAD_ARG;
    } else if (epilogue != HVS_GEMM_EPI_NONE) {
        return HVS_ERR_UNSUPPORTED;
    }
    p.kb0 = (K0 + kBK - 1) / kBK; p.kb1 = (K1 + kBK - 1) / kBK;
    p.stages = kStageBudget / (kABytes + p.BN * 128);
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    p.out_f32 = out_dtype == HVS_DTYPE_F32;
    p.ln_eps = ln_eps;
    const int64_t m_tiles = (M + kBM - 1) / kBM;
    p.n_outer = N / (p.BN * p.n_sub);
    p.num_tiles = (int)(m_tiles * p.n_outer);
    CUtensorMap ta0, tb0, ta1, tb1;
    int rc = make_tmap_bf16_2d_ld(&ta0, a0, (uint64_t)M, (uint64_t)K0, (uint64_t)lda0, kBM);
    if (rc) return rc;
    rc = make_tmap_bf16_2d_ld(&tb0, b0, (uint64_t)N, (uint64_t)K0, (uint64_t)K0, (uint32_t)p.BN);
    if (rc) return rc;
    if (K1 > 0) {
        rc = make_tmap_bf16_2d_ld(&ta1, a1, (uint64_t)M, (uint64_t)K1, (uint64_t)lda1, kBM);
        if (rc) return rc;
        rc = make_tmap_bf16_2d_ld(&tb1, b1, (uint64_t)N, (uint64_t)K1, (uint64_t)K1, (uint32_t)p.BN);
        if (rc) return rc;
    } else {
        ta1 = ta0; tb1 = tb0;
    }
    HVS_SET_MAX_SMEM(k2_gemm_kernel, kSmemBytes);
    const int sms = sm_count();
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    timer_begin(4, stream);
    k2_gemm_kernel<<<grid, kThreads, kSmemBytes, stream>>>(ta0, tb0, ta1, tb1, p);
    timer_end(4, stream);
    count_launch();
    return launch_status();
}

// Fused YOLOPredictionHead tail: the 1x1 prediction convolution (yolo_head.py:193-194) as a
// tcgen05 GEMM over the head's token view [B*H*W, C_in] with YOLODecoder.forward (:241-285) as its epilogue.  The raw
// [B, 3, H, W, 85] predictions (548 MB per batch of 64 at 640 x 640 in fp32) are never written: per pixel the kernel emits
// 3 x (box, score, class).  The score threshold and the order-preserving compaction of the survivors happen where the
// reference does them (yolo_head.py:600-622), inside hvs_post_process's NMS kernel, which reads these dense outputs.
extern "C" int hvs_head_decode_fused(const void* tokens, int64_t ld_tokens, const void* weight256, const float* bias256,
                                     const float* anchor_wh, float* boxes, float* class_scores, int64_t* class_idx,
                                     float* objectness, int B, int H, int W, int C_in, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || H <= 0 || W <= 0 || C_in <= 0) return HVS_ERR_BAD_ARG;
    if (B == 0) return HVS_OK;
    if (!tokens || !weight256 || !bias256 || !anchor_wh || !boxes || !class_scores || !class_idx) return HVS_ERR_BAD_ARG;
    if (C_in % 8 || ld_tokens % 8 || ld_tokens < C_in) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(tokens) | reinterpret_cast<uintptr_t>(weight256) | reinterpret_cast<uintptr_t>(boxes)) & 15) return HVS_ERR_ALIGNMENT;
    const int64_t M = (int64_t)B * H * W;
    if (M >= ((int64_t)1 << 31)) return HVS_ERR_UNSUPPORTED;
    GemmParams p{};
    p.bias = bias256; p.M = M; p.N = 256; p.BN = 256; p.n_sub = 1; p.n_outer = 1;
    p.kb0 = (C_in + kBK - 1) / kBK; p.kb1 = 0;
    p.stages = kStageBudget / (kABytes + p.BN * 128);
    p.epilogue = kEpiYoloDecode;
    p.num_tiles = (int)((M + kBM - 1) / kBM);
    p.anchor_wh = anchor_wh; p.dec_boxes = boxes; p.dec_scores = class_scores; p.dec_cls = class_idx; p.dec_obj = objectness;
    p.grid_h = H; p.grid_w = W;
    CUtensorMap ta, tb;
    int rc = make_tmap_bf16_2d_ld(&ta, tokens, (uint64_t)M, (uint64_t)C_in, (uint64_t)ld_tokens, kBM);
    if (rc) return rc;
    rc = make_tmap_bf16_2d_ld(&tb, weight256, 256, (uint64_t)C_in, (uint64_t)C_in, 256);
    if (rc) return rc;
    HVS_SET_MAX_SMEM(k2_gemm_kernel, kSmemBytes);
    const int sms = sm_count();
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    timer_begin(5, stream);
    k2_gemm_kernel<<<grid, kThreads, kSmemBytes, stream>>>(ta, tb, ta, tb, p);
    timer_end(5, stream);
    count_launch();
    return launch_status();
}
