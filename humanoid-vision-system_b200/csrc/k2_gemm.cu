// Token-path GEMMs of the reference-literal module (K2), ManifoldHyperConnection.forward
// (src/models/manifold_layers.py:248-270):
//     z  = LN_pre(x) @ H_pre                                       EPI_NONE       (bf16 out)
//     z  = GELU(z @ W1^T + b1),  z = GELU(z @ W2^T + b2)           EPI_BIAS_GELU  (bf16 out)
//     y  = LN_post(z @ H_post + x @ H_res)                         EPI_LAYERNORM  (two operand pairs, one accumulator)
// One persistent warp-specialised kernel for sm_100a: a TMA producer thread, a tcgen05.mma issuer thread, sixteen
// epilogue warps.  D[128 x BN] fp32 accumulators live in tensor memory, double buffered (2 x 256 columns), so the
// epilogue of tile i (tcgen05.ld -> bias / GELU / LayerNorm in registers -> global) overlaps the MMAs of tile i+1.
// Operands: A [M, K] bf16 row-major (tokens x features), B [N, K] bf16 row-major (= nn.Linear's weight layout; the
// static coefficient kernel emits H_pre^T / H_post^T / H_res^T in it), both staged by TMA as 128-byte-swizzled
// K-major tiles of 64 K-elements, 4-6 stages.
// LayerNorm needs the whole output row: with N <= 256 the tile spans it; with 256 < N <= 512 the two halves of the
// row go to the two accumulator buffers and the epilogue normalises across both (no overlap for that GEMM).
//
// Training (hvs_gemm_bf16_ex) runs the SAME kernel for the module's forward and for all ten backward GEMMs:
//   * either operand may be MN-major, i.e. given as its transpose in place ([K, rows] row-major): TMA lands 64(K) x 64
//     boxes as 128-byte-swizzled MN-major atoms (8 K-rows x 128 B) and the descriptors say so.  Data gradients
//     dX = dY W read W [out, in] as it lies (B MN-major); weight gradients dW = dY^T X contract over the TOKENS with
//     both activations [T, features] as they lie (A and B MN-major).  No transposed copies anywhere.
//   * split-K for the weight gradients (tiny outputs, K = millions of tokens): fp32 partials per split, summed in a
//     fixed order by hvs_reduce_partials (bitwise reproducible).
//   * EPI_BIAS_GELU_SAVE  out2 = z = bf16(acc + b), out = dropout(GELU(z)): what the backward needs, one pass.
//   * EPI_DGELU           out = acc * GELU'(z[m, n]) * dropout mask / (1 - p): the activation backward in the
//                         data-gradient GEMM's epilogue.  The mask is a counter-based hash of (seed, row, column pair):
//                         recomputed, never stored.
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
constexpr int kEpiWarps = 16;               // four per tensor-memory lane quadrant: the epilogues are issue / latency bound
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
    int num_tiles;          // m_tiles * n_outer * splits
    int n_outer;            // N / (BN * n_sub)
    // training extensions (hvs_gemm_bf16_ex)
    int a_mn, b_mn;         // operand given MN-major (its transpose in place)
    int b_bytes;            // B bytes per stage
    int splits;             // split-K: partial s covers K blocks [s * kb_split, (s + 1) * kb_split)
    int kb_split;
    int64_t split_stride;   // elements between the partial outputs
    const __nv_bfloat16* aux;   // EPI_DGELU: pre-activation z [M, N]
    int64_t ld_aux;
    __nv_bfloat16* out2;    // EPI_BIAS_GELU_SAVE: pre-activation z [M, N]
    int64_t ldo2;
    uint32_t drop_thr;      // dropout: element dropped iff its 16 hash bits < drop_thr (0 = no dropout)
    float drop_scale;       // 1 / (1 - p)
    uint32_t seed;
    const uint32_t* seed_dev;   // optional device word mixed into the seed (a step counter that lives on the device: CUDA graphs)
    float* colsum_part;     // EPI_DGELU, optional: [4 * m_tiles, N] column sums of the output per 32-row group (bias gradients)
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
__device__ __forceinline__ uint4 ld_global_nc_v4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

// GELU'(x) = Phi(x) + x phi(x) with the same erf approximation (absolute error <= 1e-6), two values at once
__device__ __forceinline__ void gelu_grad2(float x0, float x1, float& d0, float& d1) {
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
    // phi(x) = exp(-x^2 / 2) / sqrt(2 pi) = 2^(-x^2 * log2(e) / 2) * 0.3989...
    float q0, q1;
    upk2(mul2(mul2(x, x), pk2(-0.72134752044448170368f, -0.72134752044448170368f)), q0, q1);
    float p0, p1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(q0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(q1));
    const u64 cdf = fma2(pk2(e0, e1), pk2(0.5f, 0.5f), pk2(0.5f, 0.5f));
    upk2(fma2(mul2(x, pk2(0.3989422804014327f, 0.3989422804014327f)), pk2(p0, p1), cdf), d0, d1);
}
// dropout keep decisions of the column pair (n, n + 1), n even, of row m: 16 hash bits each (lowbias32 finaliser)
__device__ __forceinline__ uint32_t drop_hash(uint32_t seed, uint32_t row, uint32_t colpair) {
    uint32_t h = (row * 0x9E3779B1u) ^ (colpair * 0x85EBCA77u) ^ seed;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <bool kYolo>                        // the decode epilogue is its own instantiation (its register needs stay out of the GEMMs')
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
    const int stage_bytes = kABytes + p.b_bytes;
    const int num_kb = p.kb0 + p.kb1;
    const int tiles_mn = p.num_tiles / p.splits;         // a tile index = split * tiles_mn + (m_blk * n_outer + n_out)

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
                const int split = tile / tiles_mn, mn = tile - split * tiles_mn;
                const int m_blk = mn / p.n_outer, n_out = mn % p.n_outer;
                const int kb_lo = split * p.kb_split, kb_hi = min(num_kb, kb_lo + p.kb_split);
                for (int sub = 0; sub < p.n_sub; ++sub) {
                    const int n_row0 = (n_out * p.n_sub + sub) * p.BN;
                    for (int kb = kb_lo; kb < kb_hi; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
                        if (kb < p.kb0) {
                            if (!p.a_mn) {
                                tma_load_2d(sa, &tm_a0, &full[stage], kb * kBK, m_blk * kBM);
                            } else {                                 // [K, M] in memory: two 64(K) x 64(M) boxes = MN-major atoms
                                tma_load_2d(sa, &tm_a0, &full[stage], m_blk * kBM, kb * kBK);
                                tma_load_2d(sa + 8192, &tm_a0, &full[stage], m_blk * kBM + 64, kb * kBK);
                            }
                            if (!p.b_mn) {
                                tma_load_2d(sa + kABytes, &tm_b0, &full[stage], kb * kBK, n_row0);
                            } else {
                                for (int i = 0; i * 8192 < p.b_bytes; ++i)
                                    tma_load_2d(sa + kABytes + i * 8192, &tm_b0, &full[stage], n_row0 + 64 * i, kb * kBK);
                            }
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
            const uint32_t idesc = umma_idesc_bf16(kBM, p.BN, p.a_mn, p.b_mn);
            // K-major: 8-row groups 1 KB apart, a K = 16 step is +32 B.  MN-major: 64-element atoms along M / N 8 KB apart
            // (one TMA box each), the next 8 K-rows 1 KB on, a K = 16 step is two of those = +2 KB.
            const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
            const uint64_t a_step = p.a_mn ? 128u : 2u, b_step = p.b_mn ? 128u : 2u;
            int stage = 0;
            uint32_t phase = 0;
            uint32_t acc_it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int split = tile / tiles_mn;
                const int kb_lo = split * p.kb_split, kb_hi = min(num_kb, kb_lo + p.kb_split);
                for (int sub = 0; sub < p.n_sub; ++sub, ++acc_it) {
                    const uint32_t buf = acc_it & 1u, aph = (acc_it >> 1) & 1u;
                    mbar_wait(&acc_empty[buf], aph ^ 1u);            // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kMaxBN;
                    for (int kb = kb_lo; kb < kb_hi; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = base + (uint32_t)(stage * stage_bytes);
                        const uint64_t adesc = umma_smem_desc(sa, a_lbo, 1024, kUmmaLayoutSw128);
                        const uint64_t bdesc = umma_smem_desc(sa + kABytes, b_lbo, 1024, kUmmaLayoutSw128);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)
                            umma_bf16_ss(d_tmem, adesc + a_step * (uint64_t)k, bdesc + b_step * (uint64_t)k, idesc, (uint32_t)((kb > kb_lo) | (k != 0)));
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
        const int part = (warp - 2) >> 2;                            // which 32-column chunks it takes (c = part mod 4)
        constexpr int kParts = kEpiWarps / 4;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t acc_it = 0;
        const int chunks = p.BN >> 5;
        const uint32_t seed = p.seed_dev != nullptr ? (__ldg(p.seed_dev) * 0x9E3779B1u) ^ p.seed : p.seed;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int split = tile / tiles_mn, mn = tile - split * tiles_mn;
            const int m_blk = mn / p.n_outer, n_out = mn % p.n_outer;
            const int64_t row = (int64_t)m_blk * kBM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const int64_t out_off = (int64_t)split * p.split_stride;   // split-K partial
            const uint32_t it0 = acc_it;
            for (int sub = 0; sub < p.n_sub; ++sub) {
                const uint32_t it = it0 + (uint32_t)sub;
                mbar_wait(&acc_full[it & 1u], (it >> 1) & 1u);
            }
            tc_fence_after();
            if constexpr (kYolo) {
                // ---- YOLODecoder.forward (yolo_head.py:241-285) on the accumulator row: a row is one pixel, its 255 columns are
                // 3 anchors x (tx, ty, tw, th, obj, 80 classes).  Same arithmetic and order as yolo_decode.cu; the raw
                // predictions never reach HBM.  Warp `part` of the quadrant takes anchor `part`.
                const uint32_t t0 = tmem_base + lane_base + (it0 & 1u) * kMaxBN;
                const int hw = p.grid_h * p.grid_w;
                const int b = (int)(row / hw), rem = (int)(row % hw);
                const int gy = rem / p.grid_w, gx = rem % p.grid_w;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    if (a != part) continue;                             // one anchor per warp of the quadrant (the fourth idles)
                    const int lo = 85 * a;
                    float tx = 0.f, ty = 0.f, tw = 0.f, th = 0.f, obj = 0.f, best = -INFINITY;
                    int besti = 0;
#pragma unroll
                    for (int c = lo / 16; c <= (lo + 84) / 16; ++c) {
                        uint32_t v[16];
                        tmem_ld16(t0 + (uint32_t)(c * 16), v);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int col = 16 * c + j, slot = col - lo;
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
                for (int c = part; c < chunks; c += kParts) {
                    const int nc = n0 + c * 32;
                    if (!p.out_f32 && (p.epilogue == HVS_GEMM_EPI_BIAS_GELU || p.epilogue == HVS_GEMM_EPI_BIAS_GELU_SAVE ||
                                       p.epilogue == HVS_GEMM_EPI_DGELU)) {
                        // the hot epilogues (the module's MLP GEMMs, forward and backward: 88 % of its FLOPs), 16 columns at a
                        // time (register budget of a 576-thread CTA), packed fp32x2 arithmetic
#pragma unroll 1
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t v[16], o[8];
                            tmem_ld16(t0 + (uint32_t)(c * 32 + hh * 16), v);
                            const int nh = nc + hh * 16;
                            if (p.epilogue == HVS_GEMM_EPI_DGELU) {
                                // training backward: d z = d a * mask / (1 - p) * GELU'(z), z read back as the forward stored it
                                const __nv_bfloat16* zp = p.aux + (row_ok ? row : 0) * p.ld_aux + nh;
                                const uint4 zq0 = ld_global_nc_v4(zp), zq1 = ld_global_nc_v4(zp + 8);
                                const uint32_t zz[8] = {zq0.x, zq0.y, zq0.z, zq0.w, zq1.x, zq1.y, zq1.z, zq1.w};
                                tmem_wait_ld();
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    float d0, d1;
                                    gelu_grad2(bf16lo(zz[j]), bf16hi(zz[j]), d0, d1);
                                    if (p.drop_thr != 0u) {
                                        const uint32_t h = drop_hash(seed, (uint32_t)row, (uint32_t)(nh >> 1) + j);
                                        d0 *= (h & 0xffffu) < p.drop_thr ? 0.f : p.drop_scale;
                                        d1 *= (h >> 16) < p.drop_thr ? 0.f : p.drop_scale;
                                    }
                                    const float e0 = __uint_as_float(v[2 * j]) * d0, e1 = __uint_as_float(v[2 * j + 1]) * d1;
                                    o[j] = pack_bf16(e0, e1);
                                    v[2 * j] = __float_as_uint(row_ok ? bf16lo(o[j]) : 0.f);       // what the bias gradient sums: the bf16 d z
                                    v[2 * j + 1] = __float_as_uint(row_ok ? bf16hi(o[j]) : 0.f);
                                }
                                if (p.colsum_part != nullptr) {
                                    // column sums over the warp's 32 rows by a transpose-reduce: after the exchange with lane ^ 16
                                    // a lane keeps 8 of its 16 columns (summed over 2 rows), then 4, 2, 1; 16 shuffles in all, the
                                    // order of the additions is fixed.  Lane pairs (bit 0) end with the same column.
                                    float k8[8], k4[4], k2[2];
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float a = __uint_as_float(v[i]), b = __uint_as_float(v[i + 8]);
                                        const bool hi = lane & 16;
                                        k8[i] = (hi ? b : a) + __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
                                    }
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const bool hi = lane & 8;
                                        k4[i] = (hi ? k8[i + 4] : k8[i]) + __shfl_xor_sync(0xffffffffu, hi ? k8[i] : k8[i + 4], 8);
                                    }
#pragma unroll
                                    for (int i = 0; i < 2; ++i) {
                                        const bool hi = lane & 4;
                                        k2[i] = (hi ? k4[i + 2] : k4[i]) + __shfl_xor_sync(0xffffffffu, hi ? k4[i] : k4[i + 2], 4);
                                    }
                                    const bool hi2 = lane & 2;
                                    float k1 = (hi2 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, hi2 ? k2[0] : k2[1], 2);
                                    k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
                                    if (!(lane & 1)) {
                                        const int col = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
                                        p.colsum_part[((size_t)m_blk * 4 + q) * p.N + nh + col] = k1;
                                    }
                                }
                            } else {
                                tmem_wait_ld();
                                const float4* b4 = reinterpret_cast<const float4*>(p.bias + nh);
                                if (p.epilogue == HVS_GEMM_EPI_BIAS_GELU) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const float4 bb = __ldg(b4 + j);
                                        o[2 * j] = gelu2_bf16(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
                                        o[2 * j + 1] = gelu2_bf16(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
                                    }
                                } else {
                                    // training forward: z = bf16(acc + b) -> out2, dropout(GELU(z)) -> out (the autocast convention: the
                                    // Linear's bf16 output is what GELU sees and what the backward differentiates at)
                                    uint32_t zz[8];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const float4 bb = __ldg(b4 + j);
                                        zz[2 * j] = pack_bf16(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y);
                                        zz[2 * j + 1] = pack_bf16(__uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
                                    }
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const uint32_t g = gelu2_bf16(bf16lo(zz[j]), bf16hi(zz[j]));
                                        if (p.drop_thr == 0u) {
                                            o[j] = g;
                                        } else {
                                            const uint32_t h = drop_hash(seed, (uint32_t)row, (uint32_t)(nh >> 1) + j);
                                            const float k0 = (h & 0xffffu) < p.drop_thr ? 0.f : p.drop_scale, k1 = (h >> 16) < p.drop_thr ? 0.f : p.drop_scale;
                                            o[j] = pack_bf16(bf16lo(g) * k0, bf16hi(g) * k1);
                                        }
                                    }
                                    if (row_ok) {
                                        __nv_bfloat16* zp = p.out2 + row * p.ldo2 + nh;
                                        st_global_v4(zp, zz[0], zz[1], zz[2], zz[3]);
                                        st_global_v4(zp + 8, zz[4], zz[5], zz[6], zz[7]);
                                    }
                                }
                            }
                            if (row_ok) {
                                __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + nh;
                                st_global_v4(op, o[0], o[1], o[2], o[3]);
                                st_global_v4(op + 8, o[4], o[5], o[6], o[7]);
                            }
                        }
                        continue;
                    }
                    uint32_t v[32];
                    tmem_ld32(t0 + (uint32_t)(c * 32), v);
                    tmem_wait_ld();
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
                            float* o = reinterpret_cast<float*>(p.out) + out_off + row * p.ldo + nc;
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

namespace hvs {
namespace {

// fixed-order sum of split-K partials: out[i] = sum_s part[s * stride + i].  Eight lanes per output float4 stride over the
// splits and combine in a butterfly (always the same order): a [64 x 64] gradient cut into 146 splits was 67 us with one
// thread walking all of them.
__global__ void __launch_bounds__(256) reduce_partials_few_kernel(const float4* __restrict__ part, int splits, int64_t stride4, int64_t n4,
                                                                  float4* __restrict__ out) {      // a handful of splits: one thread per output
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = part[i];
        for (int s = 1; s < splits; ++s) {
            const float4 b = part[(int64_t)s * stride4 + i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        out[i] = a;
    }
}
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float4* __restrict__ part, int splits, int64_t stride4, int64_t n4,
                                                              float4* __restrict__ out) {
    const int l = threadIdx.x & 7;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < ((n4 + 31) & ~(int64_t)31); i += ((int64_t)gridDim.x * blockDim.x) >> 3) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n4)
            for (int s = l; s < splits; s += 8) {
                const float4 b = part[(int64_t)s * stride4 + i];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
        }
        if (l == 0 && i < n4) out[i] = a;
    }
}

// column sums of a bf16 matrix (bias gradients): stage 1, one CTA per contiguous block of rows, 8 columns per thread;
// stage 2, a warp per column sums the partials (lanes stride over them, fixed-order butterfly)
__global__ void __launch_bounds__(256) colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int64_t rows, int cols,
                                                             int64_t rows_per_cta, float* __restrict__ part) {
    const int groups = cols >> 3;                          // 8-column groups
    const int lanes = groups < 256 ? groups : 256;         // threads along the columns
    const int rsteps = 256 / lanes;                        // rows handled concurrently
    const int cg = threadIdx.x % lanes, rg = threadIdx.x / lanes;
    __shared__ float red[256 * 8];
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    for (int g0 = 0; g0 < groups; g0 += lanes) {
        const int g = g0 + cg;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (g < groups && rg < rsteps) {
            const __nv_bfloat16* col = x + 8 * g;
            int64_t r = r0 + rg;
            for (; r + 7 * rsteps < r1; r += 8 * rsteps) {          // eight rows in flight (the loads are the latency)
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ld_global_nc_v4(col + (r + u * rsteps) * ld);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    acc[0] += bf16lo(v[u].x); acc[1] += bf16hi(v[u].x); acc[2] += bf16lo(v[u].y); acc[3] += bf16hi(v[u].y);
                    acc[4] += bf16lo(v[u].z); acc[5] += bf16hi(v[u].z); acc[6] += bf16lo(v[u].w); acc[7] += bf16hi(v[u].w);
                }
            }
            for (; r < r1; r += rsteps) {
                const uint4 v = ld_global_nc_v4(col + r * ld);
                acc[0] += bf16lo(v.x); acc[1] += bf16hi(v.x); acc[2] += bf16lo(v.y); acc[3] += bf16hi(v.y);
                acc[4] += bf16lo(v.z); acc[5] += bf16hi(v.z); acc[6] += bf16lo(v.w); acc[7] += bf16hi(v.w);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
        __syncthreads();
        if (rg == 0 && g < groups) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = red[cg * 8 + j];
            for (int k = 1; k < rsteps; ++k)
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] += red[(k * lanes + cg) * 8 + j];
            float* dst = part + (int64_t)blockIdx.x * cols + 8 * g;
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = o[j];
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, int nparts, int cols, float* __restrict__ out) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= cols) return;
    float a = 0.f;
    for (int i = lane; i < nparts; i += 32) a += part[(int64_t)i * cols + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) out[c] = a;
}
inline int64_t colsum_rows_per_cta(int64_t rows) {           // at most 4 CTAs per SM, at least 64 rows each
    const int64_t max_parts = 4 * (int64_t)sm_count();
    int64_t rpc = (rows + max_parts - 1) / max_parts;
    return rpc < 64 ? 64 : rpc;
}

// column sums of an fp32 matrix [rows, cols] (the per-32-row partials the EPI_DGELU epilogue writes): lanes along the
// columns (coalesced), warps and gridDim.y over the rows, fixed order
__global__ void __launch_bounds__(256) colsum_f32_partial_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t rows_per_block,
                                                                 float* __restrict__ part) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float a = 0.f;
    if (c < cols) {
        int64_t r = r0 + wid;
        for (; r + 24 < r1; r += 32) {
            const float v0 = x[r * cols + c], v1 = x[(r + 8) * cols + c], v2 = x[(r + 16) * cols + c], v3 = x[(r + 24) * cols + c];
            a += v0; a += v1; a += v2; a += v3;
        }
        for (; r < r1; r += 8) a += x[r * cols + c];
    }
    __shared__ float sm[8][32];
    sm[wid][lane] = a;
    __syncthreads();
    if (wid == 0 && c < cols) {
        float t = sm[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += sm[k][lane];
        part[(size_t)blockIdx.y * cols + c] = t;
    }
}
inline int64_t colsum_f32_rows_per_block(int64_t rows) {
    int64_t rpb = (rows + 63) / 64;                       // at most 64 row blocks
    return rpb < 64 ? 64 : rpb;
}

int launch_gemm(const hvs_gemm_args& g, int timer_slot, cudaStream_t stream) {
    const int64_t M = g.M;
    const int N = g.N, K0 = g.K0, K1 = g.K1;
    if (M < 0 || N <= 0 || K0 <= 0 || K1 < 0) return HVS_ERR_BAD_ARG;
    if (M == 0) return HVS_OK;
    if (!g.a0 || !g.b0 || !g.out || (K1 > 0 && (!g.a1 || !g.b1))) return HVS_ERR_BAD_ARG;
    const int a_mn = g.a_mn_major ? 1 : 0, b_mn = g.b_mn_major ? 1 : 0;
    int splits = g.split_k > 1 ? g.split_k : 1;
    // K need not be a multiple of the 64-element stage: the tensor maps carry the true extents and TMA zero-fills the
    // rest of the box on both operands (so a D = 32 layer needs no padded copies); rows must be 16-byte multiples
    // (an MN-major operand has K as its ROW count: any K)
    if (((!a_mn || !b_mn) && K0 % 8) || K1 % 8 || N % 32) return HVS_ERR_UNSUPPORTED;
    if (K1 > 0 && (a_mn || b_mn || splits > 1)) return HVS_ERR_UNSUPPORTED;
    if (a_mn ? (M % 8 || g.lda0 < M) : (g.lda0 < K0)) return HVS_ERR_UNSUPPORTED;
    if (b_mn ? (g.ldb0 < N) : (g.ldb0 < K0)) return HVS_ERR_UNSUPPORTED;
    if ((K1 > 0 && (g.lda1 < K1 || g.ldb1 < K1)) || g.ldo < N) return HVS_ERR_UNSUPPORTED;
    if (g.lda0 % 8 || g.ldb0 % 8 || (K1 > 0 && (g.lda1 % 8 || g.ldb1 % 8)) || g.ldo % 8) return HVS_ERR_ALIGNMENT;
    if (g.bias && (reinterpret_cast<uintptr_t>(g.bias) & 15)) return HVS_ERR_ALIGNMENT;
    if (M >= ((int64_t)1 << 31) || (int64_t)K0 >= ((int64_t)1 << 31)) return HVS_ERR_UNSUPPORTED;
    if (g.out_dtype != HVS_DTYPE_F32 && g.out_dtype != HVS_DTYPE_BF16) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(g.a0) | reinterpret_cast<uintptr_t>(g.b0) | reinterpret_cast<uintptr_t>(g.a1) |
         reinterpret_cast<uintptr_t>(g.b1) | reinterpret_cast<uintptr_t>(g.out) | reinterpret_cast<uintptr_t>(g.out2) |
         reinterpret_cast<uintptr_t>(g.aux)) & 15)
        return HVS_ERR_ALIGNMENT;
    GemmParams p{};
    p.bias = g.bias; p.ln_w = g.ln_w; p.ln_b = g.ln_b; p.out = g.out; p.ldo = g.ldo; p.M = M; p.N = N;
    p.BN = N >= kMaxBN ? kMaxBN : N;
    if (N % p.BN) return HVS_ERR_UNSUPPORTED;
    p.n_sub = 1;
    p.epilogue = g.epilogue;
    p.out_f32 = g.out_dtype == HVS_DTYPE_F32;
    switch (g.epilogue) {
        case HVS_GEMM_EPI_NONE: break;
        case HVS_GEMM_EPI_LAYERNORM:
            if (!g.ln_w || !g.ln_b) return HVS_ERR_BAD_ARG;
            if (N > 2 * kMaxBN || splits > 1) return HVS_ERR_UNSUPPORTED;
            p.n_sub = N / p.BN;
            break;
        case HVS_GEMM_EPI_BIAS_GELU:
            if (!g.bias) return HVS_ERR_BAD_ARG;
            if (splits > 1) return HVS_ERR_UNSUPPORTED;
            break;
        case HVS_GEMM_EPI_BIAS_GELU_SAVE:
            if (!g.bias || !g.out2) return HVS_ERR_BAD_ARG;
            if (p.out_f32 || splits > 1 || g.ldo2 < N || g.ldo2 % 8) return HVS_ERR_UNSUPPORTED;
            break;
        case HVS_GEMM_EPI_DGELU:
            if (!g.aux) return HVS_ERR_BAD_ARG;
            if (p.out_f32 || splits > 1 || g.ld_aux < N || g.ld_aux % 8) return HVS_ERR_UNSUPPORTED;
            break;
        default: return HVS_ERR_UNSUPPORTED;
    }
    if (splits > 1 && (!p.out_f32 || g.bias)) return HVS_ERR_UNSUPPORTED;       // partials are plain fp32 accumulators
    if (!(g.dropout_p >= 0.f && g.dropout_p < 1.f)) return HVS_ERR_BAD_ARG;
    p.drop_thr = (uint32_t)(g.dropout_p * 65536.0f + 0.5f);
    p.drop_scale = p.drop_thr ? 65536.0f / (65536.0f - (float)p.drop_thr) : 1.0f;   // 1 / (1 - p) for the p actually realised
    p.seed = g.dropout_seed;
    p.seed_dev = g.dropout_seed_dev;
    p.colsum_part = g.epilogue == HVS_GEMM_EPI_DGELU ? g.colsum_partials : nullptr;
    p.aux = reinterpret_cast<const __nv_bfloat16*>(g.aux); p.ld_aux = g.ld_aux;
    p.out2 = reinterpret_cast<__nv_bfloat16*>(g.out2); p.ldo2 = g.ldo2;
    p.a_mn = a_mn; p.b_mn = b_mn;
    const int sms = sm_count();
    if ((g.epilogue == HVS_GEMM_EPI_NONE || g.epilogue == HVS_GEMM_EPI_BIAS_GELU) && splits == 1 && !a_mn && !b_mn) {
        // few row tiles (a batch-1 frame: 401 tokens = 4 tiles): narrower accumulators put the same work on 2-4x the SMs, each
        // with a quarter of the weight tile to fetch and of the epilogue to run (these launches are latency-, not math-bound)
        const int64_t mt = (M + kBM - 1) / kBM;
        while (p.BN > 64 && (p.BN / 2) % 32 == 0 && N % (p.BN / 2) == 0 && mt * (N / p.BN) * 2 <= sms) p.BN /= 2;
    }
    p.b_bytes = b_mn ? ((p.BN + 63) / 64) * 8192 : p.BN * 128;
    p.kb0 = (K0 + kBK - 1) / kBK; p.kb1 = (K1 + kBK - 1) / kBK;
    const int num_kb = p.kb0 + p.kb1;
    if (splits > num_kb) splits = num_kb;
    p.kb_split = (num_kb + splits - 1) / splits;
    splits = (num_kb + p.kb_split - 1) / p.kb_split;            // every split owns at least one K block
    p.splits = splits;
    p.split_stride = g.split_stride > 0 ? g.split_stride : M * g.ldo;
    p.stages = kStageBudget / (kABytes + p.b_bytes);
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    p.ln_eps = g.ln_eps;
    const int64_t m_tiles = (M + kBM - 1) / kBM;
    p.n_outer = N / (p.BN * p.n_sub);
    if (m_tiles * p.n_outer * splits >= ((int64_t)1 << 31)) return HVS_ERR_UNSUPPORTED;
    p.num_tiles = (int)(m_tiles * p.n_outer) * splits;
    CUtensorMap ta0, tb0, ta1, tb1;
    int rc = a_mn ? make_tmap_bf16_2d_ld(&ta0, g.a0, (uint64_t)K0, (uint64_t)M, (uint64_t)g.lda0, kBK)
                  : make_tmap_bf16_2d_ld(&ta0, g.a0, (uint64_t)M, (uint64_t)K0, (uint64_t)g.lda0, kBM);
    if (rc) return rc;
    rc = b_mn ? make_tmap_bf16_2d_ld(&tb0, g.b0, (uint64_t)K0, (uint64_t)N, (uint64_t)g.ldb0, kBK)
              : make_tmap_bf16_2d_ld(&tb0, g.b0, (uint64_t)N, (uint64_t)K0, (uint64_t)g.ldb0, (uint32_t)p.BN);
    if (rc) return rc;
    if (K1 > 0) {
        rc = make_tmap_bf16_2d_ld(&ta1, g.a1, (uint64_t)M, (uint64_t)K1, (uint64_t)g.lda1, kBM);
        if (rc) return rc;
        rc = make_tmap_bf16_2d_ld(&tb1, g.b1, (uint64_t)N, (uint64_t)K1, (uint64_t)g.ldb1, (uint32_t)p.BN);
        if (rc) return rc;
    } else {
        ta1 = ta0; tb1 = tb0;
    }
    HVS_SET_MAX_SMEM(k2_gemm_kernel<false>, kSmemBytes);
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    timer_begin(timer_slot, stream);
    k2_gemm_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(ta0, tb0, ta1, tb1, p);
    timer_end(timer_slot, stream);
    count_launch();
    return launch_status();
}

}  // namespace
}  // namespace hvs

extern "C" int hvs_gemm_bf16(const void* a0, int64_t lda0, const void* b0, int K0, const void* a1, int64_t lda1,
                             const void* b1, int K1, const float* bias, const float* ln_w, const float* ln_b, float ln_eps,
                             void* out, int out_dtype, int64_t ldo, int64_t M, int N, int epilogue, void* stream_) {
    if (epilogue != HVS_GEMM_EPI_NONE && epilogue != HVS_GEMM_EPI_BIAS_GELU && epilogue != HVS_GEMM_EPI_LAYERNORM)
        return HVS_ERR_UNSUPPORTED;
    hvs_gemm_args g{};
    g.a0 = a0; g.lda0 = lda0; g.b0 = b0; g.ldb0 = K0; g.K0 = K0;
    g.a1 = a1; g.lda1 = lda1; g.b1 = b1; g.ldb1 = K1; g.K1 = K1;
    g.bias = bias; g.ln_w = ln_w; g.ln_b = ln_b; g.ln_eps = ln_eps;
    g.out = out; g.out_dtype = out_dtype; g.ldo = ldo; g.M = M; g.N = N; g.epilogue = epilogue;
    return hvs::launch_gemm(g, 4, (cudaStream_t)stream_);
}

extern "C" int hvs_gemm_bf16_ex(const hvs_gemm_args* args, void* stream_) {
    if (!args) return HVS_ERR_BAD_ARG;
    return hvs::launch_gemm(*args, 6, (cudaStream_t)stream_);
}

// Split count for a weight-gradient GEMM out[M, N] = A^T B over K tokens: enough tiles to fill the SMs, every split at
// least 8 K blocks (512 tokens) long.
extern "C" int hvs_gemm_choose_split(int64_t M, int N, int64_t K) {
    using namespace hvs;
    if (M <= 0 || N <= 0 || K <= 0) return 1;
    const int bn = N >= kMaxBN ? kMaxBN : N;
    const int64_t tiles = ((M + kBM - 1) / kBM) * ((N + bn - 1) / bn);
    const int64_t num_kb = (K + kBK - 1) / kBK;
    int64_t s = (2 * (int64_t)sm_count() + tiles - 1) / tiles;       // about two waves of tiles
    if (s > num_kb / 8) s = num_kb / 8;
    if (s < 1) s = 1;
    if (s > 256) s = 256;
    const int64_t kb_split = (num_kb + s - 1) / s;
    return (int)((num_kb + kb_split - 1) / kb_split);
}

extern "C" int hvs_reduce_partials(const float* partials, int splits, int64_t split_stride, int64_t numel, float* out, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (splits < 1 || numel < 0 || split_stride < numel) return HVS_ERR_BAD_ARG;
    if (numel == 0) return HVS_OK;
    if (!partials || !out) return HVS_ERR_BAD_ARG;
    if (numel % 4 || split_stride % 4) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(out)) & 15) return HVS_ERR_ALIGNMENT;
    const int64_t n4 = numel / 4;
    const bool many = splits >= 16;                           // eight lanes per output only pay when there are splits to share
    int64_t blocks = ((many ? n4 * 8 : n4) + 255) / 256;
    if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
    if (many)
        reduce_partials_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(partials), splits, split_stride / 4, n4,
                                                                reinterpret_cast<float4*>(out));
    else
        reduce_partials_few_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(partials), splits, split_stride / 4, n4,
                                                                    reinterpret_cast<float4*>(out));
    count_launch();
    return launch_status();
}

extern "C" size_t hvs_colsum_f32_workspace(int64_t rows, int cols) {
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t rpb = hvs::colsum_f32_rows_per_block(rows);
    return (size_t)((rows + rpb - 1) / rpb) * (size_t)cols * 4;
}

extern "C" int hvs_colsum_f32(const float* x, int64_t rows, int cols, float* out, void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows < 0 || cols <= 0 || !out) return HVS_ERR_BAD_ARG;
    if (rows == 0) { HVS_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)cols * 4, stream)); return HVS_OK; }
    if (!x || !workspace) return HVS_ERR_BAD_ARG;
    if (workspace_bytes < hvs_colsum_f32_workspace(rows, cols)) return HVS_ERR_WORKSPACE;
    const int64_t rpb = colsum_f32_rows_per_block(rows);
    const int nblk = (int)((rows + rpb - 1) / rpb);
    colsum_f32_partial_kernel<<<dim3((cols + 31) / 32, nblk), 256, 0, stream>>>(x, rows, cols, rpb, reinterpret_cast<float*>(workspace));
    colsum_final_kernel<<<(cols + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const float*>(workspace), nblk, cols, out);
    count_launch(2);
    return launch_status();
}

extern "C" size_t hvs_colsum_bf16_workspace(int64_t rows, int cols) {
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t rpc = hvs::colsum_rows_per_cta(rows);
    return (size_t)((rows + rpc - 1) / rpc) * (size_t)cols * 4;
}

extern "C" int hvs_colsum_bf16(const void* x, int64_t ld, int64_t rows, int cols, float* out, void* workspace, size_t workspace_bytes,
                               void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows < 0 || cols <= 0 || !out) return HVS_ERR_BAD_ARG;
    if (cols % 8 || ld % 8 || ld < cols) return HVS_ERR_UNSUPPORTED;
    if (rows == 0) { HVS_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)cols * 4, stream)); return HVS_OK; }
    if (!x || !workspace || (reinterpret_cast<uintptr_t>(x) & 15)) return HVS_ERR_BAD_ARG;
    if (workspace_bytes < hvs_colsum_bf16_workspace(rows, cols)) return HVS_ERR_WORKSPACE;
    const int64_t rpc = colsum_rows_per_cta(rows);
    const int nparts = (int)((rows + rpc - 1) / rpc);
    colsum_partial_kernel<<<nparts, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, rows, cols, rpc,
                                                      reinterpret_cast<float*>(workspace));
    colsum_final_kernel<<<(cols + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const float*>(workspace), nparts, cols, out);
    count_launch(2);
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
    p.splits = 1; p.kb_split = p.kb0; p.b_bytes = p.BN * 128;
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
    HVS_SET_MAX_SMEM(k2_gemm_kernel<true>, kSmemBytes);
    const int sms = sm_count();
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    timer_begin(5, stream);
    k2_gemm_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(ta, tb, ta, tb, p);
    timer_end(5, stream);
    count_launch();
    return launch_status();
}
