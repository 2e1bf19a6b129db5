// The whole token path of the reference-literal module in ONE kernel per 128-token tile
// (ManifoldHyperConnection.forward, src/models/manifold_layers.py:248-267, eval mode):
//
//     xn = LayerNorm_pre(x)                                   (:250)   in the tile prologue, from the TMA-staged x tile
//     h0 = xn @ H_pre                                         (:253)   tcgen05.mma -> tensor memory -> bf16 -> shared memory
//     h1 = GELU(h0 @ W1^T + b1)                               (:164-165)   in 64-column chunks ...
//     h2 = GELU(h1 @ W2^T + b2)                               (:167-168)   ... each chunk is at once a K block of this GEMM
//     y  = LayerNorm_post(h2 @ H_post + x @ H_res)            (:259-267)   one accumulator, normalised in the epilogue
//
// No intermediate touches HBM: per token the kernel reads D and writes D values (the five-launch path in k2_gemm.cu
// moves 12 H + 3 D values per token, which is what bounds it for the backbone's small widths: D = 32 / 64 with
// T = 6.5 M tokens at batch 64).  Built for those shapes: D in {32, 64}, hidden H = 4 D.
//
// Roles (one CTA per SM, persistent over tiles): a TMA producer thread (x tile + a ring of 32 KB weight slots, in exactly
// the order the MMAs consume them), a tcgen05.mma issuer thread, sixteen epilogue warps (four per tensor-memory lane
// quadrant, splitting the columns in 16-column units) that also do the LayerNorm prologue and write every intermediate into shared memory
// in the 128-byte-swizzled K-major layout the next MMA reads it in.  Tensor memory: [0, H) = h0's accumulator, later
// h2's; [256, 512) = two accumulators for the 64-column h1 chunks, the first of which is reused for the output.
// h1 chunks are 64 wide (one K block of the next GEMM): with 128-wide chunks their two buffers took 64 KB and left a
// two-slot weight ring at D = 64, which bound the chunk loop by TMA round trips (34 k cycles per tile).
// The h1 chunk loop is software-pipelined: the MMAs of chunk j+1 run while the epilogue warps apply GELU to chunk j.
// Tried and dropped: 64-column h1 chunks (frees 32 KB for a third / fourth weight slot) -- 5 % SLOWER (D = 64: 6.20 -> 6.56 ms
// per 6.5 M tokens): the tile time is set by the chain of ~10 MMA <-> epilogue hand-offs, not by the weight ring.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "umma_sm100.cuh"

namespace hvs {
namespace {

constexpr int kBM = 128;
constexpr int kSlotBytes = 32768;
constexpr int kEpiWarps = 16;               // four per tensor-memory lane quadrant, each taking every fourth 16-column unit
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kTmemCols = 512;
constexpr uint32_t kColAcc1 = 256;

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// same packed A&S 7.1.28 GELU as k2_gemm.cu (absolute error <= 6e-7)
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

struct ChainParams {
    const float* b1;        // [2H]
    const float* b2;        // [H]
    const float* ln_pre_w;  // [D]
    const float* ln_pre_b;
    const float* ln_post_w;
    const float* ln_post_b;
    void* out;              // [T, D]
    int64_t T;
    int num_tiles;
    int out_f32;
    float eps_pre, eps_post;
};

// byte offset of the 16-byte group g (8 bf16) of row r inside a [rows x 64] K-major tile with the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t g) { return (r >> 3) * 1024u + (r & 7u) * 128u + ((g ^ (r & 7u)) << 4); }

template <int D, int H>
struct Cfg {
    static constexpr int H2 = 2 * H;
    static constexpr int CW = 64;                        // h1 chunk width = one K block of the h2 GEMM
    static constexpr int NCH = H2 / CW;                  // h1 chunks
    static constexpr int KBH = H / 64;                   // 64-wide K blocks of an H-wide operand
    static constexpr int kOffX = 0;                      // x tile   [128 x 64] (TMA, columns >= D zero-filled)
    static constexpr int kOffXN = 16384;                 // LN(x)    [128 x 64]
    static constexpr int kOffH0 = 32768;                 // h0 / h2  [128 x H]  = KBH blocks of 16 KB
    static constexpr int kOffH1 = kOffH0 + KBH * 16384;  // h1 chunk [128 x 64] x 2 buffers
    static constexpr int kOffRing = kOffH1 + 2 * 16384;
    static constexpr int kSlots = (D == 32) ? 4 : 3;
    static constexpr int kOffBar = kOffRing + kSlots * kSlotBytes;
    static constexpr int kSmemBytes = kOffBar + 256 + 1024;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
    static_assert(H <= 256 && H % 128 == 0 && (D == 32 || D == 64), "shapes this kernel is built for");
};

enum { B_XFULL = 0, B_XFREE, B_XN, B_ACCA, B_HA, B_ACC1, B_H1 = B_ACC1 + 2, B_G3 = B_H1 + 2, B_ACC3 = B_G3 + 2, B_ACC3FREE, B_RFULL,
       B_REMPTY = B_RFULL + 4, B_COUNT = B_REMPTY + 4 };

template <int D, int H>
__global__ void __launch_bounds__(kThreads, 1)
k2_chain_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hpre,
                const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                const __grid_constant__ CUtensorMap tm_hpost, const __grid_constant__ CUtensorMap tm_hres, const ChainParams p) {
    using C = Cfg<D, H>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_hpre); tma_prefetch_desc(&tm_w1);
        tma_prefetch_desc(&tm_w2); tma_prefetch_desc(&tm_hpost); tma_prefetch_desc(&tm_hres);
        mbar_init(&bar[B_XFULL], 1); mbar_init(&bar[B_XFREE], 1);
        mbar_init(&bar[B_XN], kEpiWarps); mbar_init(&bar[B_ACCA], 1); mbar_init(&bar[B_HA], kEpiWarps);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar[B_ACC1 + b], 1); mbar_init(&bar[B_H1 + b], kEpiWarps); mbar_init(&bar[B_G3 + b], 1);
        }
        mbar_init(&bar[B_ACC3], 1); mbar_init(&bar[B_ACC3FREE], kEpiWarps);
        for (int s = 0; s < 4; ++s) { mbar_init(&bar[B_RFULL + s], 1); mbar_init(&bar[B_REMPTY + s], 1); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================== TMA producer
        if (lane == 0) {
            int slot = 0;
            uint32_t rphase = 0;
            auto acquire = [&](uint32_t bytes) -> uint8_t* {
                mbar_wait(&bar[B_REMPTY + slot], rphase ^ 1u);
                mbar_arrive_expect_tx(&bar[B_RFULL + slot], bytes);
                return smem + C::kOffRing + slot * kSlotBytes;
            };
            auto advance = [&]() { if (++slot == C::kSlots) { slot = 0; rphase ^= 1u; } };
            uint32_t t = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++t) {
                if (t > 0) mbar_wait(&bar[B_XFREE], (t - 1) & 1u);
                mbar_arrive_expect_tx(&bar[B_XFULL], 16384);
                tma_load_2d(smem + C::kOffX, &tm_x, &bar[B_XFULL], 0, tile * kBM);
                {   // H_pre^T [H rows x 64]
                    uint8_t* s = acquire(H * 128);
                    tma_load_2d(s, &tm_hpre, &bar[B_RFULL + slot], 0, 0);
                    advance();
                }
                auto load_w1 = [&](int j) {               // W1 rows [64 j, +64), all of K = H: KBH boxes of [64 x 64] in one slot
                    uint8_t* s = acquire(C::KBH * 8192);
                    for (int kb = 0; kb < C::KBH; ++kb) tma_load_2d(s + kb * 8192, &tm_w1, &bar[B_RFULL + slot], kb * 64, j * 64);
                    advance();
                };
                auto load_w2 = [&](int j) {               // W2 [H rows] x K columns [64 j, +64): one box of [H x 64]
                    uint8_t* s = acquire(H * 128);
                    tma_load_2d(s, &tm_w2, &bar[B_RFULL + slot], j * 64, 0);
                    advance();
                };
                for (int j = 0; j < C::NCH; ++j) {
                    load_w1(j);
                    if (j >= 1) load_w2(j - 1);
                }
                load_w2(C::NCH - 1);
                {   // H_post^T [D rows x H] as KBH blocks of [D x 64] (32 KB at D = 64), then H_res^T [D x 64] in a slot of its own
                    uint8_t* s = acquire(C::KBH * D * 128);
                    for (int kb = 0; kb < C::KBH; ++kb) tma_load_2d(s + kb * D * 128, &tm_hpost, &bar[B_RFULL + slot], kb * 64, 0);
                    advance();
                    s = acquire(D * 128);
                    tma_load_2d(s, &tm_hres, &bar[B_RFULL + slot], 0, 0);
                    advance();
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        if (lane == 0) {
            int slot = 0;
            uint32_t rphase = 0;
            auto slot_wait = [&]() -> uint32_t {
                mbar_wait(&bar[B_RFULL + slot], rphase);
                tc_fence_after();
                return base + C::kOffRing + slot * kSlotBytes;
            };
            auto slot_release = [&]() {
                umma_commit(&bar[B_REMPTY + slot]);
                if (++slot == C::kSlots) { slot = 0; rphase ^= 1u; }
            };
            // D[tmem] (+)= A[128 x 64-block at a_addr] * B[n rows x 64-block at b_addr]^T over `ksteps` 16-wide K steps
            auto mma_block = [&](uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, int n, int ksteps, bool first) {
                const uint32_t idesc = umma_idesc_bf16(kBM, n, 0, 0);
                const uint64_t ad = umma_smem_desc(a_addr, 16, 1024, kUmmaLayoutSw128);
                const uint64_t bd = umma_smem_desc(b_addr, 16, 1024, kUmmaLayoutSw128);
                for (int k = 0; k < ksteps; ++k)
                    umma_bf16_ss(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)(!(first && k == 0)));
            };
            uint32_t t = 0, n_ha = 0, n_h1[2] = {0, 0};
            const uint32_t accA = tmem_base, acc1 = tmem_base + kColAcc1;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++t) {
                // ---- G1: h0 = xn @ H_pre
                mbar_wait(&bar[B_XN], t & 1u);
                tc_fence_after();
                {
                    const uint32_t s = slot_wait();
                    mma_block(accA, base + C::kOffXN, s, H, D / 16, true);
                    slot_release();
                    umma_commit(&bar[B_ACCA]);
                }
                mbar_wait(&bar[B_HA], n_ha++ & 1u);                        // h0 is in shared memory
                tc_fence_after();
                if (t > 0) { mbar_wait(&bar[B_ACC3FREE], (t - 1) & 1u); tc_fence_after(); }   // last tile's output left acc1[0]
                auto g3 = [&](int j) {                                     // h2 accumulator += h1_j @ W2[:, chunk j]^T
                    const int b = j & 1;
                    mbar_wait(&bar[B_H1 + b], n_h1[b]++ & 1u);
                    tc_fence_after();
                    const uint32_t h1 = base + C::kOffH1 + b * 16384;
                    const uint32_t s = slot_wait();
                    mma_block(accA, h1, s, H, 4, j == 0);
                    slot_release();
                    umma_commit(&bar[B_G3 + b]);
                };
                for (int j = 0; j < C::NCH; ++j) {
                    // ---- G2 chunk j: acc1[j & 1] = h0 @ W1[chunk j]^T   (N = 64)
                    const uint32_t d = acc1 + (uint32_t)(j & 1) * 128u;
                    {
                        const uint32_t s = slot_wait();
                        for (int kb = 0; kb < C::KBH; ++kb) mma_block(d, base + C::kOffH0 + kb * 16384, s + kb * 8192, C::CW, 4, kb == 0);
                        slot_release();
                    }
                    umma_commit(&bar[B_ACC1 + (j & 1)]);
                    if (j >= 1) g3(j - 1);
                }
                g3(C::NCH - 1);
                umma_commit(&bar[B_ACCA]);                                 // h2's accumulator complete
                mbar_wait(&bar[B_HA], n_ha++ & 1u);                        // h2 is in shared memory
                tc_fence_after();
                {   // ---- G4 + G5: out accumulator = h2 @ H_post + x @ H_res
                    uint32_t s = slot_wait();
                    for (int kb = 0; kb < C::KBH; ++kb) mma_block(acc1, base + C::kOffH0 + kb * 16384, s + kb * D * 128, D, 4, kb == 0);
                    slot_release();
                    s = slot_wait();
                    mma_block(acc1, base + C::kOffX, s, D, D / 16, false);
                    slot_release();
                    umma_commit(&bar[B_XFREE]);
                    umma_commit(&bar[B_ACC3]);
                }
            }
        }
    } else {
        // ================================================================== epilogue warps
        const int q = warp & 3, part = (warp - 2) >> 2;
        const uint32_t r = (uint32_t)(q * 32 + lane);                      // tile row = tensor-memory lane of this thread
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t t = 0, n_acca = 0, n_acc1[2] = {0, 0}, n_g3[2] = {0, 0};
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++t) {
            const int64_t row = (int64_t)tile * kBM + r;
            // ---- LayerNorm_pre of the staged x tile (part 0: one thread per row)
            mbar_wait(&bar[B_XFULL], t & 1u);
            if (part == 0) {
                // three passes over the row in shared memory (sum, squared deviations, normalise): 8 values live at a time -- a
                // 576-thread CTA has 96 registers per thread, a whole D = 64 row in registers spilled
                auto load8 = [&](int g, float (&f)[8]) {
                    const uint4 v = lds128(base + C::kOffX + sw128(r, g));
                    f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
                    f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
                };
                float s = 0.f;
#pragma unroll
                for (int g = 0; g < D / 8; ++g) {
                    float f[8];
                    load8(g, f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) s += f[j];
                }
                const float mean = s * (1.0f / D);
                float qv = 0.f;
#pragma unroll
                for (int g = 0; g < D / 8; ++g) {
                    float f[8];
                    load8(g, f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; qv = fmaf(d, d, qv); }
                }
                const float inv = rsqrtf(qv * (1.0f / D) + p.eps_pre);
#pragma unroll
                for (int g = 0; g < D / 8; ++g) {
                    float f[8];
                    load8(g, f);
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = 8 * g + 2 * e;
                        o[e] = pack_bf16((f[2 * e] - mean) * inv * __ldg(p.ln_pre_w + j) + __ldg(p.ln_pre_b + j),
                                         (f[2 * e + 1] - mean) * inv * __ldg(p.ln_pre_w + j + 1) + __ldg(p.ln_pre_b + j + 1));
                    }
                    sts128(base + C::kOffXN + sw128(r, g), make_uint4(o[0], o[1], o[2], o[3]));
                }
                fence_proxy_async_smem();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[B_XN]);

            // ---- E0: h0 accumulator -> bf16 -> shared memory (K-major, swizzled)
            mbar_wait(&bar[B_ACCA], n_acca++ & 1u);
            tc_fence_after();
            for (int c = part; c < H / 16; c += 4) {
                uint32_t v[16];
                tmem_ld16(tmem_base + lane_base + (uint32_t)(c * 16), v);
                tmem_wait_ld();
                const uint32_t blk = base + C::kOffH0 + (uint32_t)(c >> 2) * 16384u;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint4 o = make_uint4(pack_bf16(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                                               pack_bf16(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                                               pack_bf16(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                                               pack_bf16(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
                    sts128(blk + sw128(r, (uint32_t)((c & 3) * 2 + g)), o);
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[B_HA]);

            // ---- E1: per 64-column chunk of h1: + b1, GELU, bf16 -> shared memory (each warp of a quadrant pair takes 32 columns)
            for (int j = 0; j < C::NCH; ++j) {
                const int b = j & 1;
                mbar_wait(&bar[B_ACC1 + b], n_acc1[b]++ & 1u);
                if (n_g3[b] > 0) mbar_wait(&bar[B_G3 + b], (n_g3[b] - 1) & 1u);    // the MMAs that read this h1 buffer last are done
                ++n_g3[b];
                tc_fence_after();
                for (int c = part; c < C::CW / 16; c += 4) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + lane_base + kColAcc1 + (uint32_t)(b * 128 + c * 16), v);
                    tmem_wait_ld();
                    const float4* b4 = reinterpret_cast<const float4*>(p.b1 + j * C::CW + c * 16);
                    const uint32_t blk = base + C::kOffH1 + (uint32_t)b * 16384u;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const float4 ba = __ldg(b4 + 2 * g), bb = __ldg(b4 + 2 * g + 1);
                        const uint4 o = make_uint4(gelu2_bf16(__uint_as_float(v[8 * g]) + ba.x, __uint_as_float(v[8 * g + 1]) + ba.y),
                                                   gelu2_bf16(__uint_as_float(v[8 * g + 2]) + ba.z, __uint_as_float(v[8 * g + 3]) + ba.w),
                                                   gelu2_bf16(__uint_as_float(v[8 * g + 4]) + bb.x, __uint_as_float(v[8 * g + 5]) + bb.y),
                                                   gelu2_bf16(__uint_as_float(v[8 * g + 6]) + bb.z, __uint_as_float(v[8 * g + 7]) + bb.w));
                        sts128(blk + sw128(r, (uint32_t)((c & 3) * 2 + g)), o);
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar[B_H1 + b]);
            }

            // ---- E2: h2 accumulator: + b2, GELU, bf16 -> shared memory (over h0, which every G2 MMA has finished reading)
            mbar_wait(&bar[B_ACCA], n_acca++ & 1u);
            tc_fence_after();
            for (int c = part; c < H / 16; c += 4) {
                uint32_t v[16];
                tmem_ld16(tmem_base + lane_base + (uint32_t)(c * 16), v);
                tmem_wait_ld();
                const float4* b4 = reinterpret_cast<const float4*>(p.b2 + c * 16);
                const uint32_t blk = base + C::kOffH0 + (uint32_t)(c >> 2) * 16384u;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const float4 ba = __ldg(b4 + 2 * g), bb = __ldg(b4 + 2 * g + 1);
                    const uint4 o = make_uint4(gelu2_bf16(__uint_as_float(v[8 * g]) + ba.x, __uint_as_float(v[8 * g + 1]) + ba.y),
                                               gelu2_bf16(__uint_as_float(v[8 * g + 2]) + ba.z, __uint_as_float(v[8 * g + 3]) + ba.w),
                                               gelu2_bf16(__uint_as_float(v[8 * g + 4]) + bb.x, __uint_as_float(v[8 * g + 5]) + bb.y),
                                               gelu2_bf16(__uint_as_float(v[8 * g + 6]) + bb.z, __uint_as_float(v[8 * g + 7]) + bb.w));
                    sts128(blk + sw128(r, (uint32_t)((c & 3) * 2 + g)), o);
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[B_HA]);

            // ---- E3: LayerNorm_post over the D output columns, store the row (part 0)
            mbar_wait(&bar[B_ACC3], t & 1u);
            tc_fence_after();
            if (part == 0) {
                // three passes over the accumulator row in tensor memory, 16 columns at a time (register budget, as above)
                const uint32_t ta = tmem_base + lane_base + kColAcc1;
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < D / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld16(ta + (uint32_t)(c * 16), v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) s += __uint_as_float(v[j]);
                }
                const float mean = s * (1.0f / D);
                float qv = 0.f;
#pragma unroll
                for (int c = 0; c < D / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld16(ta + (uint32_t)(c * 16), v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) { const float d = __uint_as_float(v[j]) - mean; qv = fmaf(d, d, qv); }
                }
                const float inv = rsqrtf(qv * (1.0f / D) + p.eps_post);
#pragma unroll
                for (int c = 0; c < D / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld16(ta + (uint32_t)(c * 16), v);
                    tmem_wait_ld();
                    float y[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        y[j] = (__uint_as_float(v[j]) - mean) * inv * __ldg(p.ln_post_w + c * 16 + j) + __ldg(p.ln_post_b + c * 16 + j);
                    if (row < p.T) {
                        if (p.out_f32) {
                            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * D + c * 16);
#pragma unroll
                            for (int j = 0; j < 4; ++j) o[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
                        } else {
                            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * D + c * 16);
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                o[j] = make_uint4(pack_bf16(y[8 * j], y[8 * j + 1]), pack_bf16(y[8 * j + 2], y[8 * j + 3]),
                                                  pack_bf16(y[8 * j + 4], y[8 * j + 5]), pack_bf16(y[8 * j + 6], y[8 * j + 7]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[B_ACC3FREE]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int D, int H>
int launch_chain(const void* x, const void* hpre_t, const void* w1, const void* w2, const void* hpost_t, const void* hres_t,
                 const ChainParams& p, cudaStream_t stream) {
    using C = Cfg<D, H>;
    CUtensorMap tx, thpre, tw1, tw2, thpost, thres;
    int rc;
    if ((rc = make_tmap_bf16_2d_ld(&tx, x, (uint64_t)p.T, D, D, kBM))) return rc;
    if ((rc = make_tmap_bf16_2d_ld(&thpre, hpre_t, H, D, D, H))) return rc;
    if ((rc = make_tmap_bf16_2d_ld(&tw1, w1, 2 * H, H, H, C::CW))) return rc;
    if ((rc = make_tmap_bf16_2d_ld(&tw2, w2, H, 2 * H, 2 * H, H))) return rc;
    if ((rc = make_tmap_bf16_2d_ld(&thpost, hpost_t, D, H, H, D))) return rc;
    if ((rc = make_tmap_bf16_2d_ld(&thres, hres_t, D, D, D, D))) return rc;
    HVS_SET_MAX_SMEM((k2_chain_kernel<D, H>), C::kSmemBytes);
    const int sms = sm_count();
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    timer_begin(6, stream);
    k2_chain_kernel<D, H><<<grid, kThreads, C::kSmemBytes, stream>>>(tx, thpre, tw1, tw2, thpost, thres, p);
    timer_end(6, stream);
    count_launch();
    return launch_status();
}

}  // namespace
}  // namespace hvs

extern "C" int hvs_mhc_module_fwd_supported(int D, int H) { return (D == 32 && H == 128) || (D == 64 && H == 256); }

extern "C" int hvs_mhc_module_fwd(const void* x, const void* h_pre_t, const void* w1, const float* b1, const void* w2,
                                  const float* b2, const void* h_post_t, const void* h_res_t, const float* ln_pre_w,
                                  const float* ln_pre_b, float ln_pre_eps, const float* ln_post_w, const float* ln_post_b,
                                  float ln_post_eps, void* out, int out_dtype, int64_t T, int D, int H, void* stream_) {
    using namespace hvs;
    if (T < 0) return HVS_ERR_BAD_ARG;
    if (!hvs_mhc_module_fwd_supported(D, H)) return HVS_ERR_UNSUPPORTED;
    if (T == 0) return HVS_OK;
    if (!x || !h_pre_t || !w1 || !b1 || !w2 || !b2 || !h_post_t || !h_res_t || !ln_pre_w || !ln_pre_b || !ln_post_w || !ln_post_b || !out)
        return HVS_ERR_BAD_ARG;
    if (out_dtype != HVS_DTYPE_F32 && out_dtype != HVS_DTYPE_BF16) return HVS_ERR_UNSUPPORTED;
    if (T >= ((int64_t)1 << 31) - kBM) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(h_pre_t) | reinterpret_cast<uintptr_t>(w1) |
         reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(h_post_t) | reinterpret_cast<uintptr_t>(h_res_t) |
         reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(b2)) & 15)
        return HVS_ERR_ALIGNMENT;
    ChainParams p{};
    p.b1 = b1; p.b2 = b2; p.ln_pre_w = ln_pre_w; p.ln_pre_b = ln_pre_b; p.ln_post_w = ln_post_w; p.ln_post_b = ln_post_b;
    p.out = out; p.T = T; p.num_tiles = (int)((T + kBM - 1) / kBM); p.out_f32 = out_dtype == HVS_DTYPE_F32;
    p.eps_pre = ln_pre_eps; p.eps_post = ln_post_eps;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (D == 32) return launch_chain<32, 128>(x, h_pre_t, w1, w2, h_post_t, h_res_t, p, stream);
    return launch_chain<64, 256>(x, h_pre_t, w1, w2, h_post_t, h_res_t, p, stream);
}
