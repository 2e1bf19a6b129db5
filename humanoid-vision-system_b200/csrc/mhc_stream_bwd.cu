// K1 backward (F = identity), coefficients recomputed from x -- nothing is saved by the forward.
//
// Kernel B1 (per token, HBM-bound: x and dy in, dx out; every byte crosses HBM once).  The per-token
// Sinkhorn forward + reverse sweep is a ~14k-cycle dependent chain, several times longer than a tile can
// afford to occupy a shared-memory stage.  So after pass 1 each 8-token tile is PARKED IN TENSOR MEMORY
// (256 KB per SM, unused otherwise: this kernel has no tcgen05.mma) -- every worker thread stores its own
// 32 registers of x and dy with tcgen05.st and gets them back with tcgen05.ld -- and the shared-memory
// stage is recycled at once.  Four tiles can be parked while two coefficient warps work on them:
//   P1 workers (16 warps, split-K)   raw = x.W on the warp MMA path (W = bf16(scale*phi) in registers),
//                                    sum x^2 and the per-token 4x4  G = dy x^T  as MMAs on the smem tile,
//                                    park x, dy in TMEM, release the stage
//                                    and fixed-order sum of the 16 split-K partials into a per-token record
//   3 coefficient warps (lane=token, forward gates + Sinkhorn in packed fp32x2 registers, then the exact
//     one tile each, round robin)    reverse sweep through all iterations -> dlogits, e = d raw, kappa, M^T
//   P3 workers (4 tiles later)       dx = M^T dy  +  e W^T (MMA; W^T by movmatrix from the same registers)
//                                        + kappa x from the parked registers, one rounding to bf16, staged in
//                                    shared memory and TMA-stored
//   The coefficient warps also emit E[T,24] (fp32) and per-CTA partial sums of dbias / dalpha.
// Kernel B2: dW = x^T E on the warp MMA path (x re-read once; E split into two bf16 terms), per-CTA
//   partials; finalize: dphi = scale * dW, dscale = sum_k phi * dW, dbias, dalpha (fixed order).
//
// Oracle: autograd through oracle/mhc_ref.py::stream_mhc_forward (reference primitives
// src/models/manifold_layers.py:56-77, :213-216, :449-456).
#include "common.cuh"
#include "mhc_stream_shared.cuh"
#include "ptx_sm100.cuh"

namespace hvs {
namespace {

// ------------------------------------------------------------------------------------------------ B1
constexpr int kTok = 8;                               // tokens per tile
constexpr int kSlots = 4;                             // tiles parked in TMEM at a time (4 x 128 columns)
constexpr int kWorkers = 16;
constexpr int kWorkerThreads = kWorkers * 32;
constexpr int kThreads = (kWorkers + 4) * 32;         // + coefficient warp 0, producer, coefficient warps 1, 2
constexpr int kCoefWarps = 3;
constexpr int kWorkerRegs = 104, kRoleRegs = 64;
constexpr int kStages = 2;                            // TMA landing stages (x | dy), held only until P1 is done
constexpr int kHalfBytes = kTok * kRowBytes;          // 32 KB: x tile or dy tile
constexpr int kStageBytes = 2 * kHalfBytes;
constexpr int kDxBufs = 1;                            // dx staging for the TMA store
constexpr int kBoxBytes = 32 * 128;                   // TMA box: 32 rows (8 tokens x 4 streams) x 64 bf16
// per-token record (fp32 words): P1 writes raw[0..23] ss[24] G[28..43]; the coefficient warp overwrites it in
// place with e as bf16 pairs [0..11], kappa [12], M^T [28..43] for P3
constexpr int kRec = 44, kRecSS = 24, kRecG = 28, kRecKappa = 12;
constexpr int kMaxIters = 24;                         // normalisers of every iteration are kept in shared memory
constexpr uint32_t kTmemCols = 512;

constexpr int kOffDx = kStages * kStageBytes;
constexpr int kOffPart = kOffDx + kDxBufs * kHalfBytes;
constexpr int kOffRec = kOffPart + kWorkers * kTok * kRec * 4;
constexpr int kSkWords = kTok * kMaxIters * 8;         // normaliser scratch per coefficient warp: [token][iter][row d x4 | col d x4]
constexpr int kOffSk = kOffRec + kSlots * kTok * kRec * 4;
constexpr int kOffBar = kOffSk + kCoefWarps * kSkWords * 4;
constexpr int kOffTmem = kOffBar + (2 * kStages + 2 * kDxBufs) * 8;
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kOffBar % 8 == 0, "mbarrier alignment");
static_assert(kSmemBytes <= 232448, "shared memory budget");

constexpr int kBarW1 = 1, kBarW2 = 2, kBarRed = 4 /*,5,6*/, kBarCoef = 7 /*,8,9*/;
constexpr int kAccum = kL + 3;                        // dbias[24], dalpha[3]

struct BwdParams {
    const float* phi;
    const float* bias;
    const float* alpha;
    const float* scale;
    float* e_out;          // [T,24] fp32
    float* cta_accum;      // [grid, 3, 27]
    int64_t T;
    int num_tiles;
    int sk_iters;
    float eps_rms, eps_sk;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t v) {
    uint32_t r;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// ---- tensor memory as a register parking lot (tcgen05.st / tcgen05.ld, 32 lanes x 32 bit x N columns)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
mhc_stream_bwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                      const __grid_constant__ CUtensorMap tmap_dx, const BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023u) __trap();
    float* part = reinterpret_cast<float*>(smem + kOffPart);
    float* rec = reinterpret_cast<float*>(smem + kOffRec);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kOffBar);     // TMA load landed
    uint64_t* bar_free = bar_full + kStages;                              // P1 done with the stage
    uint64_t* bar_dx_full = bar_free + kStages;                           // dx staged by the workers
    uint64_t* bar_dx_free = bar_dx_full + kDxBufs;                        // staged dx drained by the TMA store
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_local = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_free[s], kWorkerThreads); }
        for (int s = 0; s < kDxBufs; ++s) { mbar_init(&bar_dx_full[s], kWorkerThreads); mbar_init(&bar_dx_free[s], 1); }
        fence_mbar_init();
    }
    if (warp == kWorkers + 1) tmem_alloc(tmem_slot, kTmemCols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= kWorkers) {
      reg_dealloc<kRoleRegs>();
      if (warp == kWorkers + 1) {
        // ===================================================== producer / store (one lane)
        if (lane == 0) {
            tma_prefetch_desc(&tmap_x);
            tma_prefetch_desc(&tmap_dy);
            tma_prefetch_desc(&tmap_dx);
            auto load_tile = [&](int it) {
                const int s = it % kStages;
                const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * (kTok * kN);
                uint8_t* st = smem + s * kStageBytes;
                mbar_arrive_expect_tx(&bar_full[s], kStageBytes);
#pragma unroll
                for (int cb = 0; cb < kC / 64; ++cb) {
                    tma_load_2d(st + cb * kBoxBytes, &tmap_x, &bar_full[s], cb * 64, row0);
                    tma_load_2d(st + kHalfBytes + cb * kBoxBytes, &tmap_dy, &bar_full[s], cb * 64, row0);
                }
            };
            for (int it = 0; it < kStages && it < n_local; ++it) load_tile(it);
            for (int j = 0; j < n_local + kSlots; ++j) {
                const int k = j - kSlots;                       // tile whose dx is staged in this iteration
                if (k >= 0) {
                    const int b = k % kDxBufs;
                    mbar_wait(&bar_dx_full[b], (k / kDxBufs) & 1);
                    const int row0 = ((int)blockIdx.x + k * (int)gridDim.x) * (kTok * kN);
#pragma unroll
                    for (int cb = 0; cb < kC / 64; ++cb)
                        tma_store_2d(&tmap_dx, smem + kOffDx + b * kHalfBytes + cb * kBoxBytes, cb * 64, row0);
                    bulk_commit();
                    bulk_wait_read<0>();
                    mbar_arrive(&bar_dx_free[b]);
                }
                if (j < n_local && j + kStages < n_local) {
                    mbar_wait(&bar_free[j % kStages], (j / kStages) & 1);     // P1(j) has released its stage
                    load_tile(j + kStages);
                }
            }
            bulk_wait<0>();
        }
      } else {
        // ===================================================== coefficient warps 0..2 (warps 16, 18, 19): tile k
        // goes to warp k % 3.  Four lanes per token: lane (tk, i) owns row i of the token's 4x4 block as two
        // packed fp32x2 registers A = (p_i0,p_i1), B = (p_i2,p_i3) and gate i of H_pre / H_post.  Row sums are
        // local, column sums take two xor-shuffles inside the 4-lane group.
        const int cw = warp == kWorkers ? 0 : warp - (kWorkers + 1);
        const int tk = lane >> 2, i = lane & 3, gbase = lane & ~3;
        const float b_pre = __ldg(p.bias + i), b_post = __ldg(p.bias + kN + i);
        const float4 b_res = __ldg(reinterpret_cast<const float4*>(p.bias + 2 * kN) + i);
        const float a_pre = __ldg(p.alpha + 0), a_post = __ldg(p.alpha + 1), a_res = __ldg(p.alpha + 2);
        const float eps = p.eps_sk;
        const u64 eps2 = pk2(eps, eps);
        float acc_b[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // dbias: pre_i, post_i, res_i0..3
        float acc_a[3] = {0.f, 0.f, 0.f};                  // dalpha terms of this lane
        float* skl = reinterpret_cast<float*>(smem + kOffSk) + cw * kSkWords + tk * (kMaxIters * 8);
        auto gsum2 = [](u64 v) {                           // sum over the 4 lanes of a group, both halves
            float a, b;
            upk2(v, a, b);
            a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2);
            return pk2(a, b);
        };
        for (int it = cw; it < n_local; it += kCoefWarps) {
            bar_sync(kBarRed + cw, kWorkerThreads + 32);
            float* r = rec + ((it % kSlots) * kTok + tk) * kRec;
            const float inv_rms = __fdiv_rn(1.0f, __fsqrt_rn(fmaf(r[kRecSS], 1.0f / kRow, p.eps_rms)));
            const float raw_pre = r[i], raw_post = r[kN + i];
            const float4 raw_res = *reinterpret_cast<const float4*>(r + 2 * kN + 4 * i);
            const float4 gq = *reinterpret_cast<const float4*>(r + kRecG + 4 * i);          // row i of G = dy x^T
            const float hpre = sigmoid_f32(fmaf(a_pre, raw_pre * inv_rms, b_pre));
            const float hpost = 2.0f * sigmoid_f32(fmaf(a_post, raw_post * inv_rms, b_post));
            u64 A, B;
            {
                const float l0 = fmaf(a_res, raw_res.x * inv_rms, b_res.x), l1 = fmaf(a_res, raw_res.y * inv_rms, b_res.y);
                const float l2 = fmaf(a_res, raw_res.z * inv_rms, b_res.z), l3 = fmaf(a_res, raw_res.w * inv_rms, b_res.w);
                const float mx = fmaxf(fmaxf(l0, l1), fmaxf(l2, l3));
                const float e0 = fast_exp(l0 - mx), e1 = fast_exp(l1 - mx), e2 = fast_exp(l2 - mx), e3 = fast_exp(l3 - mx);
                const float r4 = 4.0f * rcp_approx((e0 + e1) + (e2 + e3));
                A = pk2(e0 * r4, e1 * r4);
                B = pk2(e2 * r4, e3 * r4);
            }
            // ---- forward Sinkhorn, the normalisers are kept for the reverse sweep
            for (int k = 0; k < p.sk_iters; ++k) {
                float sa, sb;
                upk2(add2(A, B), sa, sb);
                const float dr = (sa + sb) + eps;
                const float rr = rcp_approx(dr);
                const u64 rr2 = pk2(rr, rr);
                A = mul2(A, rr2); B = mul2(B, rr2);
                const u64 cA = add2(gsum2(A), eps2), cB = add2(gsum2(B), eps2);
                float c0, c1, c2, c3;
                upk2(cA, c0, c1); upk2(cB, c2, c3);
                A = mul2(A, pk2(rcp_approx(c0), rcp_approx(c1)));
                B = mul2(B, pk2(rcp_approx(c2), rcp_approx(c3)));
                skl[k * 8 + i] = dr;
                if (i == 0) *reinterpret_cast<float4*>(skl + k * 8 + 4) = make_float4(c0, c1, c2, c3);
            }
            // ---- M = P + hpost (x) hpre needs every H_pre of the token; gate gradients from G
            const float h0 = __shfl_sync(0xffffffffu, hpre, gbase + 0), h1 = __shfl_sync(0xffffffffu, hpre, gbase + 1);
            const float h2 = __shfl_sync(0xffffffffu, hpre, gbase + 2), h3 = __shfl_sync(0xffffffffu, hpre, gbase + 3);
            float p0, p1, p2, p3;
            upk2(A, p0, p1); upk2(B, p2, p3);
            const float dhpost = fmaf(gq.w, h3, fmaf(gq.z, h2, fmaf(gq.y, h1, gq.x * h0)));
            float t0 = gq.x * hpost, t1 = gq.y * hpost, t2 = gq.z * hpost, t3 = gq.w * hpost;   // dhpre[j] = sum_i G[i][j] hpost[i]
            t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
            t2 += __shfl_xor_sync(0xffffffffu, t2, 1); t3 += __shfl_xor_sync(0xffffffffu, t3, 1);
            t0 += __shfl_xor_sync(0xffffffffu, t0, 2); t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
            t2 += __shfl_xor_sync(0xffffffffu, t2, 2); t3 += __shfl_xor_sync(0xffffffffu, t3, 2);
            const float dhpre = i == 0 ? t0 : i == 1 ? t1 : i == 2 ? t2 : t3;
            const float dl_pre = dhpre * hpre * (1.0f - hpre);
            const float dl_post = dhpost * hpost * (1.0f - 0.5f * hpost);
            __syncwarp();                                   // every lane of the group has read its G row and raw logits
            r[kRecG + 0 * 4 + i] = fmaf(hpost, h0, p0);     // M^T[j][i] = M[i][j] replaces G in the record
            r[kRecG + 1 * 4 + i] = fmaf(hpost, h1, p1);
            r[kRecG + 2 * 4 + i] = fmaf(hpost, h2, p2);
            r[kRecG + 3 * 4 + i] = fmaf(hpost, h3, p3);
            __syncwarp();                                   // normalisers written by lane 0 of the group are visible
            // ---- exact reverse sweep through the iterations (dP = G)
            u64 Da = pk2(gq.x, gq.y), Db = pk2(gq.z, gq.w);
            for (int k = p.sk_iters - 1; k >= 0; --k) {
                const float4 cd = *reinterpret_cast<const float4*>(skl + k * 8 + 4);
                const float dr = skl[k * 8 + i];
                // column step y = x / c:  dx = (dy - sum_rows dy*y) / c ;  x = y * c
                const u64 qA = gsum2(mul2(Da, A)), qB = gsum2(mul2(Db, B));
                const float r0 = rcp_approx(cd.x), r1 = rcp_approx(cd.y), r2 = rcp_approx(cd.z), r3 = rcp_approx(cd.w);
                Da = fma2(Da, pk2(r0, r1), mul2(qA, pk2(-r0, -r1)));
                Db = fma2(Db, pk2(r2, r3), mul2(qB, pk2(-r2, -r3)));
                A = mul2(A, pk2(cd.x, cd.y)); B = mul2(B, pk2(cd.z, cd.w));
                // row step y = x / dr
                float qa, qb;
                upk2(fma2(Db, B, mul2(Da, A)), qa, qb);
                const float rr = rcp_approx(dr);
                const float nq = -(qa + qb) * rr;
                const u64 rr2 = pk2(rr, rr), nq2 = pk2(nq, nq), dd2 = pk2(dr, dr);
                Da = fma2(Da, rr2, nq2); Db = fma2(Db, rr2, nq2);
                A = mul2(A, dd2); B = mul2(B, dd2);
            }
            // softmax * 4 backward (A,B are back at the softmax output): dl = s * (d - sum(d*s)/4)
            float dl0, dl1, dl2, dl3;
            {
                float qa, qb;
                upk2(fma2(Db, B, mul2(Da, A)), qa, qb);
                const float nqs = -0.25f * (qa + qb);
                const u64 nq2 = pk2(nqs, nqs);
                upk2(mul2(A, add2(Da, nq2)), dl0, dl1);
                upk2(mul2(B, add2(Db, nq2)), dl2, dl3);
            }
            // ---- e = d raw, kappa (RMSNorm backward), dbias / dalpha terms
            acc_b[0] += dl_pre; acc_b[1] += dl_post; acc_b[2] += dl0; acc_b[3] += dl1; acc_b[4] += dl2; acc_b[5] += dl3;
            acc_a[0] = fmaf(dl_pre, raw_pre * inv_rms, acc_a[0]);
            acc_a[1] = fmaf(dl_post, raw_post * inv_rms, acc_a[1]);
            acc_a[2] += fmaf(dl3, raw_res.w, fmaf(dl2, raw_res.z, fmaf(dl1, raw_res.y, dl0 * raw_res.x))) * inv_rms;
            const float e_pre = a_pre * dl_pre * inv_rms, e_post = a_post * dl_post * inv_rms;
            const float e0 = a_res * dl0 * inv_rms, e1 = a_res * dl1 * inv_rms;
            const float e2 = a_res * dl2 * inv_rms, e3 = a_res * dl3 * inv_rms;
            // d inv_rms = sum_k e_k raw_k / inv_rms ;  kappa = -d inv_rms * inv_rms^3 / N
            float dsum = e_pre * raw_pre + e_post * raw_post + (e0 * raw_res.x + e1 * raw_res.y + e2 * raw_res.z + e3 * raw_res.w);
            dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
            dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
            __nv_bfloat16* eb = reinterpret_cast<__nv_bfloat16*>(r);        // e as bf16 pairs over raw[0..11]
            eb[i] = __float2bfloat16_rn(e_pre);
            eb[kN + i] = __float2bfloat16_rn(e_post);
            *reinterpret_cast<uint2*>(eb + 2 * kN + 4 * i) = make_uint2(pack_bf16(e0, e1), pack_bf16(e2, e3));
            if (i == 0) r[kRecKappa] = -dsum * inv_rms * inv_rms * (1.0f / kRow);
            const int64_t tok = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kTok + tk;
            if (tok < p.T) {
                float* o = p.e_out + tok * kL;
                o[i] = e_pre;
                o[kN + i] = e_post;
                *reinterpret_cast<float4*>(o + 2 * kN + 4 * i) = make_float4(e0, e1, e2, e3);
            }   // padded rows have dy = 0, hence G = 0 and every dl = 0: they add nothing to the sums above
            __threadfence_block();
            bar_arrive(kBarCoef + cw, kWorkerThreads + 32);
        }
        // fold the 8 token groups of the warp (lanes with equal i) in a fixed order, then the group for dalpha
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < 6; ++k) acc_b[k] += __shfl_xor_sync(0xffffffffu, acc_b[k], o);
#pragma unroll
            for (int k = 0; k < 3; ++k) acc_a[k] += __shfl_xor_sync(0xffffffffu, acc_a[k], o);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            acc_a[k] += __shfl_xor_sync(0xffffffffu, acc_a[k], 1);
            acc_a[k] += __shfl_xor_sync(0xffffffffu, acc_a[k], 2);
        }
        if (lane < 4) {
            float* o = p.cta_accum + ((size_t)blockIdx.x * kCoefWarps + cw) * kAccum;
            o[i] = acc_b[0];
            o[kN + i] = acc_b[1];
            o[2 * kN + 4 * i + 0] = acc_b[2]; o[2 * kN + 4 * i + 1] = acc_b[3];
            o[2 * kN + 4 * i + 2] = acc_b[4]; o[2 * kN + 4 * i + 3] = acc_b[5];
            if (i == 0) { o[kL] = acc_a[0]; o[kL + 1] = acc_a[1]; o[kL + 2] = acc_a[2]; }
        }
      }
    } else {
        // ===================================================== worker warps
        reg_alloc<kWorkerRegs>();
        const int w = warp, g = lane >> 2, t = lane & 3;
        uint32_t bfrag[kN][2][3][2];
#pragma unroll
        for (int j = 0; j < kN; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k0 = j * kC + 32 * w + 8 * t + 4 * q;
                float sc[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) sc[e] = __ldg(p.scale + k0 + e);
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    const int col = nt * 8 + g;
                    float f[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) f[e] = __ldg(p.phi + (size_t)(k0 + e) * kL + col) * sc[e];
                    bfrag[j][q][nt][0] = pack_bf16(f[0], f[1]);
                    bfrag[j][q][nt][1] = pack_bf16(f[2], f[3]);
                }
            }
        const int cb = w >> 1, hh = w & 1;
        uint32_t off[kN];                                   // own 16-byte chunk of (token g, stream j)
#pragma unroll
        for (int j = 0; j < kN; ++j) {
            const int row = g * kN + j;
            off[j] = cb * kBoxBytes + row * 128 + (((4 * hh + t) ^ (row & 7)) << 4);
        }
        const int lm = lane >> 3, lr = lane & 7;            // ldmatrix: matrix index / row of an x4 load
        const uint32_t stage0 = smem_u32(smem);
        // this thread's parking columns: TMEM lane = 32*(warp%4) + lane (the only quadrant the warp may touch),
        // 32 columns per (slot, warp/4): x words [0,16), dy words [16,32)
        const uint32_t tm_thread = tmem_base + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)((w >> 2) * 32);

        for (int j = 0; j < n_local + kSlots; ++j) {
            const int k = j - kSlots;
            if (k >= 0) {
                // ============ P3: dx for token g of tile k, operands back from tensor memory
                bar_sync(kBarCoef + k % kCoefWarps, kWorkerThreads + 32);                // the tile's coefficients are ready
                const float* c = rec + ((k % kSlots) * kTok + g) * kRec;
                const uint32_t* ew = reinterpret_cast<const uint32_t*>(c);
                const uint32_t ea0 = ew[t], ea2 = ew[t + 4], eb0 = ew[t + 8];     // e[2t..], e[2t+8..], e[2t+16..]
                const float kappa = c[kRecKappa];
                const uint32_t tm = tm_thread + (uint32_t)((k % kSlots) * 128);
                uint32_t xr[16], dyr[16];
                tmem_ld16(tm, xr);
                tmem_ld16(tm + 16, dyr);
                tmem_wait_ld();
                const int b = k % kDxBufs;
                mbar_wait(&bar_dx_free[b], ((k / kDxBufs) & 1) ^ 1);     // staging buffer drained (first use passes)
                const uint32_t dbase = stage0 + kOffDx + b * kHalfBytes;
#pragma unroll
                for (int jj = 0; jj < kN; ++jj) {
                    const float4 mt = *reinterpret_cast<const float4*>(c + kRecG + 4 * jj);       // M[0..3][jj]
                    uint32_t out[4];
#pragma unroll
                    for (int qq = 0; qq < 2; ++qq)
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            // dx_proj for K index (jj, 32w + 8t + 4qq + 2rr + {0,1}) of token g: e . W^T
                            float acc[4] = {0.f, 0.f, 0.f, 0.f};
                            mma_bf16_16816(acc, ea0, 0u, ea2, 0u, movmatrix_trans(bfrag[jj][qq][0][rr]), movmatrix_trans(bfrag[jj][qq][1][rr]));
                            mma_bf16_16816(acc, eb0, 0u, 0u, 0u, movmatrix_trans(bfrag[jj][qq][2][rr]), 0u);
                            const int e = 2 * qq + rr;
                            float lo = fmaf(kappa, bf16lo(xr[4 * jj + e]), acc[0]);
                            float hi = fmaf(kappa, bf16hi(xr[4 * jj + e]), acc[1]);
                            lo = fmaf(mt.x, bf16lo(dyr[e]), lo);      hi = fmaf(mt.x, bf16hi(dyr[e]), hi);
                            lo = fmaf(mt.y, bf16lo(dyr[4 + e]), lo);  hi = fmaf(mt.y, bf16hi(dyr[4 + e]), hi);
                            lo = fmaf(mt.z, bf16lo(dyr[8 + e]), lo);  hi = fmaf(mt.z, bf16hi(dyr[8 + e]), hi);
                            lo = fmaf(mt.w, bf16lo(dyr[12 + e]), lo); hi = fmaf(mt.w, bf16hi(dyr[12 + e]), hi);
                            out[e] = pack_bf16(lo, hi);
                        }
                    sts128(dbase + off[jj], make_uint4(out[0], out[1], out[2], out[3]));
                }
                fence_proxy_async_smem();
                mbar_arrive(&bar_dx_full[b]);
            }
            if (j < n_local) {
                // ============ P1: raw^T = W^T x^T (tokens are the MMA N), sum x^2 on the diagonal of x x^T
                const int s = j % kStages;
                const uint32_t sbase = stage0 + s * kStageBytes;
                const uint32_t tm = tm_thread + (uint32_t)((j % kSlots) * 128);
                mbar_wait(&bar_full[s], (j / kStages) & 1);
                float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, accs[4] = {0.f, 0.f, 0.f, 0.f};
                {
                    uint32_t xr[16];
#pragma unroll
                    for (int jj = 0; jj < kN; ++jj) {
                        const uint4 xv = lds128(sbase + off[jj]);
                        xr[4 * jj] = xv.x; xr[4 * jj + 1] = xv.y; xr[4 * jj + 2] = xv.z; xr[4 * jj + 3] = xv.w;
                        mma_bf16_16816(acc0, bfrag[jj][0][0][0], bfrag[jj][0][1][0], bfrag[jj][0][0][1], bfrag[jj][0][1][1], xv.x, xv.y);
                        mma_bf16_16816(acc1, bfrag[jj][0][2][0], 0u, bfrag[jj][0][2][1], 0u, xv.x, xv.y);
                        mma_bf16_16816(accs, xv.x, 0u, xv.y, 0u, xv.x, xv.y);
                        mma_bf16_16816(acc0, bfrag[jj][1][0][0], bfrag[jj][1][1][0], bfrag[jj][1][0][1], bfrag[jj][1][1][1], xv.z, xv.w);
                        mma_bf16_16816(acc1, bfrag[jj][1][2][0], 0u, bfrag[jj][1][2][1], 0u, xv.z, xv.w);
                        mma_bf16_16816(accs, xv.z, 0u, xv.w, 0u, xv.z, xv.w);
                    }
                    tmem_st16(tm, xr);                                   // park x
                }
                {
                    uint32_t dyr[16];
#pragma unroll
                    for (int ii = 0; ii < kN; ++ii) {
                        const uint4 v = lds128(sbase + kHalfBytes + off[ii]);
                        dyr[4 * ii] = v.x; dyr[4 * ii + 1] = v.y; dyr[4 * ii + 2] = v.z; dyr[4 * ii + 3] = v.w;
                    }
                    tmem_st16(tm + 16, dyr);                             // park dy
                }
                // G = dy x^T per token: rows (token, i) of dy against rows (token, j) of x, block diagonal
                float gacc[2][2][4];
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) gacc[m][h2][0] = gacc[m][h2][1] = gacc[m][h2][2] = gacc[m][h2][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    uint32_t bx[4][2];
#pragma unroll
                    for (int nn = 0; nn < 4; nn += 2) {
                        const int row = 8 * nn + (lm >> 1) * 8 + lr;
                        const int chunk = 4 * hh + 2 * ks + (lm & 1);
                        ldmatrix_x4(sbase + cb * kBoxBytes + row * 128 + ((chunk ^ (row & 7)) << 4),
                                    bx[nn][0], bx[nn][1], bx[nn + 1][0], bx[nn + 1][1]);
                    }
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        uint32_t a0, a1, a2, a3;
                        const int row = 16 * m + (lm & 1) * 8 + lr;
                        const int chunk = 4 * hh + 2 * ks + (lm >> 1);
                        ldmatrix_x4(sbase + kHalfBytes + cb * kBoxBytes + row * 128 + ((chunk ^ (row & 7)) << 4), a0, a1, a2, a3);
                        mma_bf16_16816(gacc[m][0], a0, a1, a2, a3, bx[2 * m][0], bx[2 * m][1]);
                        mma_bf16_16816(gacc[m][1], a0, a1, a2, a3, bx[2 * m + 1][0], bx[2 * m + 1][1]);
                    }
                }
                tmem_wait_st();
                mbar_arrive(&bar_free[s]);                        // every read of the stage has landed in registers
                bar_sync(kBarW1, kWorkerThreads);                 // the previous tile's reduction has read `part`
                {
                    float* pw = part + (size_t)w * kTok * kRec;
                    pw[(2 * t) * kRec + g] = acc0[0];      pw[(2 * t + 1) * kRec + g] = acc0[1];
                    pw[(2 * t) * kRec + 8 + g] = acc0[2];  pw[(2 * t + 1) * kRec + 8 + g] = acc0[3];
                    pw[(2 * t) * kRec + 16 + g] = acc1[0]; pw[(2 * t + 1) * kRec + 16 + g] = acc1[1];
                    if (t == (g >> 1)) pw[g * kRec + kRecSS] = accs[g & 1];       // diagonal of x x^T
                    if ((g >> 2) == (t >> 1)) {
                        const int i = g & 3, jq = 2 * (t & 1);
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            const int tokA = 4 * m + (g >> 2), tokB = 4 * m + 2 + (g >> 2);
                            *reinterpret_cast<float2*>(pw + tokA * kRec + kRecG + 4 * i + jq) = make_float2(gacc[m][0][0], gacc[m][0][1]);
                            *reinterpret_cast<float2*>(pw + tokB * kRec + kRecG + 4 * i + jq) = make_float2(gacc[m][1][2], gacc[m][1][3]);
                        }
                    }
                }
                bar_sync(kBarW2, kWorkerThreads);
                // fixed-order (tree) sum of the 16 split-K partials into the tile's record slot
                if (threadIdx.x < kTok * kRec) {
                    const int col = threadIdx.x % kRec;
                    if (!(col > kRecSS && col < kRecG)) {
                        float v[kWorkers];
#pragma unroll
                        for (int ww = 0; ww < kWorkers; ++ww) v[ww] = part[ww * kTok * kRec + threadIdx.x];
#pragma unroll
                        for (int st = 1; st < kWorkers; st <<= 1)
#pragma unroll
                            for (int ww = 0; ww < kWorkers; ww += 2 * st) v[ww] += v[ww + st];
                        rec[(j % kSlots) * kTok * kRec + threadIdx.x] = v[0];
                    }
                }
                __threadfence_block();
                bar_arrive(kBarRed + j % kCoefWarps, kWorkerThreads + 32);
            }
        }
    }
    // tensor memory is released by the warp that allocated it, after every parked register has been read back
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kWorkers + 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ B2: dW = x^T E
constexpr int kDwTok = 16;
constexpr int kDwThreads = 512;                       // 16 MMA warps; thread 0 also issues the TMA loads
constexpr int kDwStages = 3;
constexpr int kDwStageBytes = kDwTok * kRowBytes;     // 64 KB, 32 boxes of [16 tokens x 64 cols]
constexpr int kDwBoxBytes = kDwTok * 128;
constexpr int kDwOffE = kDwStages * kDwStageBytes;    // E^T as two bf16 terms: [2 buffers][hi|lo][24][8 words]
constexpr int kDwEWords = 2 * kL * 8;
constexpr int kDwOffBar = kDwOffE + 2 * kDwEWords * 4;
constexpr int kDwSmemBytes = kDwOffBar + kDwStages * 8;

__global__ void __launch_bounds__(kDwThreads, 1)
mhc_stream_dw_kernel(const __grid_constant__ CUtensorMap tmap_xt, const float* __restrict__ e_in, float* __restrict__ dw_part,
                     int64_t T, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023u) __trap();
    uint32_t* et = reinterpret_cast<uint32_t*>(smem + kDwOffE);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kDwOffBar);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_local = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto load_tile = [&](int it) {
        const int s = it % kDwStages;
        const int tok0 = ((int)blockIdx.x + it * (int)gridDim.x) * kDwTok;
        mbar_arrive_expect_tx(&bar_full[s], kDwStageBytes);
        for (int b = 0; b < kRow / 64; ++b)
            tma_load_2d(smem + s * kDwStageBytes + b * kDwBoxBytes, &tmap_xt, &bar_full[s], b * 64, tok0);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kDwStages; ++s) mbar_init(&bar_full[s], 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmap_xt);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (n_local > 0) load_tile(0);
        if (n_local > 1) load_tile(1);
    }
    // warp w owns K rows [128w, 128w+128) -> 8 m-tiles of 16; accumulators [8][3][4]
    const int w = warp, g = lane >> 2, t = lane & 3;
    float acc[8][3][4];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
    const int lm = lane >> 3, lr = lane & 7;
    const uint32_t stage0 = smem_u32(smem);
    const int tid = threadIdx.x;
    for (int it = 0; it < n_local; ++it) {
        const int s = it % kDwStages, eb = it & 1;
        const int64_t tok0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kDwTok;
        // E tile -> two bf16 terms, transposed to [logit][token pair]
        if (tid < kDwTok * kL / 2) {
            const int logit = tid % kL, pair = tid / kL;               // tokens 2*pair, 2*pair+1
            const int64_t ta = tok0 + 2 * pair, tb = ta + 1;
            const float va = ta < T ? __ldg(e_in + ta * kL + logit) : 0.f;
            const float vb = tb < T ? __ldg(e_in + tb * kL + logit) : 0.f;
            const __nv_bfloat16 ha = __float2bfloat16_rn(va), hb = __float2bfloat16_rn(vb);
            const float la = va - __bfloat162float(ha), lb = vb - __bfloat162float(hb);
            uint32_t* dst = et + eb * kDwEWords;
            dst[logit * 8 + pair] = pack_bf16(__bfloat162float(ha), __bfloat162float(hb));
            dst[kL * 8 + logit * 8 + pair] = pack_bf16(la, lb);
        }
        __syncthreads();            // E^T visible; every warp is done with tile it-1, so its stage is free
        if (tid == 0 && it + 2 < n_local) load_tile(it + 2);
        mbar_wait(&bar_full[s], (it / kDwStages) & 1);
        const uint32_t sbase = stage0 + s * kDwStageBytes;
        const uint32_t* ehi = et + eb * kDwEWords;
        uint32_t bh[3][2], bl[3][2];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            bh[nt][0] = ehi[(nt * 8 + g) * 8 + t];            bh[nt][1] = ehi[(nt * 8 + g) * 8 + t + 4];
            bl[nt][0] = ehi[kL * 8 + (nt * 8 + g) * 8 + t];   bl[nt][1] = ehi[kL * 8 + (nt * 8 + g) * 8 + t + 4];
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            // A = x^T: K rows 128w + 16m .. +15 (two 8-column chunks), k = 16 tokens; transposed 8x8 loads
            const int box = 2 * w + (m >> 2), chunk = 2 * (m & 3) + (lm & 1);
            const int row = (lm >> 1) * 8 + lr;               // token
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4_trans(sbase + box * kDwBoxBytes + row * 128 + ((chunk ^ (row & 7)) << 4), a0, a1, a2, a3);
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                mma_bf16_16816(acc[m][nt], a0, a1, a2, a3, bh[nt][0], bh[nt][1]);
                mma_bf16_16816(acc[m][nt], a0, a1, a2, a3, bl[nt][0], bl[nt][1]);
            }
        }
    }
    float* out = dw_part + (size_t)blockIdx.x * kRow * kL;
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
            const int k0 = 128 * w + 16 * m + g;
            *reinterpret_cast<float2*>(out + (size_t)k0 * kL + nt * 8 + 2 * t) = make_float2(acc[m][nt][0], acc[m][nt][1]);
            *reinterpret_cast<float2*>(out + (size_t)(k0 + 8) * kL + nt * 8 + 2 * t) = make_float2(acc[m][nt][2], acc[m][nt][3]);
        }
}

// dphi = scale * dW, dscale = sum_k phi * dW (straight-through the bf16 rounding of scale*phi),
// dphi / dscale from the per-CTA dW partials, dbias / dalpha from the per-CTA accumulators.  Fixed summation order
// (four thread groups take a quarter of the CTAs each, eight interleaved partial sums per thread, combined pairwise),
// so the result is bitwise reproducible.  One CTA = 8 rows x 24 columns x 4 groups; every load of a warp is 128
// contiguous bytes and 32 loads per element are in flight (the kernel is latency-bound: 29 MB behind 148-deep sums).
constexpr int kFinRows = 8, kFinGroups = 4;
__global__ void __launch_bounds__(kFinRows * kL * kFinGroups)
mhc_stream_bwd_finalize_kernel(const float* __restrict__ dw_part, int dw_ctas, const float* __restrict__ cta_accum,
                               int acc_ctas, const float* __restrict__ phi, const float* __restrict__ scale,
                               float* __restrict__ dphi, float* __restrict__ dscale, float* __restrict__ dbias,
                               float* __restrict__ dalpha) {
    __shared__ float part[kFinGroups][kFinRows * kL];
    __shared__ float prod[kFinRows * kL];
    const int e = threadIdx.x % (kFinRows * kL), grp = threadIdx.x / (kFinRows * kL);
    const size_t idx = (size_t)blockIdx.x * (kFinRows * kL) + e;               // element of the [2048, 24] matrix
    const int per = (dw_ctas + kFinGroups - 1) / kFinGroups;
    const int c0 = grp * per, c1 = min(dw_ctas, c0 + per);
    const float* src = dw_part + idx;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += src[(size_t)(c + u) * kRow * kL];
    }
    for (int u = 0; c < c1; ++c, ++u) acc[u] += src[(size_t)c * kRow * kL];
    part[grp][e] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    __syncthreads();
    if (grp == 0) {
        const float dwv = (part[0][e] + part[1][e]) + (part[2][e] + part[3][e]);
        const int row = (int)(idx / kL);
        dphi[idx] = dwv * scale[row];
        prod[e] = dwv * phi[idx];
    }
    __syncthreads();
    if (threadIdx.x < kFinRows) {
        float ds = 0.f;
#pragma unroll
        for (int l = 0; l < kL; ++l) ds += prod[threadIdx.x * kL + l];
        dscale[blockIdx.x * kFinRows + threadIdx.x] = ds;
    }
    if (blockIdx.x == 0 && threadIdx.x < kAccum) {
        float s0 = 0.f;
        for (int k = 0; k < acc_ctas; ++k) s0 += cta_accum[(size_t)k * kAccum + threadIdx.x];
        if (threadIdx.x < kL) dbias[threadIdx.x] = s0; else dalpha[threadIdx.x - kL] = s0;
    }
}

inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct BwdWs {
    float* e; float* dw_part; float* cta_accum;
    size_t total;
};
BwdWs carve(void* base, int64_t T, int ctas) {
    BwdWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return reinterpret_cast<uint8_t*>(base) + o; };
    w.e = reinterpret_cast<float*>(take((size_t)T * kL * 4));
    w.dw_part = reinterpret_cast<float*>(take((size_t)ctas * kRow * kL * 4));
    w.cta_accum = reinterpret_cast<float*>(take((size_t)ctas * kCoefWarps * kAccum * 4));
    w.total = off;
    return w;
}

}  // namespace

// shared with the fused (saved-statistics) backward in mhc_stream_bwd_fused.cu
int launch_bwd_finalize(const float* dw_part, int dw_ctas, const float* cta_accum, int acc_ctas, const float* phi,
                        const float* scale, float* dphi, float* dscale, float* dbias, float* dalpha, cudaStream_t stream) {
    mhc_stream_bwd_finalize_kernel<<<kRow / kFinRows, kFinRows * kL * kFinGroups, 0, stream>>>(dw_part, dw_ctas, cta_accum, acc_ctas, phi, scale, dphi,
                                                                 dscale, dbias, dalpha);
    count_launch();
    return launch_status();
}

}  // namespace hvs

namespace hvs {
bool generic_stream_shape_ok(int n, int C);
size_t generic_stream_bwd_workspace(int64_t T, int n, int C);
int launch_generic_stream_bwd(const void* x, const void* dy, const float* phi, const float* bias, const float* alpha, const float* scale,
                              void* dx, float* dphi, float* dbias, float* dalpha, float* dscale, int64_t T, int n, int C, int iters,
                              float eps_rms, float eps_sk, void* workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace hvs

extern "C" size_t hvs_mhc_stream_bwd_workspace(int64_t T, int n, int C) {
    using namespace hvs;
    if (T >= 0 && (n != kN || C != kC) && generic_stream_shape_ok(n, C)) return generic_stream_bwd_workspace(T, n, C);
    if (T < 0 || n != kN || C != kC) return 0;
    return carve(nullptr, T, sm_count()).total;
}

extern "C" int hvs_mhc_stream_bwd(const void* x, const void* dy, const float* phi, const float* bias,
                                  const float* alpha, const float* scale, void* dx, float* dphi, float* dbias,
                                  float* dalpha, float* dscale, int64_t T, int n, int C, int sk_iters,
                                  float eps_rms, float eps_sk, uint32_t flags, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (T < 0) return HVS_ERR_BAD_ARG;
    if (flags & HVS_MHC_SPLIT_PHI) return HVS_ERR_UNSUPPORTED;
    if (!phi || !bias || !alpha || !scale || !dphi || !dbias || !dalpha || !dscale) return HVS_ERR_BAD_ARG;
    if (T > 0 && (!x || !dy || !dx)) return HVS_ERR_BAD_ARG;
    if ((n != kN || C != kC) && generic_stream_shape_ok(n, C)) {
        // every other stream shape (n in {2, 4}, C % 8 == 0, C <= 1024): the general three-launch backward
        if (T >= (int64_t)1 << 31) return HVS_ERR_UNSUPPORTED;
        if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) return HVS_ERR_ALIGNMENT;
        return launch_generic_stream_bwd(x, dy, phi, bias, alpha, scale, dx, dphi, dbias, dalpha, dscale, T, n, C, sk_iters, eps_rms, eps_sk,
                                         workspace, workspace_bytes, stream);
    }
    if (n != kN || C != kC || sk_iters < 0 || sk_iters > kMaxIters) return HVS_ERR_UNSUPPORTED;
    if (T * kN >= (int64_t)1 << 31) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15)
        return HVS_ERR_ALIGNMENT;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return HVS_ERR_ALIGNMENT;
    const int sms = sm_count();
    const BwdWs ws = carve(workspace, T, sms);
    if (workspace_bytes < ws.total) return HVS_ERR_WORKSPACE;
    int grid1 = 0, grid2 = 0;
    if (T > 0) {
        CUtensorMap tx, tdy, tdx, txt;
        int rc = make_tmap_bf16_2d(&tx, x, (uint64_t)T * kN, kC, kTok * kN);
        if (rc) return rc;
        rc = make_tmap_bf16_2d(&tdy, dy, (uint64_t)T * kN, kC, kTok * kN);
        if (rc) return rc;
        rc = make_tmap_bf16_2d(&tdx, dx, (uint64_t)T * kN, kC, kTok * kN);
        if (rc) return rc;
        rc = make_tmap_bf16_2d(&txt, x, (uint64_t)T, kRow, kDwTok);
        if (rc) return rc;
        HVS_SET_MAX_SMEM(mhc_stream_bwd_kernel, kSmemBytes);
        HVS_SET_MAX_SMEM(mhc_stream_dw_kernel, kDwSmemBytes);
        BwdParams p;
        p.phi = phi; p.bias = bias; p.alpha = alpha; p.scale = scale;
        p.e_out = ws.e; p.cta_accum = ws.cta_accum;
        p.T = T;
        p.num_tiles = (int)((T + kTok - 1) / kTok);
        p.sk_iters = sk_iters; p.eps_rms = eps_rms; p.eps_sk = eps_sk;
        grid1 = p.num_tiles < sms ? p.num_tiles : sms;
        timer_begin(1, stream);
        mhc_stream_bwd_kernel<<<grid1, kThreads, kSmemBytes, stream>>>(tx, tdy, tdx, p);
        timer_end(1, stream);
        count_launch();
        int rc2 = launch_status();
        if (rc2) return rc2;
        const int tiles2 = (int)((T + kDwTok - 1) / kDwTok);
        grid2 = tiles2 < sms ? tiles2 : sms;
        timer_begin(2, stream);
        mhc_stream_dw_kernel<<<grid2, kDwThreads, kDwSmemBytes, stream>>>(txt, ws.e, ws.dw_part, T, tiles2);
        timer_end(2, stream);
        count_launch();
        rc2 = launch_status();
        if (rc2) return rc2;
    }
    timer_begin(3, stream);
    mhc_stream_bwd_finalize_kernel<<<kRow / kFinRows, kFinRows * kL * kFinGroups, 0, stream>>>(ws.dw_part, grid2, ws.cta_accum, kCoefWarps * grid1, phi, scale, dphi,
                                                                 dscale, dbias, dalpha);
    timer_end(3, stream);
    count_launch();
    return launch_status();
}
