// K1 backward for training, ONE fused kernel: dx and the parameter gradients in a single pass over x and dy
// (every byte crosses HBM once).  The forward saved the un-normalised projection and the sum of squares per
// token (hvs_mhc_stream_fwd_save, 112 B/token).  The two dense contractions go to the tcgen05 tensor cores with
// accumulators in tensor memory, operands straight from the TMA-landed token tile.  A small tcgen05.mma occupies
// the tensor pipe ~55 cycles whatever its shape (tools/micro/umma.cu), so they are few and big, 32 per tile:
//   G = dy x^T (per-token 4x4)   16 x tcgen05.mma D[64x64] += dy rows * (x rows)^T: the 512 channels as two halves side
//                                 by side (rows = (stream, half, token), K = the 256 channels of a half); the read-out
//                                 keeps the (half, token) diagonal and adds the halves
//   dW = x^T E (2048 x 24)       16 x tcgen05.mma D[128x32] += x-atoms^T (MN-major A, read in place) * [E_hi ; E_lo]:
//                                 the two bf16 terms of E ride the two halves of K = 16 against the SAME eight token
//                                 rows of x (stride-0 K step); accumulated over the whole kernel in 384 tensor-memory
//                                 columns (16 blocks of 128 channels, 24 columns apart: the 8 spare columns of N = 32
//                                 add zeros to the neighbour)
// Roles (640 threads, one CTA per SM, 3 stages of 8 tokens: x | dy = 64 KB each, one 4-D TMA box per tensor):
//   front warp          (convergent; lane 0 issues) TMA loads, the dW MMAs, TMA stores of dx (one output stream at
//                       a time, during the dx pass), stage recycling
//   3 coefficient warps (one per stage, four lanes per token) forward Sinkhorn AHEAD of the tile from the saved
//                       record (fetched by the warp itself), tracking the cumulative scalings; the G MMAs when the tile
//                       has landed; once G is there: gate gradients, reverse sweep in the scaling form (no reciprocals,
//                       one shuffle step per iteration), softmax / RMSNorm backward
//   16 worker warps     G out of tensor memory the moment it completes; e = d raw as bf16 hi/lo into the E operand
//                       tile, dbias; dx = M^T dy + kappa x + W e   (W e on the warp MMA path, scale*phi in A-fragment
//                       order: 32 registers + 16 parked in tensor memory; mixing in packed fp32x2 FMAs), written IN
//                       PLACE over dy, one rounding
// Oracle: autograd through oracle/mhc_ref.py::stream_mhc_forward (reference primitives
// src/models/manifold_layers.py:56-77, :213-216, :449-456).
#include "common.cuh"
#include "mhc_stream_shared.cuh"
#include "ptx_sm100.cuh"
#include "umma_sm100.cuh"

namespace hvs {

int launch_bwd_finalize(const float* dw_part, int dw_ctas, const float* cta_accum, int acc_ctas, const float* phi,
                        const float* scale, float* dphi, float* dscale, float* dbias, float* dalpha, cudaStream_t stream);

namespace {

constexpr int kTok = 8;
constexpr int kStages = 3;
constexpr int kWorkers = 16;
constexpr int kWorkerThreads = kWorkers * 32;
constexpr int kThreads = (kWorkers + 4) * 32;         // + coefficient warp 0, front warp, coefficient warps 1, 2
constexpr int kStageBytes = 65536;                    // [x 32 KB][dy 32 KB]; each = 32 swizzle atoms (stream j, 64-channel block cb) of 8 tokens x 128 B
constexpr int kHalf = 32768;
constexpr int kSaved = HVS_MHC_SAVED_STRIDE;          // floats per token in the saved record: raw[24], sum x^2, pad
constexpr int kMaxIters = 24;                        // normalisers of every iteration are kept in shared memory
constexpr int kCoefWarps = 3;
constexpr int kAccum = kL + 3;

constexpr int kSavedBytes = kTok * kSaved * 4;        // 896
constexpr int kOffSaved = kStages * kStageBytes;
constexpr int kOffWrec = kOffSaved + kStages * kSavedBytes;       // per stage: e bf16 pairs [8][12] | M pairs [4][4][4][2] | kappa [8]
                                                                  //            | alpha_g * inv_rms [8][3] | d logits [8][24]
constexpr int kWrecBytes = 1152, kWrecM = 384, kWrecK = 960, kWrecS = 992, kMpStride = 36;   // M pairs: 4 token pairs x (32 + 4 pad) floats   // G [8][16] fp32 overlays bytes [0, 512) until M is written
constexpr int kOffEt = kOffWrec + kStages * kWrecBytes;           // per stage: E tile, no swizzle: two K chunks (bf16 hi | lo of e over the 8 tokens) of
                                                                  // 5 row groups (8 rows x 16 B): zero | logits 0-7 | 8-15 | 16-23 | zero
constexpr int kEtBytes = 1280, kEtChunk = 640, kEtRow0 = 128;
constexpr int kSkWords = kTok * (kMaxIters + 1) * 8;     // slot 0 = the scalings before the first iteration (ones)
constexpr int kOffSk = kOffEt + kStages * kEtBytes;                 // per coefficient warp: scalings u | v of every Sinkhorn iteration
constexpr int kOffDl = kOffSk + kCoefWarps * kSkWords * 4;        // per stage: d logits [8][24] fp32 (coefficient warp -> workers)
constexpr int kDlBytes = 768;
constexpr int kOffPart = kOffDl + kStages * kDlBytes;             // end-of-kernel dbias fold: [8 tokens][24] fp32
constexpr int kOffBias = kOffPart + kTok * kL * 4;                // bias[24] staged once (float4 broadcast reads)
constexpr int kOffBar = kOffBias + 128;
constexpr int kOffTmem = kOffBar + (7 * kStages + 4 * kStages + 1) * 8;
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kOffBar % 8 == 0 && kOffSaved % 16 == 0 && kOffWrec % 16 == 0 && kOffEt % 128 == 0 && kOffSk % 16 == 0 && kOffDl % 16 == 0, "alignment");
static_assert(kSmemBytes <= 232448, "shared memory budget");

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColDw = 0;                        // dW: 16 blocks of 128 channels (lane = channel in the block), block p in columns [24p, 24p+24)
constexpr uint32_t kColGs = 384;                      // dy x^T of a tile in two channel halves: 64 columns
constexpr uint32_t kColW = 448;                       // 64 columns = 16 registers per worker thread: the K = 8 step of its W
                                                      //     fragments lives here instead of spilling (no L1 to speak of)

constexpr int kBarRec = 1 /*,2,3*/, kBarW = 4;

struct FusedParams {
    const float* phi;
    const float* bias;
    const float* alpha;
    const float* scale;
    const float* saved;    // [T, 28]
    float* dw_part;        // [grid, 2048, 24]
    float* cta_accum;      // [grid, 3, 27]
    long long* dbg;        // optional [grid, 4, 8] cycle counters: coefficient warps 0..2, front thread (development aid)
    int64_t T;
    int num_tiles;
    int sk_iters;
    float eps_rms, eps_sk;
};

// development aid (compile with -DHVS_FUSED_TRACE): cycle counters of the coefficient / front warps and per-tile
// event timestamps of CTA 0 (first 64 tiles), written to p.dbg.  Compiled out by default.
#ifdef HVS_FUSED_TRACE
#define HVS_TR(tile, ev) do { if (p.dbg && blockIdx.x == 0 && (tile) < 64) p.dbg[148 * 4 * 8 + (tile) * 12 + (ev)] = clock64(); } while (0)
#define HVS_TRACE_ON 1
#else
#define HVS_TR(tile, ev) do { } while (0)
#define HVS_TRACE_ON 0
#endif

typedef unsigned long long u64;
__device__ __forceinline__ uint32_t sel32(uint32_t a, uint32_t b, uint32_t c) {       // c != 0 ? a : b, kept out of the optimiser's sight
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.b32 %0, %1, %2, p;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// D(16x8,f32) += A(16x8,bf16,row) * B(8x8,bf16,col)
__device__ __forceinline__ void mma_bf16_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}

// kAdaptive (HVS_MHC_ADAPTIVE_ITERS): the forward replay stops at the first iteration that changes no entry of P of any token
// of the warp by more than 2^-20 relative (see mhc_stream_fwd.cu: what the remaining iterations would change is ~2e-6), and
// the reverse sweep differentiates the iterations actually run.
__device__ __forceinline__ bool rel_close2(u64 a, u64 b) {      // |a - b| <= 2^-20 a for both halves (a > 0)
    float a0, a1, b0, b1;
    upk2(a, a0, a1); upk2(b, b0, b1);
    return fabsf(a0 - b0) <= 9.5367431640625e-07f * a0 && fabsf(a1 - b1) <= 9.5367431640625e-07f * a1;
}
template <bool kAdaptive>
__global__ void __launch_bounds__(kThreads, 1)
mhc_stream_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                            const __grid_constant__ CUtensorMap tmap_dx, const FusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023u) __trap();
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kOffBar);     // tile + saved records landed
    uint64_t* bar_ed = bar_full + kStages;                                // E tile written (worker warps)
    uint64_t* bar_dxr = bar_ed + kStages;                                 // [stage][stream] dx of one output stream staged by the workers
    uint64_t* bar_dw = bar_dxr + kN * kStages;                            // dW MMAs of the tile complete
    uint64_t* bar_sv = bar_dw + kStages;                                  // saved records of the stage's next tile landed
    uint64_t* bar_cd = bar_sv + kStages;                                  // coefficients of the tile written (coefficient warp)
    uint64_t* bar_gs = bar_cd + kStages;                                  // G MMAs of the tile complete
    uint64_t* bar_gr = bar_gs + kStages;                                  // G buffer read out by the 16 worker warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);

    // (the shuffle tells the compiler that the warp index is warp-uniform: everything derived from it -- stage addresses,
    // tensor-core descriptors -- then lives in uniform registers instead of going through an elect / R2UR loop per MMA)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int wtid = threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int n_local = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_ed[s], kTok * kL);         // every thread that builds a piece of the E tile arrives itself
            for (int j = 0; j < kN; ++j) mbar_init(&bar_dxr[s * kN + j], kWorkerThreads);
            mbar_init(&bar_dw[s], 1);
            mbar_init(&bar_sv[s], 1);
            mbar_init(&bar_cd[s], 32);
            mbar_init(&bar_gs[s], 1);
            if (s < 1) mbar_init(&bar_gr[s], kWorkers);
        }
        fence_mbar_init();
    }
    if (threadIdx.x < kL) reinterpret_cast<float*>(smem + kOffBias)[threadIdx.x] = __ldg(p.bias + threadIdx.x);
    if (warp == kWorkers + 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the dW accumulators start at zero (every MMA below accumulates); so do the 8 padding rows of the E tiles
    for (int i = threadIdx.x; i < kStages * kEtBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem + kOffEt)[i] = 0u;
    if (warp < kWorkers) {
        const uint32_t zero[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t tq = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + kColDw + (uint32_t)((warp >> 2) * 96);
#pragma unroll
        for (int c = 0; c < 6; ++c) tmem_st16(tq + 16 * c, zero);
        tmem_wait_st();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= kWorkers) {
      if (warp == kWorkers + 1) {
        // ===================================================== front warp: loads, dW MMA issue, stores.  The whole warp
        // runs the loop (every address / descriptor stays warp-uniform, i.e. in uniform registers); lane 0 issues the
        // asynchronous operations.  (The G MMAs of a tile are issued by that tile's coefficient warp, which is the
        // one waiting for them: a tcgen05.mma costs its issuer ~55 cycles, 64 of them per tile are too many for one warp.)
        {
            const bool leader = lane == 0;
            if (leader) {
                tma_prefetch_desc(&tmap_x);
                tma_prefetch_desc(&tmap_dy);
                tma_prefetch_desc(&tmap_dx);
            }
            const uint32_t s0 = smem_u32(smem);
            // a tile = two 32 KB boxes on one mbarrier; the x half of a stage is free as soon as the dx pass and the dW MMAs
            // are done with it, the dy half only once the dx store has read it
            auto load_x = [&](int it) {
                const int s = it % kStages;
                const int tok0 = ((int)blockIdx.x + it * (int)gridDim.x) * kTok;
                if (leader) {
                    mbar_arrive_expect_tx(&bar_full[s], kStageBytes);
                    tma_load_4d(smem + s * kStageBytes, &tmap_x, &bar_full[s], 0, tok0, 0, 0);
                    HVS_TR(it, 0);
                }
            };
            auto load_dy = [&](int it) {
                const int s = it % kStages;
                const int tok0 = ((int)blockIdx.x + it * (int)gridDim.x) * kTok;
                if (leader) tma_load_4d(smem + s * kStageBytes + kHalf, &tmap_dy, &bar_full[s], 0, tok0, 0, 0);
            };
            long long facc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            long long fprev = clock64();
#define HVS_FTICK(slot) do { if (HVS_TRACE_ON && p.dbg) { const long long tn = clock64(); facc[slot] += tn - fprev; fprev = tn; } } while (0)
            const uint32_t id_dw = umma_idesc_bf16(128, 32, 1, 0);
            auto retire = [&](int k) {
                const int s = k % kStages;
                const uint32_t ph = (uint32_t)(k / kStages) & 1u;
                const int tok0 = ((int)blockIdx.x + k * (int)gridDim.x) * kTok;
                // dW += x^T E for the tile whose coefficients are ready: 16 blocks of 128 channels,
                // K = 16 = the 8 token rows twice (stride 0) against [E_hi ; E_lo]
                mbar_wait(&bar_ed[s], ph);
                HVS_FTICK(2);
                tc_fence_after();
                // M = 128 channels (two atoms, 1 KB apart), N = 32 = 24 logits + 8 zero rows: the accumulators sit 24 columns
                // apart, so the 8 extra columns add 0 to the next block's first columns (same issuer, in order).  The last
                // block puts the zero rows first instead (E tile one row group earlier, accumulator 8 columns lower).
                const uint32_t et = s0 + kOffEt + s * kEtBytes;
                const uint64_t bdesc = umma_smem_desc(et + kEtRow0, kEtChunk, 128, kUmmaLayoutNone);
                const uint64_t bdesc_last = umma_smem_desc(et, kEtChunk, 128, kUmmaLayoutNone);
                const uint64_t adesc0 = umma_smem_desc(s0 + s * kStageBytes, 1024, 0, kUmmaLayoutSw128);
#pragma unroll
                for (int b = 0; b < 16; ++b)           // block b = atoms 2b, 2b+1 at +2 KB each (= +128 in the address field)
                    if (leader)
                        umma_bf16_ss(tmem_base + kColDw + (uint32_t)(b * 24 - (b == 15 ? 8 : 0)), adesc0 + (uint64_t)(b * 128),
                                     b == 15 ? bdesc_last : bdesc, id_dw, 1u);
                if (leader) umma_commit(&bar_dw[s]);
                HVS_FTICK(3);
                // dx of the tile (written in place over dy) -> HBM, one output stream at a time while the dx pass is still
                // working on the next: when the pass ends only the last 8 KB are left to drain before dy can be re-loaded
#pragma unroll
                for (int j = 0; j < kN; ++j) {
                    mbar_wait(&bar_dxr[s * kN + j], ph);
                    if (leader) {
                        if (j == kN - 1) HVS_TR(k, 10);
                        tma_store_4d(&tmap_dx, smem + s * kStageBytes + kHalf + j * 8192, 0, tok0, 0, j);
                        bulk_commit();
                    }
                }
                HVS_FTICK(4);
                mbar_wait(&bar_dw[s], ph);                       // the tensor core is done reading x of this stage
                if (k + kStages < n_local) load_x(k + kStages);
                HVS_FTICK(6);
                if (leader) bulk_wait_read<0>();
                __syncwarp();
                if (leader) HVS_TR(k, 11);
                HVS_FTICK(5);
                if (k + kStages < n_local) load_dy(k + kStages);
                HVS_FTICK(7);
            };
            for (int it = 0; it < kStages && it < n_local; ++it) { load_x(it); load_dy(it); }
            for (int k = 0; k < n_local; ++k) retire(k);
            if (leader) bulk_wait<0>();
            if (HVS_TRACE_ON && p.dbg && leader) for (int q = 0; q < 8; ++q) p.dbg[((size_t)blockIdx.x * 4 + 3) * 8 + q] = facc[q];
        }
      } else {
        // ===================================================== coefficient warps (warps 16, 18, 19 <-> stage 0, 1, 2)
        // Four lanes per token: lane 4 tk + i owns row i of the token's 4x4 blocks as two packed fp32x2 registers, so
        // row sums / row dot products are local and column sums are two xor-shuffle steps inside the quad.  These warps
        // share their scheduler with four worker warps each, so what counts is the number of instructions per
        // iteration (about half of a lane-per-token layout, whose other 24 lanes only repeat work).
        //   forward  (reference arithmetic, P / (sum + eps)) runs AHEAD of the tile: it needs only the saved record,
        //            which this warp fetches itself; it also tracks the cumulative scalings u, v with
        //            P_k = diag(u_k) K diag(v_k), K = the softmax start
        //   backward differentiates that scaling form exactly: u_k = 1 / (K v_{k-1}), v_k = 1 / (K^T u_k)  (the eps of
        //            the reference, 1e-8 against sums of 1, is below fp32 resolution), so the sweep needs neither
        //            reciprocals nor a reconstruction of P: per iteration four 4x4 mat-vecs / rank-1 updates.
        const int cw = warp == kWorkers ? 0 : warp - (kWorkers + 1);     // warps 16, 18, 19 -> 0, 1, 2
        const int s = cw;
        const int tk = lane >> 2, i4 = lane & 3, qb = lane & ~3;
        const float a_pre = __ldg(p.alpha + 0), a_post = __ldg(p.alpha + 1), a_res = __ldg(p.alpha + 2);
        const float b_pre4 = __ldg(p.bias + i4), b_post4 = __ldg(p.bias + kN + i4);
        const float4 b_res4 = __ldg(reinterpret_cast<const float4*>(p.bias + 2 * kN) + i4);
        const float eps = p.eps_sk;
        const u64 eps2 = pk2(eps, eps);
        const uint32_t id_gs = umma_idesc_bf16(64, 64, 0, 0);
        float acc_a[3] = {0.f, 0.f, 0.f};                  // dalpha terms (row-0 lanes)
        // scalings of every iteration: [iter][ u: 32 floats, lane-indexed | v: 8 tokens x 4 ]
        float* sku = reinterpret_cast<float*>(smem + kOffSk) + cw * kSkWords + lane;
        float* skv = reinterpret_cast<float*>(smem + kOffSk) + cw * kSkWords + 32 + tk * 4;
        sku[0] = 1.f;                                      // slot 0: u_0 = v_0 = 1 (iteration k writes slot k + 1)
        if (i4 == 0) *reinterpret_cast<float4*>(skv) = make_float4(1.f, 1.f, 1.f, 1.f);
        __syncwarp();
        const float* rsv = reinterpret_cast<const float*>(smem + kOffSaved + s * kSavedBytes);
        uint8_t* wrec = smem + kOffWrec + s * kWrecBytes;
        auto fetch_saved = [&](int it) {                   // lane 0: the 8 saved records of tile `it` -> shared memory
            const int64_t tok0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kTok;
            const int64_t left = p.T - tok0;
            const uint32_t bytes = (left >= kTok ? kTok : (uint32_t)left) * kSaved * 4;
            mbar_arrive_expect_tx(&bar_sv[s], bytes);
            bulk_load_1d(smem + kOffSaved + s * kSavedBytes, p.saved + tok0 * kSaved, bytes, &bar_sv[s]);
        };
        auto quad_sum2 = [](u64 v) {                       // sum over the four lanes of a token, identical in all four
            v = add2(v, __shfl_xor_sync(0xffffffffu, v, 1));
            return add2(v, __shfl_xor_sync(0xffffffffu, v, 2));
        };
        if (lane == 0 && cw < n_local) fetch_saved(cw);
        long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tprev = clock64();
#define HVS_TICK(slot) do { if (HVS_TRACE_ON && p.dbg) { const long long tn = clock64(); tacc[slot] += tn - tprev; tprev = tn; } } while (0)
        for (int it = cw; it < n_local; it += kCoefWarps) {
            const uint32_t ph = (uint32_t)(it / kStages) & 1u;
            mbar_wait(&bar_sv[s], ph);
            if (lane == 0) HVS_TR(it, 3);
            HVS_TICK(0);
            const int64_t tok0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kTok;
            // ---- prologue: inverse RMS, gates, softmax start of row i4.
            // Rows past T: x = dy = 0 (TMA fill) and the records were not loaded; they run on zeros.
            const bool valid = tok0 + tk < p.T;
            const float* r = rsv + tk * kSaved;
            const float inv_rms = __fdiv_rn(1.0f, __fsqrt_rn(fmaf(valid ? r[kL] : 1.0f, 1.0f / kRow, p.eps_rms)));
            const float raw_pre = valid ? r[i4] : 0.f, raw_post = valid ? r[kN + i4] : 0.f;
            const float4 rr = valid ? *reinterpret_cast<const float4*>(r + 2 * kN + 4 * i4) : make_float4(0.f, 0.f, 0.f, 0.f);
            // the record now lives in registers: fetch the next tile's into the same buffer right away, a whole tile ahead of
            // its use (fetched at the end of the tile it cost this warp 1.5 k cycles of HBM latency per tile)
            __syncwarp();
            if (lane == 0 && it + kCoefWarps < n_local) {
                fence_proxy_async_smem();
                fetch_saved(it + kCoefWarps);
            }
            u64 K01, K23;                                  // row i4 of K = 4 softmax(logits): the Sinkhorn start
            {
                const float l0 = fmaf(a_res, rr.x * inv_rms, b_res4.x), l1 = fmaf(a_res, rr.y * inv_rms, b_res4.y);
                const float l2 = fmaf(a_res, rr.z * inv_rms, b_res4.z), l3 = fmaf(a_res, rr.w * inv_rms, b_res4.w);
                const float mx = fmaxf(fmaxf(l0, l1), fmaxf(l2, l3));
                const float e0 = fast_exp(l0 - mx), e1 = fast_exp(l1 - mx), e2 = fast_exp(l2 - mx), e3 = fast_exp(l3 - mx);
                const float r4 = 4.0f * rcp_approx((e0 + e1) + (e2 + e3));
                K01 = pk2(e0 * r4, e1 * r4);
                K23 = pk2(e2 * r4, e3 * r4);
            }
            const float hpre = sigmoid_f32(fmaf(a_pre, raw_pre * inv_rms, b_pre4));
            const float hpost = 2.0f * sigmoid_f32(fmaf(a_post, raw_post * inv_rms, b_post4));
            HVS_TICK(1);
            // ---- forward Sinkhorn; the cumulative scalings after every iteration are kept for the reverse sweep
            u64 P01 = K01, P23 = K23;
            int iters_run = p.sk_iters;                   // iterations whose scalings are in the history (kAdaptive: maybe fewer)
            {
                float ui = 1.f;
                u64 V01 = pk2(1.f, 1.f), V23 = V01;
                for (int k = 0; k < p.sk_iters; ++k) {
                    const u64 Q01 = P01, Q23 = P23;
                    float sa, sb;
                    upk2(add2(P01, P23), sa, sb);
                    const float rw = rcp_approx((sa + sb) + eps);
                    const u64 rw2 = pk2(rw, rw);
                    P01 = mul2(P01, rw2);
                    P23 = mul2(P23, rw2);
                    ui *= rw;
                    const u64 c01 = add2(quad_sum2(P01), eps2), c23 = add2(quad_sum2(P23), eps2);
                    float c0, c1, c2, c3;
                    upk2(c01, c0, c1); upk2(c23, c2, c3);
                    const u64 rc01 = pk2(rcp_approx(c0), rcp_approx(c1)), rc23 = pk2(rcp_approx(c2), rcp_approx(c3));
                    P01 = mul2(P01, rc01);
                    P23 = mul2(P23, rc23);
                    V01 = mul2(V01, rc01);
                    V23 = mul2(V23, rc23);
                    sku[(k + 1) * 64] = ui;
                    if (i4 == 0) {
                        float v0, v1, v2, v3;
                        upk2(V01, v0, v1); upk2(V23, v2, v3);
                        *reinterpret_cast<float4*>(skv + (k + 1) * 64) = make_float4(v0, v1, v2, v3);
                    }
                    if (kAdaptive && __all_sync(0xffffffffu, rel_close2(P01, Q01) && rel_close2(P23, Q23))) { iters_run = k + 1; break; }
                }
            }
            if (lane == 0) HVS_TR(it, 4);
            HVS_TICK(2);
            // gathered while the tile is still in flight: all four rows of K (this lane's own among them) and the four H_pre
            u64 KR[4][2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                KR[k][0] = __shfl_sync(0xffffffffu, K01, qb + k);
                KR[k][1] = __shfl_sync(0xffffffffu, K23, qb + k);
            }
            const float hp0 = __shfl_sync(0xffffffffu, hpre, qb), hp1 = __shfl_sync(0xffffffffu, hpre, qb + 1);
            const float hp2 = __shfl_sync(0xffffffffu, hpre, qb + 2), hp3 = __shfl_sync(0xffffffffu, hpre, qb + 3);
            // ---- G = dy x^T of the landed tile on the tensor core.  Every tcgen05.mma costs ~55 cycles whatever its shape,
            //      so the 512 channels go in as two halves side by side: A rows = (stream, half, token) of dy, B rows = the
            //      same of x (row groups 4 KB apart), K = the 256 channels of a half: 16 MMAs of 64 x 64; the read-out adds
            //      the two (half, half) diagonal blocks.  This warp issues them (lane 0; descriptors warp-uniform); the 16
            //      worker warps read the token diagonal out of tensor memory into the tile's record.
            mbar_wait(&bar_full[s], ph);
            if (it >= 1) {
                // one G buffer, three issuing warps: strictly in tile order.  (G of tile it - 1 complete first -- only then
                // is the parity wait on the read-out barrier unambiguous.)
                mbar_wait(&bar_gs[(it - 1) % kStages], (uint32_t)((it - 1) / kStages) & 1u);
                mbar_wait(&bar_gr[0], (uint32_t)(it - 1) & 1u);                      // buffer read out (tile it - 1)
            }
            if (lane == 0) HVS_TR(it, 1);
            tc_fence_after();
            {
                const uint32_t st = smem_u32(smem) + s * kStageBytes;
                const uint64_t xdesc0 = umma_smem_desc(st, 16, 4096, kUmmaLayoutSw128);
                const uint64_t ydesc0 = umma_smem_desc(st + kHalf, 16, 4096, kUmmaLayoutSw128);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t o = (uint64_t)(c4 * 64 + ks * 2);                 // +1 KB per block, +32 B per K step
                        if (lane == 0) umma_bf16_ss(tmem_base + kColGs, ydesc0 + o, xdesc0 + o, id_gs, (uint32_t)((c4 | ks) != 0));
                    }
                if (lane == 0) umma_commit(&bar_gs[s]);
                if (lane == 0) HVS_TR(it, 9);
            }
            bar_sync(kBarRec + s, kWorkerThreads + 32);
            HVS_TICK(3);
            // ---- M = P + hpost (x) hpre for the workers, gate gradients from G (row i4 of the token's G)
            const float4 g = *reinterpret_cast<const float4*>(wrec + lane * 16);
            __syncwarp();                                   // every lane has its G row: M may overwrite the record
            float dl_pre, dl_post;
            {
                const float dhpost = fmaf(g.w, hp3, fmaf(g.z, hp2, fmaf(g.y, hp1, g.x * hp0)));
                dl_post = dhpost * hpost * (1.0f - 0.5f * hpost);
                const u64 hq2 = pk2(hpost, hpost);
                const u64 d01 = quad_sum2(mul2(pk2(g.x, g.y), hq2)), d23 = quad_sum2(mul2(pk2(g.z, g.w), hq2));
                float d0, d1, d2, d3;
                upk2(d01, d0, d1); upk2(d23, d2, d3);
                const float dhpre = i4 == 0 ? d0 : i4 == 1 ? d1 : i4 == 2 ? d2 : d3;
                dl_pre = dhpre * hpre * (1.0f - hpre);
                float p0, p1, p2, p3;
                upk2(P01, p0, p1); upk2(P23, p2, p3);
                float* mp = reinterpret_cast<float*>(wrec + kWrecM) + (tk >> 1) * kMpStride + i4 * 2 + (tk & 1);
                mp[0] = fmaf(hpost, hp0, p0);               // M[i4][j] at j * 8
                mp[8] = fmaf(hpost, hp1, p1);
                mp[16] = fmaf(hpost, hp2, p2);
                mp[24] = fmaf(hpost, hp3, p3);
            }
            if (lane == 0) HVS_TR(it, 5);
            HVS_TICK(4);
            // ---- reverse sweep in the scaling form; dK accumulates row i4; ub / vb are the adjoints of the current u / v.
            //      This loop is the longest serial piece of a tile's life and its warp is short of issue slots, so it is
            //      written for instruction count: the history is walked with two pointers (slot 0 holds the ones before the
            //      first iteration: no special case for k = 0), and the signs are folded (tn = vb v^2 = -tb, ubn = -ub).
            u64 dK0, dK1;
            {
                const int last = iters_run - 1;
                const float* pu = sku + (last + 1) * 64;                        // u_k | v_k of the iteration being undone
                const float* pv = skv + (last + 1) * 64;
                float un = *pu;
                float4 vn = *reinterpret_cast<const float4*>(pv);
                // adjoints of the output P = diag(u) K diag(v):  dK = G u v^T,  ub_i = sum_j G_ij K_ij v_j,  vb_j = sum_i G_ij K_ij u_i.
                // ubn is kept as a packed partial-sum pair (its two halves are added when it is consumed).
                u64 ubn, vb01, vb23;
                {
                    const u64 v01 = pk2(vn.x, vn.y), v23 = pk2(vn.z, vn.w), ui = pk2(un, un);
                    const u64 g0 = pk2(g.x, g.y), g1 = pk2(g.z, g.w);
                    const u64 gk0 = mul2(g0, K01), gk1 = mul2(g1, K23);
                    ubn = mul2(fma2(gk1, v23, mul2(gk0, v01)), pk2(-1.f, -1.f));
                    vb01 = quad_sum2(mul2(gk0, ui));
                    vb23 = quad_sum2(mul2(gk1, ui));
                    dK0 = mul2(mul2(g0, v01), ui);
                    dK1 = mul2(mul2(g1, v23), ui);
                }
                for (int k = last; k >= 0; --k) {
                    const float uu = un;
                    const u64 v01 = pk2(vn.x, vn.y), v23 = pk2(vn.z, vn.w);
                    pu -= 64; pv -= 64;
                    un = *pu;                                                   // u_{k-1}, v_{k-1}
                    vn = *reinterpret_cast<const float4*>(pv);
                    const u64 vp01 = pk2(vn.x, vn.y), vp23 = pk2(vn.z, vn.w);
                    // v_k = 1 / (K^T u_k):  tb = -vb v_k^2 = -tn ;  ub += K tb ;  dK += u_k tb^T
                    const u64 tn01 = mul2(mul2(vb01, v01), v01), tn23 = mul2(mul2(vb23, v23), v23);
                    const u64 nui = pk2(-uu, -uu);
                    ubn = fma2(K23, tn23, fma2(K01, tn01, ubn));
                    dK0 = fma2(tn01, nui, dK0);
                    dK1 = fma2(tn23, nui, dK1);
                    // u_k = 1 / (K v_{k-1}):  sb = -ub u_k^2 = ubn u_k^2
                    float ua, ub_;
                    upk2(ubn, ua, ub_);
                    const float sb = (ua + ub_) * (uu * uu);
                    ubn = pk2(0.f, 0.f);
                    // vb = K^T sb ;  dK += sb v_{k-1}^T.  The four sb of the token are gathered in ONE shuffle step (independent
                    // shuffles) and every lane forms the whole column sum from its copy of K -- the two dependent steps of a
                    // butterfly sum were the longest link of the iteration.
                    const u64 s2 = pk2(sb, sb);
                    {
                        const float t0 = __shfl_sync(0xffffffffu, sb, qb), t1 = __shfl_sync(0xffffffffu, sb, qb + 1);
                        const float t2 = __shfl_sync(0xffffffffu, sb, qb + 2), t3 = __shfl_sync(0xffffffffu, sb, qb + 3);
                        const u64 q0 = pk2(t0, t0), q1 = pk2(t1, t1), q2 = pk2(t2, t2), q3 = pk2(t3, t3);
                        vb01 = add2(fma2(KR[1][0], q1, mul2(KR[0][0], q0)), fma2(KR[3][0], q3, mul2(KR[2][0], q2)));
                        vb23 = add2(fma2(KR[1][1], q1, mul2(KR[0][1], q0)), fma2(KR[3][1], q3, mul2(KR[2][1], q2)));
                    }
                    dK0 = fma2(vp01, s2, dK0);
                    dK1 = fma2(vp23, s2, dK1);
                }
            }
            HVS_TICK(5);
            // ---- softmax * 4 backward: dl = K (dK - sum_j(dK K) / 4); the sums for kappa (RMSNorm backward) and dalpha
            {
                float qa, qb_;
                upk2(fma2(dK1, K23, mul2(dK0, K01)), qa, qb_);
                const float nqs = -0.25f * (qa + qb_);
                const u64 nq2 = pk2(nqs, nqs);
                float d0, d1, d2, d3;
                upk2(mul2(K01, add2(dK0, nq2)), d0, d1);
                upk2(mul2(K23, add2(dK1, nq2)), d2, d3);
                const u64 sums = quad_sum2(pk2(dl_pre * raw_pre, dl_post * raw_post));
                const u64 sumr = quad_sum2(pk2(fmaf(d3, rr.w, fmaf(d2, rr.z, fmaf(d1, rr.y, d0 * rr.x))), 0.f));
                float da_pre, da_post, da_res, unused;
                upk2(sums, da_pre, da_post);
                upk2(sumr, da_res, unused);
                // d inv_rms = sum_k e_k raw_k / inv_rms with e = alpha_g * dl * inv_rms ;  kappa = -d inv_rms * inv_rms^3 / N
                if (i4 == 0) {
                    const float dsum = (a_pre * da_pre + a_post * da_post + a_res * da_res) * inv_rms;
                    acc_a[0] = fmaf(da_pre, inv_rms, acc_a[0]);
                    acc_a[1] = fmaf(da_post, inv_rms, acc_a[1]);
                    acc_a[2] = fmaf(da_res, inv_rms, acc_a[2]);
                    reinterpret_cast<float*>(wrec + kWrecK)[tk] = -dsum * inv_rms * inv_rms * (1.0f / kRow);
                    float* sc = reinterpret_cast<float*>(wrec + kWrecS) + tk * 3;
                    sc[0] = a_pre * inv_rms; sc[1] = a_post * inv_rms; sc[2] = a_res * inv_rms;
                }
                // d logits of the token for the workers (they scale to e, split into bf16 hi/lo and build the E tile)
                float* dq = reinterpret_cast<float*>(smem + kOffDl + s * kDlBytes) + tk * kL;
                dq[i4] = dl_pre;
                dq[kN + i4] = dl_post;
                *reinterpret_cast<float4*>(dq + 2 * kN + 4 * i4) = make_float4(d0, d1, d2, d3);
            }
            __syncwarp();
            mbar_arrive(&bar_cd[s]);                        // release: every lane's stores above are visible to the waiter
            if (lane == 0) HVS_TR(it, 6);
            HVS_TICK(6);
        }
        if (HVS_TRACE_ON && p.dbg && lane == 0)
            for (int q = 0; q < 8; ++q) p.dbg[((size_t)blockIdx.x * 4 + cw) * 8 + q] = tacc[q];
        // dalpha: sum the 8 tokens of the warp (row-0 lanes; the others hold 0) in a fixed order; dbias comes from the workers
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) acc_a[k] += __shfl_xor_sync(0xffffffffu, acc_a[k], o);
        }
        {
            float* o = p.cta_accum + ((size_t)blockIdx.x * kCoefWarps + cw) * kAccum;
            if (lane == 0) { o[kL] = acc_a[0]; o[kL + 1] = acc_a[1]; o[kL + 2] = acc_a[2]; }
            if (cw != 0 && lane < kL) o[lane] = 0.f;       // dbias of the CTA goes to row 0 (workers)
        }
      }
    } else {
        // ===================================================== worker warps
        const int w = warp, g = lane >> 2, t = lane & 3;
        // W = bf16(scale * phi) as the A operand of  dx_proj^T = W e^T : m-tile (jj, mt) rows g / g+8 are the
        // channel pair 32w + 4g + 2mt + {0,1} of stream jj, k = logits.  48 registers, resident for the whole kernel.
        uint32_t wf[kN][2][4];
        uint32_t wk[16];
#pragma unroll
        for (int jj = 0; jj < kN; ++jj)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int ka = jj * kC + 32 * w + 4 * g + 2 * mt;
                const float sa = __ldg(p.scale + ka), sb = __ldg(p.scale + ka + 1);
                const float* ra = p.phi + (size_t)ka * kL;
                const float* rb = ra + kL;
                wf[jj][mt][0] = pack_bf16(__ldg(ra + 2 * t) * sa, __ldg(ra + 2 * t + 1) * sa);
                wf[jj][mt][1] = pack_bf16(__ldg(rb + 2 * t) * sb, __ldg(rb + 2 * t + 1) * sb);
                wf[jj][mt][2] = pack_bf16(__ldg(ra + 2 * t + 8) * sa, __ldg(ra + 2 * t + 9) * sa);
                wf[jj][mt][3] = pack_bf16(__ldg(rb + 2 * t + 8) * sb, __ldg(rb + 2 * t + 9) * sb);
                wk[(jj * 2 + mt) * 2] = pack_bf16(__ldg(ra + 2 * t + 16) * sa, __ldg(ra + 2 * t + 17) * sa);
                wk[(jj * 2 + mt) * 2 + 1] = pack_bf16(__ldg(rb + 2 * t + 16) * sb, __ldg(rb + 2 * t + 17) * sb);
            }
        // the K = 8 step (logits 16..23) of the fragments is parked in tensor memory: 4 registers per stream are
        // fetched when needed (tcgen05.ld), 32 registers stay resident
        const uint32_t tm_w = tmem_base + ((uint32_t)(32 * (w & 3)) << 16) + kColW + 16u * (uint32_t)(w >> 2);
        tmem_st16(tm_w, wk);
        tmem_wait_st();
        // this thread's 8 bytes (channels 32w + 4g .. +3) of (token 2t, stream jj); token 2t+1 is the next row with
        // the swizzle bit flipped
        const int cb = w >> 1, hh = w & 1;
        // (stream jj adds the constant 8 KB * jj: an immediate in the load / store)
        const uint32_t offa0 = cb * 1024 + (2 * t) * 128 + (((4 * hh + (g >> 1)) ^ (2 * t)) << 4) + (g & 1) * 8;
        const uint32_t offb0 = (offa0 + 128) ^ 16;
        const uint32_t stage0 = smem_u32(smem);
        float acc_db = 0.f;                               // dbias of logit tid % 24 over tokens tid / 24 (threads < 192)
        const int q = w & 3, jcol = w >> 2;               // tensor-memory lane quadrant / G column group of this warp
        const uint32_t tm_gs = tmem_base + ((uint32_t)(32 * q) << 16) + kColGs + 16u * jcol;

        // G of a tile out of tensor memory into its record: lanes 0..15 of quadrant q hold the rows (dy stream q, half
        // lane / 8, token lane % 8); a warp takes the columns of x stream jcol (both halves), keeps the (half, token)
        // diagonal and adds the halves.
        auto g_tile = [&](int tile, int s) {
            tc_fence_after();
            // column lane % 16 of this lane's row, 8 columns at a time (a select tree on the lane bits; an indexed array
            // would go to the stack, and the stack is an L2 round trip away)
            const uint32_t b0 = lane & 1, b1 = lane & 2, b2 = lane & 4, b3 = lane & 8;
            uint32_t v[8];
            tmem_ld8(tm_gs, v);
            tmem_wait_ld();
            const uint32_t lo8 = sel32(sel32(sel32(v[7], v[6], b0), sel32(v[5], v[4], b0), b1),
                                       sel32(sel32(v[3], v[2], b0), sel32(v[1], v[0], b0), b1), b2);
            tmem_ld8(tm_gs + 8, v);
            tmem_wait_ld();
            const uint32_t hi8 = sel32(sel32(sel32(v[7], v[6], b0), sel32(v[5], v[4], b0), b1),
                                       sel32(sel32(v[3], v[2], b0), sel32(v[1], v[0], b0), b1), b2);
            const uint32_t val = sel32(hi8, lo8, b3);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_gr[0]);
            const float sum = __uint_as_float(val) + __shfl_xor_sync(0xffffffffu, __uint_as_float(val), 8);
            if (lane < 8) reinterpret_cast<float*>(smem + kOffWrec + s * kWrecBytes)[lane * 16 + q * 4 + jcol] = sum;
            __threadfence_block();
            bar_arrive(kBarRec + s, kWorkerThreads + 32);
            if (wtid == 64) HVS_TR(tile, 2);
        };
        // Worker schedule: dx of tile k as soon as its coefficients are done.  Every warp also moves its part of a
        // finished G out of tensor memory the moment it completes -- while they wait, and between
        // the four stream steps of a dx pass -- because that read-out heads the next tiles' coefficient chains.
        // (the poll sits in the dx pass's inner loop: barrier address and parity of the next tile are carried along
        // instead of being derived from the tile index every time)
        int g_next = 0, g_stage = 0;
        uint32_t g_par = 0;
        auto g_poll = [&]() {
            if (g_next < n_local && mbar_test_wait(&bar_gs[g_stage], g_par)) {
                g_tile(g_next, g_stage);
                ++g_next;
                if (++g_stage == kStages) { g_stage = 0; g_par ^= 1u; }
            }
        };
        for (int d_next = 0; d_next < n_local;) {
            {
                // ============ dx for tokens 2t, 2t+1 of tile k, channels 32w + 4g .. +3 of every stream
                const int k = d_next++;
                const int s = k % kStages;
                while (!mbar_try_wait(&bar_cd[s], (uint32_t)(k / kStages) & 1u)) g_poll();   // (acquires the coefficient warp's stores)
                if (wtid == 0) HVS_TR(k, 7);
                const uint32_t sb = stage0 + s * kStageBytes;
                uint8_t* wrec = smem + kOffWrec + s * kWrecBytes;
                const int eidx = wtid;                          // (spreading the 192 elements over all 16 warps, 12 lanes each: +1 %)
                if (wtid < kTok * kL) {
                    // e = alpha_g * inv_rms * d logit of (token, logit) = (idx / 24, idx % 24): bf16 hi for the W e MMA,
                    // hi and lo into the E tile of the dW MMA (rows = logits, K = token | 8 + token); dbias in registers
                    const int etok = eidx / kL, er = eidx - etok * kL;
                    const float dlv = reinterpret_cast<const float*>(smem + kOffDl + s * kDlBytes)[eidx];
                    const float e = dlv * reinterpret_cast<const float*>(wrec + kWrecS)[etok * 3 + (er < kN ? 0 : er < 2 * kN ? 1 : 2)];
                    acc_db += dlv;
                    const __nv_bfloat16 hi = __float2bfloat16_rn(e);
                    const __nv_bfloat16 lo = __float2bfloat16_rn(e - __bfloat162float(hi));
                    uint8_t* dst = smem + kOffEt + s * kEtBytes + kEtRow0 + (er >> 3) * 128 + (er & 7) * 16 + etok * 2;
                    *reinterpret_cast<__nv_bfloat16*>(dst) = hi;
                    *reinterpret_cast<__nv_bfloat16*>(dst + kEtChunk) = lo;
                    fence_proxy_async_smem();               // the E tile is read by the tensor core (async proxy)
                    mbar_arrive(&bar_ed[s]);
                }
                // every thread forms the three bf16 pairs of e it multiplies W with straight from the d logits: no worker-wide
                // barrier between the coefficient warp's hand-off and the dx pass (the 192 threads that build the E tile used to
                // publish packed pairs through shared memory behind a 512-thread barrier: -3.5 % on the kernel without it)
                uint32_t eb0, eb1, eb2;
                {
                    const float* dq = reinterpret_cast<const float*>(smem + kOffDl + s * kDlBytes) + g * kL;
                    const float* sc = reinterpret_cast<const float*>(wrec + kWrecS) + g * 3;
                    const float2 d0 = *reinterpret_cast<const float2*>(dq + 2 * t), d1 = *reinterpret_cast<const float2*>(dq + 2 * t + 8),
                                 d2 = *reinterpret_cast<const float2*>(dq + 2 * t + 16);
                    const float s0 = sc[t < 2 ? 0 : 1], s2 = sc[2];
                    eb0 = pack_bf16(d0.x * s0, d0.y * s0);
                    eb1 = pack_bf16(d1.x * s2, d1.y * s2);
                    eb2 = pack_bf16(d2.x * s2, d2.y * s2);
                }
                const float2 kp = *reinterpret_cast<const float2*>(wrec + kWrecK + 8 * t);
                const u64 kp2 = pk2(kp.x, kp.y);
                uint2 dya[kN], dyb[kN];
                {
#pragma unroll
                for (int ii = 0; ii < kN; ++ii) {
                    dya[ii] = lds64(sb + kHalf + offa0 + ii * 8192);
                    dyb[ii] = lds64(sb + kHalf + offb0 + ii * 8192);
                }
#pragma unroll
                for (int jj = 0; jj < kN; ++jj) {
                    g_poll();
                    const float4* mq = reinterpret_cast<const float4*>(wrec + kWrecM + t * (kMpStride * 4)) + jj * 2;
                    const float4 m01 = mq[0], m23 = mq[1];                        // (M_a[0],M_b[0],M_a[1],M_b[1]) (M_a[2],...)
                    const u64 mp[kN] = {pk2(m01.x, m01.y), pk2(m01.z, m01.w), pk2(m23.x, m23.y), pk2(m23.z, m23.w)};
                    uint32_t w8[4];
                    tmem_ld4(tm_w + 4u * jj, w8);
                    const uint2 xa = lds64(sb + offa0 + jj * 8192);
                    const uint2 xb = lds64(sb + offb0 + jj * 8192);
                    uint32_t oa[2], ob[2];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint32_t xwa = mt ? xa.y : xa.x, xwb = mt ? xb.y : xb.x;
                        const u64 c01i = mul2(kp2, pk2(bf16lo(xwa), bf16lo(xwb)));
                        const u64 c23i = mul2(kp2, pk2(bf16hi(xwa), bf16hi(xwb)));
                        float c[4];
                        upk2(c01i, c[0], c[1]);
                        upk2(c23i, c[2], c[3]);
                        mma_bf16_16816(c, wf[jj][mt][0], wf[jj][mt][1], wf[jj][mt][2], wf[jj][mt][3], eb0, eb1);
                        if (mt == 0) tmem_wait_ld();
                        mma_bf16_1688(c, w8[2 * mt], w8[2 * mt + 1], eb2);
                        u64 c01 = pk2(c[0], c[1]), c23 = pk2(c[2], c[3]);
#pragma unroll
                        for (int ii = 0; ii < kN; ++ii) {
                            const uint32_t da = mt ? dya[ii].y : dya[ii].x, db = mt ? dyb[ii].y : dyb[ii].x;
                            c01 = fma2(mp[ii], pk2(bf16lo(da), bf16lo(db)), c01);
                            c23 = fma2(mp[ii], pk2(bf16hi(da), bf16hi(db)), c23);
                        }
                        upk2(c01, c[0], c[1]);
                        upk2(c23, c[2], c[3]);
                        oa[mt] = pack_bf16(c[0], c[2]);
                        ob[mt] = pack_bf16(c[1], c[3]);
                    }
                    sts64(sb + kHalf + offa0 + jj * 8192, oa[0], oa[1]);
                    sts64(sb + kHalf + offb0 + jj * 8192, ob[0], ob[1]);
                    fence_proxy_async_smem();
                    mbar_arrive(&bar_dxr[s * kN + jj]);
                }
                }
                if (wtid == 0) HVS_TR(k, 8);
            }
        }
        // ============ dbias of this CTA: fold the 8 tokens in a fixed order
        {
            float* red = reinterpret_cast<float*>(smem + kOffPart);
            if (wtid < kTok * kL) red[wtid] = acc_db;
            bar_sync(kBarW, kWorkerThreads);
            if (wtid < kL) {
                float v = 0.f;
#pragma unroll
                for (int tt = 0; tt < kTok; ++tt) v += red[tt * kL + wtid];
                p.cta_accum[(size_t)blockIdx.x * kCoefWarps * kAccum + wtid] = v;
            }
        }
        // ============ dW of this CTA out of tensor memory (every MMA has been committed before the last barrier phase)
        {
            const int last = n_local - 1;
            mbar_wait(&bar_dw[last % kStages], (uint32_t)(last / kStages) & 1u);
            tc_fence_after();
            float* out = p.dw_part + (size_t)blockIdx.x * kRow * kL;
#pragma unroll
            for (int pi = 0; pi < 4; ++pi) {
                const int pr = (w >> 2) * 4 + pi;          // 128-channel block; this lane's row = channel 32q + lane of it
                const int kidx = pr * 128 + 32 * q + lane;
                const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + kColDw + (uint32_t)(pr * 24);
                uint32_t v[8];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    tmem_ld8(ta + 8 * c, v);
                    tmem_wait_ld();
                    float4* o = reinterpret_cast<float4*>(out + (size_t)kidx * kL + 8 * c);
                    o[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
                    o[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWorkers + 1) tmem_dealloc(tmem_base, kTmemCols);
}

long long* g_fused_dbg = nullptr;

inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct FusedWs {
    float* dw_part; float* cta_accum;
    size_t total;
};
FusedWs carve(void* base, int ctas) {
    FusedWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return reinterpret_cast<uint8_t*>(base) + o; };
    w.dw_part = reinterpret_cast<float*>(take((size_t)ctas * kRow * kL * 4));
    w.cta_accum = reinterpret_cast<float*>(take((size_t)ctas * kCoefWarps * kAccum * 4));
    w.total = off;
    return w;
}

}  // namespace
}  // namespace hvs

#ifdef HVS_FUSED_TRACE
// development aid (trace builds only, tools/time_fused.py): device buffer [SMs, 3, 8] int64 receiving the
// coefficient warps' cycle counters (NULL = off).  Not part of the shipped ABI.
extern "C" int hvs_debug_fused_timing(void* device_buffer, int mode) {
    hvs::g_fused_dbg = reinterpret_cast<long long*>(device_buffer);
    (void)mode;
    return HVS_OK;
}
#endif

extern "C" size_t hvs_mhc_stream_bwd_saved_workspace(int64_t T, int n, int C) {
    using namespace hvs;
    if (T < 0 || n != kN || C != kC) return 0;
    return carve(nullptr, sm_count()).total;
}

extern "C" int hvs_mhc_stream_bwd_saved(const void* x, const void* dy, const float* saved, const float* phi,
                                        const float* bias, const float* alpha, const float* scale, void* dx,
                                        float* dphi, float* dbias, float* dalpha, float* dscale, int64_t T, int n,
                                        int C, int sk_iters, float eps_rms, float eps_sk, uint32_t flags,
                                        void* workspace, size_t workspace_bytes, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (T < 0) return HVS_ERR_BAD_ARG;
    if (n != kN || C != kC || sk_iters < 0 || sk_iters > kMaxIters) return HVS_ERR_UNSUPPORTED;
    if (flags & HVS_MHC_SPLIT_PHI) return HVS_ERR_UNSUPPORTED;
    if (!phi || !bias || !alpha || !scale || !dphi || !dbias || !dalpha || !dscale) return HVS_ERR_BAD_ARG;
    if (T > 0 && (!x || !dy || !dx || !saved)) return HVS_ERR_BAD_ARG;
    if (T >= (int64_t)1 << 31) return HVS_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
         reinterpret_cast<uintptr_t>(saved)) & 15)
        return HVS_ERR_ALIGNMENT;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return HVS_ERR_ALIGNMENT;
    const int sms = sm_count();
    const FusedWs ws = carve(workspace, sms);
    if (workspace_bytes < ws.total) return HVS_ERR_WORKSPACE;
    int grid = 0;
    if (T > 0) {
        CUtensorMap tx, tdy, tdx;
        int rc = make_tmap_bf16_streams4d(&tx, x, (uint64_t)T, kTok);
        if (rc) return rc;
        rc = make_tmap_bf16_streams4d(&tdy, dy, (uint64_t)T, kTok);
        if (rc) return rc;
        rc = make_tmap_bf16_streams4d(&tdx, dx, (uint64_t)T, kTok, 1);      // dx leaves one stream (8 KB) at a time
        if (rc) return rc;
        const bool adaptive = (flags & HVS_MHC_ADAPTIVE_ITERS) != 0;
        if (adaptive) HVS_SET_MAX_SMEM(mhc_stream_bwd_fused_kernel<true>, kSmemBytes);
        else HVS_SET_MAX_SMEM(mhc_stream_bwd_fused_kernel<false>, kSmemBytes);
        FusedParams p;
        p.phi = phi; p.bias = bias; p.alpha = alpha; p.scale = scale; p.saved = saved;
        p.dw_part = ws.dw_part; p.cta_accum = ws.cta_accum;
        p.dbg = g_fused_dbg;
        p.T = T;
        p.num_tiles = (int)((T + kTok - 1) / kTok);
        p.sk_iters = sk_iters; p.eps_rms = eps_rms; p.eps_sk = eps_sk;
        grid = p.num_tiles < sms ? p.num_tiles : sms;
        timer_begin(1, stream);
        if (adaptive) mhc_stream_bwd_fused_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(tx, tdy, tdx, p);
        else mhc_stream_bwd_fused_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(tx, tdy, tdx, p);
        timer_end(1, stream);
        count_launch();
        const int rc2 = launch_status();
        if (rc2) return rc2;
    }
    timer_begin(3, stream);
    const int rc3 = launch_bwd_finalize(ws.dw_part, grid, ws.cta_accum, kCoefWarps * grid, phi, scale, dphi, dscale, dbias, dalpha, stream);
    timer_end(3, stream);
    return rc3;
}
