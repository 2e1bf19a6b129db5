// Greedy NMS, bit-exact against the reference's two implementations:
//   class-agnostic  YOLODetectionHead.non_max_suppression / compute_iou   src/models/yolo_head.py:678-755
//   class-aware     NMSFilter.apply / _standard_nms / _compute_iou        src/inference/postprocessing.py:505-607, 772-802
// and the two-stage multi-scale merge YOLODetectionHead.post_process      src/models/yolo_head.py:571-676.
//
// Two kernels.  Sets of 64 < N <= 24000 candidates (both stages of post_process): nms_sorted_kernel below -- one stable
// sort in shared memory (radix above 1024 passing candidates, one rank-by-counting pass below), then every candidate is
// tested only against the boxes already kept.  Smaller sets: nms_kernel.
// One CTA per candidate set.  Both reference loops keep at most max_det boxes in descending score
// order and never look back, so instead of a sort + N x N mask the CTA repeats, at most max_det
// times: (1) block-wide arg-max over the still-alive scores (ties -> lower index; scores cached in
// shared memory, a suppressed candidate becomes -inf), (2) one IoU of the winner against every alive
// candidate, 32 candidates per warp, the warp's verdicts gathered with __ballot_sync into the 32-bit
// alive word of that group.  Work is O(kept x N) integer/compare work on L2-resident boxes.
//
// IoU is evaluated with explicitly rounded fp32 intrinsics (no FMA contraction), in the reference's
// operation order: inter / (((a1 + a2) - inter) + 1e-6).  max/min/clamp propagate NaN like torch.
#include <math.h>

#include "common.cuh"

namespace hvs {
namespace {

constexpr int kMaxGroups = 4;

struct NmsGroup {
    const float* boxes;       // [problems, stride, 4]
    const float* scores;      // [problems, stride]
    const int64_t* classes;   // [problems, stride] or null
    int64_t stride;           // elements between consecutive problems of this group
    int n;                    // candidates per problem (ignored when counts/offsets are given)
};

struct NmsArgs {
    NmsGroup g[kMaxGroups];
    int num_groups;            // problem p -> group p % num_groups, batch item p / num_groups
    int batch;                 // problems per group
    int order[kMaxGroups];     // groups by descending n: CTA x works on group order[x / batch], item x % batch -- the big
                               // candidate sets start first and the small ones fill in behind them (in problem order the
                               // 64 + 128 sets of a batch of 64 ran as two waves with big sets in both)
    const int64_t* offsets;    // optional [P+1]: problem p = [offsets[p], offsets[p+1]) of group 0
    const int32_t* counts;     // optional [P]: candidates of problem p (<= g.n)
    float score_thr, iou_thr;
    int max_det, mode;
    int64_t* keep_idx;         // [P, max_det] index into the compacted (score > thr) list, may be null
    int64_t* keep_src;         // [P, max_det] index into the problem's candidates
    int32_t* keep_count;       // [P]
};

__device__ __forceinline__ int problem_of_cta(const NmsArgs& a) {
    const int x = blockIdx.x;
    if (a.offsets || a.num_groups <= 1) return x;
    const int slot = x / a.batch;
    return (x - slot * a.batch) * a.num_groups + a.order[slot];
}

__device__ __forceinline__ float max_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
__device__ __forceinline__ float clamp0_nan(float a) { return a < 0.f ? 0.f : a; }

__device__ __forceinline__ float4 load_box(const float* boxes, int64_t i, bool cxcywh) {
    float4 b = __ldg(reinterpret_cast<const float4*>(boxes) + i);
    if (cxcywh) {   // postprocessing.py:540-549
        const float hw = __fmul_rn(b.z, 0.5f), hh = __fmul_rn(b.w, 0.5f);
        b = make_float4(__fsub_rn(b.x, hw), __fsub_rn(b.y, hh), __fadd_rn(b.x, hw), __fadd_rn(b.y, hh));
    }
    return b;
}

__device__ __forceinline__ float iou_ref(const float4& a, float area_a, const float4& b) {
    const float ix1 = max_nan(a.x, b.x), iy1 = max_nan(a.y, b.y);
    const float ix2 = min_nan(a.z, b.z), iy2 = min_nan(a.w, b.w);
    const float iw = clamp0_nan(__fsub_rn(ix2, ix1)), ih = clamp0_nan(__fsub_rn(iy2, iy1));
    const float inter = __fmul_rn(iw, ih);
    const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float uni = __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), 1e-6f);
    return __fdiv_rn(inter, uni);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) nms_kernel(const NmsArgs a) {
    extern __shared__ float s_score[];                 // [n] alive score or -inf
    __shared__ float s_wbest[THREADS / 32];
    __shared__ int s_wbesti[THREADS / 32];
    __shared__ int s_pick;
    const int p = problem_of_cta(a);
    const int gi = a.offsets ? 0 : p % a.num_groups;
    const int bi = a.offsets ? 0 : p / a.num_groups;
    const NmsGroup& grp = a.g[gi];
    int64_t start = (int64_t)bi * grp.stride;
    int n = grp.n;
    if (a.offsets) { start = a.offsets[p]; n = (int)(a.offsets[p + 1] - start); }
    if (a.counts) n = a.counts[p];
    const float* boxes = grp.boxes + start * 4;
    const float* scores = grp.scores + start;
    const int64_t* classes = grp.classes ? grp.classes + start : nullptr;
    const bool class_aware = (a.mode & 15) == HVS_NMS_CLASS_AWARE;
    const bool cxcywh = class_aware && !(a.mode & HVS_NMS_BOXES_XYXY);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < n; i += THREADS) {
        const float s = scores[i];
        s_score[i] = (s > a.score_thr) ? s : -INFINITY;     // strict >, NaN never passes (yolo_head.py:605)
    }
    __syncthreads();

    int kept = 0;
    while (kept < a.max_det) {
        // ---- (1) arg-max over alive candidates, ties -> lower index
        float best = -INFINITY;
        int besti = 0x7fffffff;
        for (int i = tid; i < n; i += THREADS) {
            const float s = s_score[i];
            if (s > best) { best = s; besti = i; }          // ascending i: first maximum wins
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) { s_wbest[warp] = best; s_wbesti[warp] = besti; }
        __syncthreads();
        if (warp == 0) {
            best = lane < THREADS / 32 ? s_wbest[lane] : -INFINITY;
            besti = lane < THREADS / 32 ? s_wbesti[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
            }
            if (lane == 0) s_pick = (best == -INFINITY) ? -1 : besti;   // +inf scores are legal, -inf is "dead"
        }
        __syncthreads();
        const int pick = s_pick;
        if (pick < 0) break;
        if (tid == 0) {
            a.keep_src[(int64_t)p * a.max_det + kept] = pick;
            s_score[pick] = -INFINITY;
        }
        ++kept;
        if (kept >= a.max_det) break;                        // yolo_head.py:710-711
        // ---- (2) suppress against the winner
        const float4 wb = load_box(boxes, pick, cxcywh);
        const float warea = __fmul_rn(__fsub_rn(wb.z, wb.x), __fsub_rn(wb.w, wb.y));
        const int64_t wcls = class_aware ? classes[pick] : 0;
        __syncthreads();                                     // winner's own slot is dead before the sweep
        for (int i0 = warp * 32; i0 < n; i0 += THREADS) {
            const int i = i0 + lane;
            bool alive = i < n && s_score[i] != -INFINITY;
            if (__ballot_sync(0xffffffffu, alive) == 0u) continue;
            bool kill = false;
            if (alive) {
                if (class_aware) {
                    if (classes[i] == wcls) kill = iou_ref(wb, warea, load_box(boxes, i, cxcywh)) > a.iou_thr;  // postprocessing.py:594
                } else {
                    kill = !(iou_ref(wb, warea, load_box(boxes, i, false)) < a.iou_thr);                      // yolo_head.py:727
                }
            }
            const unsigned dead = __ballot_sync(0xffffffffu, kill);
            if ((dead >> lane) & 1u) s_score[i] = -INFINITY;
        }
        __syncthreads();
    }
    if (tid == 0) a.keep_count[p] = kept;
    // index into the compacted list of candidates that passed the score threshold
    if (a.keep_idx != nullptr) {
        __syncthreads();
        for (int r = warp; r < kept; r += THREADS / 32) {
            const int src = (int)a.keep_src[(int64_t)p * a.max_det + r];
            int cnt = 0;
            for (int i0 = 0; i0 < src; i0 += 32) {
                const int i = i0 + lane;
                cnt += __popc(__ballot_sync(0xffffffffu, i < src && scores[i] > a.score_thr));
            }
            if (lane == 0) a.keep_idx[(int64_t)p * a.max_det + r] = cnt;
        }
    }
}

// ---- large candidate sets: sort once, then test a candidate only against the boxes already KEPT ------------------------
// The arg-max kernel above does kept x N IoUs (one sweep of every alive candidate per kept box: 100 x 19 200 per problem
// when a random-init model lets every candidate through).  Greedy NMS only ever lets KEPT boxes suppress,
// so a candidate's fate depends on the kept boxes that precede it in score order alone: sort the passing candidates
// once (stable LSD radix sort on the score bits, ties -> lower index, 4-bit digits, all in shared memory), then walk
// them in order 1024 at a time -- every thread tests its candidate against the boxes kept so far, warp 0 then resolves
// the survivors of the chunk in order with ballots.  IoUs drop from kept x N to about (candidates examined) x kept, and
// the walk stops at max_det.  Same semantics, same IoU arithmetic, same outputs as nms_kernel.
constexpr int kSortThreads = 1024;
constexpr int kSortMaxN = 24000;

__device__ __forceinline__ uint32_t desc_key(float s) {                // ascending order of this key = descending score
    const uint32_t u = __float_as_uint(s);
    return ~((u & 0x80000000u) ? ~u : (u | 0x80000000u));
}

__device__ __forceinline__ bool suppresses(const float4& kb, float karea, int64_t kcls, const float4& cb, int64_t ccls,
                                           bool class_aware, float thr) {
    if (class_aware) return kcls == ccls && iou_ref(kb, karea, cb) > thr;     // postprocessing.py:594
    return !(iou_ref(kb, karea, cb) < thr);                                   // yolo_head.py:727
}

__global__ void __launch_bounds__(kSortThreads) nms_sorted_kernel(const NmsArgs a) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ int s_warp_tot[32];
    __shared__ int s_total, s_kept;
    __shared__ uint32_t s_alive[32], s_sup[32], s_pair[32];
    const int p = problem_of_cta(a);
    const int gi = a.offsets ? 0 : p % a.num_groups;
    const int bi = a.offsets ? 0 : p / a.num_groups;
    const NmsGroup& grp = a.g[gi];
    int64_t start = (int64_t)bi * grp.stride;
    int n = grp.n;
    if (a.offsets) { start = a.offsets[p]; n = (int)(a.offsets[p + 1] - start); }
    if (a.counts) n = a.counts[p];
    const float* boxes = grp.boxes + start * 4;
    const float* scores = grp.scores + start;
    const int64_t* classes = grp.classes ? grp.classes + start : nullptr;
    const bool class_aware = (a.mode & 15) == HVS_NMS_CLASS_AWARE;
    const bool cxcywh = class_aware && !(a.mode & HVS_NMS_BOXES_XYXY);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cap = (n + 1) & ~1;
    // shared memory: keys [cap] u32 | order A [cap] u16 | order B [cap] u16 | digit counters [16][1024] u16 | kept boxes
    uint32_t* keys = reinterpret_cast<uint32_t*>(s_raw);
    uint16_t* ord_a = reinterpret_cast<uint16_t*>(keys + cap);
    uint16_t* ord_b = ord_a + cap;
    uint16_t* cnt = ord_b + cap;
    float4* kbox = reinterpret_cast<float4*>(cnt + 16 * kSortThreads);
    float* karea = reinterpret_cast<float*>(kbox + a.max_det);
    int64_t* kcls = reinterpret_cast<int64_t*>(karea + ((a.max_det + 1) & ~1));

    // ---- 1. order-preserving compaction of the candidates with score > thr (strict; NaN never passes, yolo_head.py:605)
    const int per = (n + kSortThreads - 1) / kSortThreads;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    int c = 0;
    for (int i = lo; i < hi; ++i) c += scores[i] > a.score_thr;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = s_warp_tot[lane], w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
        s_warp_tot[lane] = w - v;
        if (lane == 31) s_total = w;
    }
    __syncthreads();
    int pos = s_warp_tot[warp] + incl - c;
    for (int i = lo; i < hi; ++i) {
        const float sc = scores[i];
        if (sc > a.score_thr) { keys[i] = desc_key(sc); ord_a[pos++] = (uint16_t)i; }
    }
    const int m = s_total;
    if (tid == 0) s_kept = 0;
    __syncthreads();

    // ---- 2. stable LSD radix sort of the compacted indices by key (8 passes of 4 bits)
    uint16_t* src = ord_a;
    uint16_t* dst = ord_b;
    const int mper = (m + kSortThreads - 1) / kSortThreads;
    const int mlo = min(tid * mper, m), mhi = min(mlo + mper, m);
    if (m > 1 && m <= kSortThreads) {
        // up to one candidate per thread (the cross-scale stage: at most 3 x max_det): its place in the order is the number of
        // candidates that sort before it -- smaller key, or the same key earlier in the list (= lower index).  One pass.
        if (tid < m) {
            const uint32_t mine = keys[src[tid]];
            int rank = 0;
            for (int q = 0; q < m; ++q) {
                const uint32_t k = keys[src[q]];
                rank += (k < mine || (k == mine && q < tid)) ? 1 : 0;
            }
            dst[rank] = src[tid];
        }
        __syncthreads();
        uint16_t* t = src; src = dst; dst = t;
    }
    for (int shift = 0; shift < 32 && m > kSortThreads; shift += 4) {
#pragma unroll
        for (int d = 0; d < 16; ++d) cnt[d * kSortThreads + tid] = 0;
        for (int q = mlo; q < mhi; ++q) cnt[((keys[src[q]] >> shift) & 15u) * kSortThreads + tid] += 1;
        __syncthreads();
        // exclusive scan of the 16 x 1024 counters in (digit, thread) order: thread t owns counters [16 t, 16 t + 16)
        int local[16];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) { local[k] = cnt[tid * 16 + k]; sum += local[k]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) s_warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int v = s_warp_tot[lane], w = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
            s_warp_tot[lane] = w - v;
        }
        __syncthreads();
        int run = s_warp_tot[warp] + inc - sum;
#pragma unroll
        for (int k = 0; k < 16; ++k) { cnt[tid * 16 + k] = (uint16_t)run; run += local[k]; }
        __syncthreads();
        for (int q = mlo; q < mhi; ++q) {
            const uint16_t v = src[q];
            const uint32_t d = (keys[v] >> shift) & 15u;
            dst[cnt[d * kSortThreads + tid]++] = v;
        }
        __syncthreads();
        uint16_t* t = src; src = dst; dst = t;
    }

    // ---- 3. walk the sorted candidates, 1024 at a time.  A chunk: every thread tests its candidate against the boxes kept
    //      in earlier chunks; then the chunk's 32 groups of 32 are resolved in order, each in two block-wide steps:
    //        A  all 1024 threads: candidate `lane` of the group against the boxes kept earlier in this chunk (keep w, w + 32, ..
    //           for warp w) and against candidate `warp` of the same group (the 32 x 32 pair matrix, one IoU per thread)
    //        B  warp 0: greedy over the group's alive word with the pair masks -- a few integer instructions per kept box
    //      (one warp doing both with an IoU per kept box on its dependent chain was 70 % of the kernel: ~100 us per set).
    __syncthreads();                                         // the digit counters are dead: the chunk buffers live there
    float4* cbox = reinterpret_cast<float4*>(cnt);
    int64_t* ccls = reinterpret_cast<int64_t*>(cbox + kSortThreads);
    for (int base = 0; base < m; base += kSortThreads) {
        const int kept0 = s_kept;
        if (kept0 >= a.max_det) break;
        {
            const int q = base + tid;
            bool alive = q < m;
            float4 mb = make_float4(0.f, 0.f, 0.f, 0.f);
            int64_t mc = 0;
            if (alive) {
                const int me = src[q];
                mb = load_box(boxes, me, cxcywh);
                mc = class_aware ? classes[me] : 0;
                for (int k = 0; k < kept0 && alive; ++k) alive = !suppresses(kbox[k], karea[k], kcls[k], mb, mc, class_aware, a.iou_thr);
            }
            cbox[tid] = mb;
            ccls[tid] = mc;
            const uint32_t word = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) s_alive[warp] = word;
        }
        __syncthreads();
        for (int g = 0; g < 32; ++g) {
            const int kept = s_kept;                             // (block-uniform: written before the last barrier)
            if (kept >= a.max_det) break;
            const uint32_t aw = s_alive[g];
            if (aw == 0u) continue;
            const float4 cb = cbox[g * 32 + lane];
            const int64_t cc = ccls[g * 32 + lane];
            const bool mine = (aw >> lane) & 1u;
            bool sup = false;
            if (mine)
                for (int k = kept0 + warp; k < kept && !sup; k += 32) sup = suppresses(kbox[k], karea[k], kcls[k], cb, cc, class_aware, a.iou_thr);
            bool pair = false;
            if (mine && lane > warp && ((aw >> warp) & 1u)) {
                const float4 wb = cbox[g * 32 + warp];
                const float warea = __fmul_rn(__fsub_rn(wb.z, wb.x), __fsub_rn(wb.w, wb.y));
                pair = suppresses(wb, warea, ccls[g * 32 + warp], cb, cc, class_aware, a.iou_thr);
            }
            const uint32_t supw = __ballot_sync(0xffffffffu, sup), pairw = __ballot_sync(0xffffffffu, pair);
            if (lane == 0) { s_sup[warp] = supw; s_pair[warp] = pairw; }
            __syncthreads();
            if (warp == 0) {
                uint32_t alive_w = aw & ~__reduce_or_sync(0xffffffffu, s_sup[lane]);
                uint32_t keep_w = 0u;
                int kk = kept;
                while (alive_w != 0u && kk < a.max_det) {
                    const int ld = __ffs(alive_w) - 1;           // best remaining candidate of the group: keep it
                    keep_w |= 1u << ld;
                    alive_w &= ~(s_pair[ld] | (1u << ld));
                    ++kk;
                }
                if ((keep_w >> lane) & 1u) {
                    const int r = kept + __popc(keep_w & ((1u << lane) - 1u));
                    kbox[r] = cb;
                    karea[r] = __fmul_rn(__fsub_rn(cb.z, cb.x), __fsub_rn(cb.w, cb.y));
                    kcls[r] = cc;
                    a.keep_src[(int64_t)p * a.max_det + r] = src[base + g * 32 + lane];
                }
                if (lane == 0) s_kept = kk;
            }
            __syncthreads();
        }
    }
    const int kept = s_kept;
    if (tid == 0) a.keep_count[p] = kept;
    if (a.keep_idx != nullptr) {      // index into the compacted (score > thr) list
        for (int r = warp; r < kept; r += kSortThreads / 32) {
            const int sidx = (int)a.keep_src[(int64_t)p * a.max_det + r];
            int cn = 0;
            for (int i0 = 0; i0 < sidx; i0 += 32) {
                const int i = i0 + lane;
                cn += __popc(__ballot_sync(0xffffffffu, i < sidx && scores[i] > a.score_thr));
            }
            if (lane == 0) a.keep_idx[(int64_t)p * a.max_det + r] = cn;
        }
    }
}

inline size_t sorted_smem_bytes(int64_t max_n, int max_det) {
    const size_t cap = ((size_t)max_n + 1) & ~(size_t)1;
    return cap * 4 + cap * 2 * 2 + 16 * kSortThreads * 2 + (size_t)max_det * 16 + (((size_t)max_det + 1) & ~(size_t)1) * 4 + (size_t)max_det * 8 + 64;
}

int launch_nms(const NmsArgs& a, int num_problems, int64_t max_n, cudaStream_t stream) {
    if (num_problems == 0) return HVS_OK;
    if (max_n > 49152) return HVS_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(max_n > 0 ? max_n : 1) * sizeof(float);
    constexpr size_t kSortSmemMax = 232448 - 1024;            // the kernel also has static shared memory
    // (sets of up to 64 candidates stay on the arg-max kernel: a handful of rounds there; above that the arg-max kernel's
    //  max_det block-wide rounds -- 150 us for the 300-candidate cross-scale stage -- lose to one sort + one walk)
    if (max_n > 64 && max_n <= kSortMaxN && sorted_smem_bytes(max_n, a.max_det) <= kSortSmemMax) {
        const size_t sb = sorted_smem_bytes(max_n, a.max_det);
        HVS_SET_MAX_SMEM(nms_sorted_kernel, (int)kSortSmemMax);
        nms_sorted_kernel<<<num_problems, kSortThreads, sb, stream>>>(a);
        count_launch();
        return launch_status();
    }
    if (max_n <= 4096) {
        nms_kernel<256><<<num_problems, 256, smem, stream>>>(a);
    } else {
        HVS_SET_MAX_SMEM(nms_kernel<1024>, 49152 * 4);
        nms_kernel<1024><<<num_problems, 1024, smem, stream>>>(a);
    }
    count_launch();
    return launch_status();
}

// ---- post_process glue -------------------------------------------------------------------------
struct GatherArgs {
    NmsGroup g[kMaxGroups];
    int num_groups, max_det;
    const int64_t* s1_src;     // [B*S, max_det]
    const int32_t* s1_count;   // [B*S]
    float* cat_boxes; float* cat_scores; int64_t* cat_labels; int32_t* cat_count;   // [B, S*max_det(,4)]
};

// concatenate the per-scale survivors in scale order, kept order inside a scale (yolo_head.py:646-654)
__global__ void gather_scales_kernel(const GatherArgs a) {
    const int b = blockIdx.x;
    const int cap = a.num_groups * a.max_det;
    int base = 0;
    for (int s = 0; s < a.num_groups; ++s) {
        const int p = b * a.num_groups + s;
        const int cnt = a.s1_count[p];
        const NmsGroup& g = a.g[s];
        for (int r = threadIdx.x; r < cnt; r += blockDim.x) {
            const int64_t src = (int64_t)b * g.stride + a.s1_src[(int64_t)p * a.max_det + r];
            const int64_t dst = (int64_t)b * cap + base + r;
            reinterpret_cast<float4*>(a.cat_boxes)[dst] = __ldg(reinterpret_cast<const float4*>(g.boxes) + src);
            a.cat_scores[dst] = g.scores[src];
            a.cat_labels[dst] = g.classes[src];
        }
        base += cnt;
    }
    if (threadIdx.x == 0) a.cat_count[b] = base;
}

__global__ void gather_final_kernel(const float* cat_boxes, const float* cat_scores, const int64_t* cat_labels,
                                    const int64_t* s2_src, const int32_t* s2_count, int cap, int max_det,
                                    float* det_boxes, float* det_scores, int64_t* det_labels, int32_t* det_count) {
    const int b = blockIdx.x;
    const int cnt = s2_count[b];
    for (int r = threadIdx.x; r < max_det; r += blockDim.x) {
        const int64_t dst = (int64_t)b * max_det + r;
        if (r < cnt) {
            const int64_t src = (int64_t)b * cap + s2_src[dst];
            reinterpret_cast<float4*>(det_boxes)[dst] = reinterpret_cast<const float4*>(cat_boxes)[src];
            det_scores[dst] = cat_scores[src];
            det_labels[dst] = cat_labels[src];
        } else {
            reinterpret_cast<float4*>(det_boxes)[dst] = make_float4(0.f, 0.f, 0.f, 0.f);
            det_scores[dst] = 0.f;
            det_labels[dst] = -1;
        }
    }
    if (threadIdx.x == 0) det_count[b] = cnt;
}

inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct PostWs {
    float* cat_boxes; float* cat_scores; int64_t* cat_labels; int32_t* cat_count;
    int64_t* s1_src; int32_t* s1_count; int64_t* s2_src; int32_t* s2_count;
    size_t total;
};

PostWs carve(void* base, int B, int S, int max_det) {
    PostWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return reinterpret_cast<uint8_t*>(base) + o; };
    const size_t cap = (size_t)S * max_det;
    w.cat_boxes = reinterpret_cast<float*>(take((size_t)B * cap * 16));
    w.cat_scores = reinterpret_cast<float*>(take((size_t)B * cap * 4));
    w.cat_labels = reinterpret_cast<int64_t*>(take((size_t)B * cap * 8));
    w.cat_count = reinterpret_cast<int32_t*>(take((size_t)B * 4));
    w.s1_src = reinterpret_cast<int64_t*>(take((size_t)B * S * max_det * 8));
    w.s1_count = reinterpret_cast<int32_t*>(take((size_t)B * S * 4));
    w.s2_src = reinterpret_cast<int64_t*>(take((size_t)B * max_det * 8));
    w.s2_count = reinterpret_cast<int32_t*>(take((size_t)B * 4));
    w.total = off;
    return w;
}

}  // namespace
}  // namespace hvs

extern "C" int hvs_nms(const float* boxes, const float* scores, const int64_t* classes, const int64_t* offsets,
                       int num_problems, int64_t max_n, float score_thr, float iou_thr, int max_det, int mode,
                       int64_t* keep_idx, int64_t* keep_src, int32_t* keep_count, void* stream) {
    using namespace hvs;
    if (num_problems < 0 || max_n < 0 || max_det <= 0 || !keep_src || !keep_count || !offsets) return HVS_ERR_BAD_ARG;
    if (num_problems > 0 && max_n > 0 && (!boxes || !scores)) return HVS_ERR_BAD_ARG;
    if ((mode & 15) == HVS_NMS_CLASS_AWARE && max_n > 0 && !classes) return HVS_ERR_BAD_ARG;
    if ((mode & 15) > HVS_NMS_CLASS_AWARE) return HVS_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(boxes) & 15) return HVS_ERR_ALIGNMENT;
    NmsArgs a{};
    a.g[0] = NmsGroup{boxes, scores, classes, 0, 0};
    a.num_groups = 1; a.batch = num_problems > 0 ? num_problems : 1; a.order[0] = 0;
    a.offsets = offsets;
    a.counts = nullptr;
    a.score_thr = score_thr; a.iou_thr = iou_thr; a.max_det = max_det; a.mode = mode;
    a.keep_idx = keep_idx; a.keep_src = keep_src; a.keep_count = keep_count;
    return launch_nms(a, num_problems, max_n, (cudaStream_t)stream);
}

extern "C" size_t hvs_post_process_workspace(int B, int num_scales, int max_det) {
    if (B <= 0 || num_scales <= 0 || max_det <= 0) return 0;
    return hvs::carve(nullptr, B, num_scales, max_det).total;
}

extern "C" int hvs_post_process(const float* const* boxes_host, const float* const* class_scores_host,
                                const int64_t* const* class_idx_host, const int* n_per_scale_host, int num_scales,
                                int B, float conf_thr, float iou_thr, int max_det, float* det_boxes,
                                float* det_scores, int64_t* det_labels, int32_t* det_count, void* workspace,
                                size_t workspace_bytes, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!boxes_host || !class_scores_host || !class_idx_host || !n_per_scale_host || !det_boxes || !det_scores ||
        !det_labels || !det_count || B < 0 || max_det <= 0)
        return HVS_ERR_BAD_ARG;
    if (num_scales <= 0 || num_scales > kMaxGroups) return HVS_ERR_UNSUPPORTED;
    if (B == 0) return HVS_OK;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return HVS_ERR_ALIGNMENT;
    const PostWs w = carve(workspace, B, num_scales, max_det);
    if (workspace_bytes < w.total) return HVS_ERR_WORKSPACE;
    // stage 1: per (image, scale), class-agnostic, score > conf (yolo_head.py:605, :625-629)
    NmsArgs a{};
    int64_t max_n = 0;
    for (int s = 0; s < num_scales; ++s) {
        if (!boxes_host[s] || !class_scores_host[s] || !class_idx_host[s] || n_per_scale_host[s] < 0) return HVS_ERR_BAD_ARG;
        a.g[s] = NmsGroup{boxes_host[s], class_scores_host[s], class_idx_host[s], n_per_scale_host[s], n_per_scale_host[s]};
        if (n_per_scale_host[s] > max_n) max_n = n_per_scale_host[s];
    }
    a.num_groups = num_scales; a.batch = B;
    for (int s = 0; s < num_scales; ++s) a.order[s] = s;
    for (int i = 1; i < num_scales; ++i)            // insertion sort, descending n, stable
        for (int j = i; j > 0 && a.g[a.order[j]].n > a.g[a.order[j - 1]].n; --j) { const int t = a.order[j]; a.order[j] = a.order[j - 1]; a.order[j - 1] = t; }
    a.score_thr = conf_thr; a.iou_thr = iou_thr; a.max_det = max_det; a.mode = HVS_NMS_AGNOSTIC;
    a.keep_idx = nullptr; a.keep_src = w.s1_src; a.keep_count = w.s1_count;
    int rc = launch_nms(a, B * num_scales, max_n, stream);
    if (rc) return rc;
    // concatenate survivors in scale order (:646-654)
    GatherArgs ga{};
    for (int s = 0; s < num_scales; ++s) ga.g[s] = a.g[s];
    ga.num_groups = num_scales; ga.max_det = max_det;
    ga.s1_src = w.s1_src; ga.s1_count = w.s1_count;
    ga.cat_boxes = w.cat_boxes; ga.cat_scores = w.cat_scores; ga.cat_labels = w.cat_labels; ga.cat_count = w.cat_count;
    gather_scales_kernel<<<B, 128, 0, stream>>>(ga);
    count_launch();
    // stage 2: NMS across scales on the concatenation, no score threshold (:658-662)
    const int cap = num_scales * max_det;
    NmsArgs b2{};
    b2.g[0] = NmsGroup{w.cat_boxes, w.cat_scores, w.cat_labels, cap, cap};
    b2.num_groups = 1; b2.batch = B; b2.order[0] = 0;
    b2.counts = w.cat_count;
    b2.score_thr = -INFINITY; b2.iou_thr = iou_thr; b2.max_det = max_det; b2.mode = HVS_NMS_AGNOSTIC;
    b2.keep_idx = nullptr; b2.keep_src = w.s2_src; b2.keep_count = w.s2_count;
    rc = launch_nms(b2, B, cap, stream);
    if (rc) return rc;
    gather_final_kernel<<<B, 128, 0, stream>>>(w.cat_boxes, w.cat_scores, w.cat_labels, w.s2_src, w.s2_count, cap,
                                                max_det, det_boxes, det_scores, det_labels, det_count);
    count_launch();
    return launch_status();
}
