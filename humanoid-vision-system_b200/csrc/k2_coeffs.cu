// Static (parameter-only) coefficients of the reference-literal module for ALL layers of a model in ONE launch:
//   ManifoldHyperConnection.constrained_matrices   src/models/manifold_layers.py:205-221
//   SinkhornKnoppProjection.forward (2-D input)                                :32-93
// and the matching backward (d H_pre / d H_post / d H_res -> d raw parameters), which replaces autograd's unrolled
// 20-iteration graph (120 tiny kernels and 60 host syncs per layer per step in the reference).
//
// One cooperative kernel, one CTA per SM.  Every D x D matrix is cut into row slabs of about equal element count
// (a 1792 x 1792 matrix becomes ~50 slabs spread over ~50 SMs, a 32 x 32 one is a single slab), so a large matrix is
// iterated by many CTAs.  The iteration runs on the scalings:  P_k = diag(u_k) K diag(v_k),  K = softmax(raw) * m,
//   u_k = u_{k-1} / (u_{k-1} * (K v_{k-1}) + eps),   v_k = v_{k-1} / (v_{k-1} * (K^T u_k) + eps)
// which is the reference's P / (rowsum + eps), P / (colsum + eps) written for the factors; K (fp32, L2-resident, in
// the H_res output buffer) is only READ during the iterations.  Per iteration: one pass over K (row dot products
// with v, then column partial sums of K^T u in registers), one grid-wide barrier.  All reductions have a fixed order
// (lane stride, butterfly, warp order, slab order): results are bitwise reproducible.
// The same launch applies the sigmoid gates and writes the bf16, transposed, K-padded copies of the three matrices
// that the token-path GEMMs (k2_gemm.cu) consume as their B operands.
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace hvs {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxD = 4096;

struct Slab {
    int job, row0, nrows;
    int part_off;      // float offset of this slab's column partials
};
struct JobDev {
    hvs_coeff_job j;
    int first_slab, nslabs;
    int vec_off;       // float offset of this job's u / v / scratch vectors (each D long)
    int mat_off;       // float offset of this job's D x D scratch matrix (backward)
    hvs_coeff_grad g;  // backward only
};
struct TransOp {       // dst_t[c][r] = bf16(f(src[r][c])), dst_f[r][c] = f(src[r][c]);  f = gain * sigmoid or identity
    const float* src;
    float* dst_f;      // may be null, may alias src (identity ops)
    __nv_bfloat16* dst_t;   // may be null; [C][ld_t], columns r in [R, ld_t) zero-filled
    int R, C, ld_t;
    float gain;        // 0 = identity
    int phase;         // 0 = before the iterations (gates), 1 = after (H_res)
};

struct Plan {
    const JobDev* jobs;
    const Slab* slabs;
    const int* cta_slab;      // [grid + 1]
    const TransOp* ops;
    float* vec_u;             // current u per job row
    float* vec_v;             // current v per job column (ping-pong of 2)
    float* partials;          // [2][part_total] column partials per slab (ping-pong by pass parity)
    float* row_partials;      // [2][nslabs] sum of row sums (convergence history)
    float* kbuf;              // backward: K = softmax(raw) * m per job (mat_off), [sum D*D]
    float* bwd_hist;          // backward: per job [2 * iters][D]: tbar^k | sbar^k
    int njobs, nslabs, nops, vec_total, part_total;
    int iters;
    float eps;
    int max_d;
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

// 32 x 32 tiles through shared memory: coalesced reads of src rows, coalesced writes of dst_t rows
__device__ void run_trans_ops(const Plan& pl, int phase, float (*tile)[33]) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    for (int o = 0; o < pl.nops; ++o) {
        const TransOp op = pl.ops[o];
        if (op.phase != phase) continue;
        const int rt = (max(op.R, op.ld_t) + 31) / 32, ct = (op.C + 31) / 32;
        for (int t = blockIdx.x; t < rt * ct; t += gridDim.x) {
            const int r0 = (t / ct) * 32, c0 = (t % ct) * 32;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = r0 + ty + 8 * k, c = c0 + tx;
                float v = 0.f;
                if (r < op.R && c < op.C) {
                    v = op.src[(size_t)r * op.C + c];
                    if (op.gain != 0.f) v = op.gain * sigmoid_ref(v);
                    if (op.dst_f != nullptr && (op.gain != 0.f || op.dst_f != op.src)) op.dst_f[(size_t)r * op.C + c] = v;
                }
                tile[ty + 8 * k][tx] = v;
            }
            __syncthreads();
            if (op.dst_t != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = c0 + ty + 8 * k, r = r0 + tx;
                    if (c < op.C && r < op.ld_t) op.dst_t[(size_t)c * op.ld_t + r] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
                }
            }
        }
    }
}

// v of a job for the coming pass, from the previous pass's slab partials (ping-pong buffers: a fast CTA may already
// be writing the next pass's partials while a slow one still reads these):
//   t_j = sum_slabs partial;  v_j <- v_prev_j / (v_prev_j * t_j + eps)
// Every CTA that owns a slab of the job computes the same values in the same order; the job's first slab ("owner")
// also records them, and the reference's convergence history |mean(row_sum) - 1| (:76-77) of the previous pass.
__device__ void load_v_forward(const Plan& pl, const JobDev& jd, int iter, float* sm_v, bool owner) {
    const int D = jd.j.D;
    const float* vprev = pl.vec_v + (size_t)((iter & 1) ^ 1) * pl.vec_total + jd.vec_off;
    float* vcur = pl.vec_v + (size_t)(iter & 1) * pl.vec_total + jd.vec_off;
    const float* parts = pl.partials + (size_t)((iter - 1) & 1) * pl.part_total;
    for (int j = threadIdx.x; j < D; j += kThreads) {
        float v = 1.0f;
        if (iter > 0) {
            // (a job's slabs own consecutive D-float pieces of the partial buffer: direct addresses, so the loads of the sum are
            //  independent -- through the slab table every term was two dependent L2 round trips, 50 in a row for a 1792 matrix)
            const float* pj = parts + pl.slabs[jd.first_slab].part_off + j;
            float t = 0.f;
#pragma unroll 8
            for (int s = 0; s < jd.nslabs; ++s) t += pj[(size_t)s * D];
            const float vp = iter > 1 ? vprev[j] : 1.0f;
            v = __fdiv_rn(vp, vp * t + pl.eps);
            if (owner) {
                vcur[j] = v;
                if (jd.j.uv_history != nullptr) jd.j.uv_history[(size_t)(2 * iter + 1) * D + j] = v;
            }
        } else if (owner && jd.j.uv_history != nullptr) {
            jd.j.uv_history[(size_t)D + j] = 1.0f;
            jd.j.uv_history[j] = 1.0f;
        }
        sm_v[j] = v;
    }
    if (owner && iter > 0 && threadIdx.x == 0 && jd.j.convergence != nullptr) {
        const float* rp = pl.row_partials + (size_t)((iter - 1) & 1) * pl.nslabs;
        float t = 0.f;
        for (int s = 0; s < jd.nslabs; ++s) t += rp[jd.first_slab + s];
        jd.j.convergence[iter - 1] = fabsf(t / (float)D - 1.0f);
    }
    __syncthreads();
}

template <int NPL>
__device__ void forward_pass_slab(const Plan& pl, const Slab& sl, int slab_index, const JobDev& jd, int iter,
                                  const float* sm_v, float* sm_part, float* sm_rs) {
    const int D = jd.j.D, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* K = jd.j.h_res;
    float* u = pl.vec_u + jd.vec_off;
    float cp[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) cp[k] = 0.f;
    float rs_acc = 0.f;
    for (int r = warp; r < sl.nrows; r += kWarps) {
        const int i = sl.row0 + r;
        const float* row = K + (size_t)i * D;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int j = lane + 32 * k;
            if (j < D) s = fmaf(row[j], sm_v[j], s);
        }
        s = wsum(s);
        const float up = iter > 0 ? u[i] : 1.0f;
        rs_acc += up * s;                                          // the reference's row sum of P before this row step
        const float un = __fdiv_rn(up, up * s + pl.eps);
        if (lane == 0) {
            u[i] = un;
            if (jd.j.uv_history != nullptr) jd.j.uv_history[(size_t)(2 * (iter + 1)) * D + i] = un;
        }
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int j = lane + 32 * k;
            if (j < D) cp[k] = fmaf(row[j], un, cp[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
        const int j = lane + 32 * k;
        if (j < D) sm_part[warp * D + j] = cp[k];
    }
    if (lane == 0) sm_rs[warp] = rs_acc;
    __syncthreads();
    float* out = pl.partials + (size_t)(iter & 1) * pl.part_total + sl.part_off;
    for (int j = threadIdx.x; j < D; j += kThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += sm_part[w * D + j];
        out[j] = t;
    }
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kWarps; ++w) t += sm_rs[w];
        pl.row_partials[(size_t)(iter & 1) * pl.nslabs + slab_index] = t;
    }
    __syncthreads();
}

__device__ void forward_pass_dispatch(const Plan& pl, const Slab& sl, int slab_index, const JobDev& jd, int iter,
                                      const float* sm_v, float* sm_part, float* sm_rs) {
    const int npl = (jd.j.D + 31) / 32;
    if (npl <= 1) forward_pass_slab<1>(pl, sl, slab_index, jd, iter, sm_v, sm_part, sm_rs);
    else if (npl <= 2) forward_pass_slab<2>(pl, sl, slab_index, jd, iter, sm_v, sm_part, sm_rs);
    else if (npl <= 4) forward_pass_slab<4>(pl, sl, slab_index, jd, iter, sm_v, sm_part, sm_rs);
    else if (npl <= 8) forward_pass_slab<8>(pl, sl, slab_index, jd, iter, sm_v, sm_part, sm_rs);
    else if (npl <= 16) forward_pass_slab<16>(pl, sl, slab_index, jd, iter, sm_v, sm_part, sm_rs);
    else if (npl <= 32) forward_pass_slab<32>(pl, sl, slab_index, jd, iter, sm_v, sm_part, sm_rs);
    else if (npl <= 64) forward_pass_slab<64>(pl, sl, slab_index, jd, iter, sm_v, sm_part, sm_rs);
    else forward_pass_slab<128>(pl, sl, slab_index, jd, iter, sm_v, sm_part, sm_rs);
}

__global__ void __launch_bounds__(kThreads, 1) static_coeffs_fwd_kernel(const Plan pl) {
    extern __shared__ float sm[];
    float* sm_v = sm;                                  // [max_d]
    float* sm_part = sm + pl.max_d;                    // [kWarps][max_d]
    __shared__ float sm_rs[kWarps];
    __shared__ float tile[32][33];
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s_beg = pl.cta_slab[blockIdx.x], s_end = pl.cta_slab[blockIdx.x + 1];

    // ---- phase 0: gates (+ transposed bf16 copies) and K = softmax(raw, -1) * m  (:56-57) into the H_res buffer
    run_trans_ops(pl, 0, tile);
    for (int s = s_beg; s < s_end; ++s) {
        const Slab sl = pl.slabs[s];
        const JobDev& jd = pl.jobs[sl.job];
        const int D = jd.j.D;
        for (int r = warp; r < sl.nrows; r += kWarps) {
            const float* src = jd.j.h_res_raw + (size_t)(sl.row0 + r) * D;
            float* dst = jd.j.h_res + (size_t)(sl.row0 + r) * D;
            float mx = -INFINITY;
            for (int j = lane; j < D; j += 32) mx = fmaxf(mx, src[j]);
            mx = wmax(mx);
            float sum = 0.f;
            for (int j = lane; j < D; j += 32) sum += expf(src[j] - mx);
            sum = wsum(sum);
            for (int j = lane; j < D; j += 32) dst[j] = __fdiv_rn(expf(src[j] - mx), sum) * (float)D;
        }
    }
    grid.sync();

    // ---- the iterations (:64-77): one pass over K and one grid barrier each
    for (int it = 0; it < pl.iters; ++it) {
        for (int s = s_beg; s < s_end; ++s) {
            const Slab& sl = pl.slabs[s];
            const JobDev& jd = pl.jobs[sl.job];
            load_v_forward(pl, jd, it, sm_v, s == jd.first_slab);
            forward_pass_dispatch(pl, sl, s, jd, it, sm_v, sm_part, sm_rs);
        }
        grid.sync();
    }

    // ---- last column step, convergence history, P = diag(u) K diag(v) in place
    for (int s = s_beg; s < s_end; ++s) {
        const Slab sl = pl.slabs[s];
        const JobDev& jd = pl.jobs[sl.job];
        const int D = jd.j.D;
        load_v_forward(pl, jd, pl.iters, sm_v, s == jd.first_slab);
        const float* u = pl.vec_u + jd.vec_off;
        for (int r = warp; r < sl.nrows; r += kWarps) {
            const int i = sl.row0 + r;
            float* row = jd.j.h_res + (size_t)i * D;
            const float ui = pl.iters > 0 ? u[i] : 1.0f;
            for (int j = lane; j < D; j += 32) row[j] = ui * row[j] * sm_v[j];
        }
        __syncthreads();
    }
    grid.sync();
    run_trans_ops(pl, 1, tile);
}

}  // namespace
}  // namespace hvs

// =============================================================================================== backward
// Reverse sweep through the scaling iterations (eps = 1e-8 against sums of ~1 is below fp32 resolution, so the
// iteration is differentiated as u_k = 1 / (K v_{k-1}), v_k = 1 / (K^T u_k), P = diag(u_n) K diag(v_n)):
//   Kbar  = G o (u_n v_n^T);  ubar_i = sum_j G_ij K_ij v_n,j;  vbar_j = sum_i G_ij K_ij u_n,i
//   for k = n..1:  tbar = -vbar o v_k^2;  ubar += K tbar;  Kbar += u_k tbar^T
//                  sbar = -ubar o u_k^2;  vbar  = K^T sbar; Kbar += sbar v_{k-1}^T;  ubar = 0
//   d raw_ij = K_ij (Kbar_ij - sum_l K_il Kbar_il / m)            (softmax * m backward)
// Kbar is never materialised: the 2n vectors tbar^k, sbar^k are kept and the rank-2n update is applied in the final
// pass.  Same slab decomposition, same one-pass-one-barrier structure as the forward.
namespace hvs {
namespace {

__device__ void load_t_backward(const Plan& pl, const JobDev& jd, int pass, float* sm_v, bool owner) {
    const int D = jd.j.D;
    const int k = pl.iters - pass + 1;                              // iteration being reversed (pass >= 1)
    const float* parts = pl.partials + (size_t)((pass - 1) & 1) * pl.part_total;
    const float* vk = jd.j.uv_history + (size_t)(2 * k + 1) * D;
    float* hist = pl.bwd_hist + (size_t)2 * pl.iters * jd.vec_off + (size_t)(2 * (k - 1)) * D;
    for (int j = threadIdx.x; j < D; j += kThreads) {
        const float* pj = parts + pl.slabs[jd.first_slab].part_off + j;
        float vb = 0.f;
#pragma unroll 8
        for (int s = 0; s < jd.nslabs; ++s) vb += pj[(size_t)s * D];
        const float v = vk[j];
        const float tb = -vb * v * v;
        if (owner) hist[j] = tb;
        sm_v[j] = tb;
    }
    __syncthreads();
}

template <int NPL>
__device__ void backward_pass_slab(const Plan& pl, const Slab& sl, const JobDev& jd, int pass, const float* sm_v,
                                   float* sm_part) {
    const int D = jd.j.D, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* K = pl.kbuf + jd.mat_off;
    const float* G = jd.g.d_h_res;
    const int n = pl.iters;
    float cp[NPL];
#pragma unroll
    for (int q = 0; q < NPL; ++q) cp[q] = 0.f;
    if (pass == 0) {
        const float* un = jd.j.uv_history + (size_t)(2 * n) * D;
        const float* vn = jd.j.uv_history + (size_t)(2 * n + 1) * D;
        float* ubar = pl.vec_u + jd.vec_off;
        for (int r = warp; r < sl.nrows; r += kWarps) {
            const int i = sl.row0 + r;
            const float* krow = K + (size_t)i * D;
            const float* grow = G + (size_t)i * D;
            const float ui = un[i];
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const int j = lane + 32 * q;
                if (j < D) {
                    const float gk = grow[j] * krow[j];
                    s = fmaf(gk, vn[j], s);
                    cp[q] = fmaf(gk, ui, cp[q]);
                }
            }
            s = wsum(s);
            if (lane == 0) ubar[i] = s;
        }
    } else {
        const int k = n - pass + 1;
        const float* uk = jd.j.uv_history + (size_t)(2 * k) * D;
        const float* ubar0 = pl.vec_u + jd.vec_off;
        float* shist = pl.bwd_hist + (size_t)2 * n * jd.vec_off + (size_t)(2 * (k - 1) + 1) * D;
        // Narrow matrices have hundreds of rows per warp and one or two elements per lane in each: a row at a time is a chain of
        // L2 round trips (a pass over the model's D = 32 .. 128 layers took 60-180 us, 20 passes per step).  RU rows in flight.
        constexpr int RU = NPL <= 2 ? 8 : (NPL <= 8 ? 4 : (NPL <= 16 ? 2 : 1));
        if constexpr (RU == 1) {                              // wide rows: many loads per lane already, the row streams twice
            for (int r = warp; r < sl.nrows; r += kWarps) {
                const int i = sl.row0 + r;
                const float* krow = K + (size_t)i * D;
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int j = lane + 32 * q;
                    if (j < D) a = fmaf(krow[j], sm_v[j], a);
                }
                a = wsum(a);
                if (pass == 1) a += ubar0[i];
                const float u = uk[i];
                const float sb = -a * u * u;
                if (lane == 0) shist[i] = sb;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int j = lane + 32 * q;
                    if (j < D) cp[q] = fmaf(krow[j], sb, cp[q]);
                }
            }
        } else
        for (int r0 = warp * RU; r0 < sl.nrows; r0 += kWarps * RU) {
            float kr[RU][NPL], a[RU];
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const bool ok = r0 + u < sl.nrows;
                const float* krow = K + (size_t)(sl.row0 + (ok ? r0 + u : r0)) * D;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int j = lane + 32 * q;
                    kr[u][q] = (ok && j < D) ? krow[j] : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                a[u] = 0.f;
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    const int j = lane + 32 * q;
                    if (j < D) a[u] = fmaf(kr[u][q], sm_v[j], a[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < RU; ++u) a[u] = wsum(a[u]);
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                if (r0 + u < sl.nrows) {
                    const int i = sl.row0 + r0 + u;
                    float av = a[u];
                    if (pass == 1) av += ubar0[i];
                    const float uu = uk[i];
                    const float sb = -av * uu * uu;
                    if (lane == 0) shist[i] = sb;
#pragma unroll
                    for (int q = 0; q < NPL; ++q) cp[q] = fmaf(kr[u][q], sb, cp[q]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
        const int j = lane + 32 * q;
        if (j < D) sm_part[warp * D + j] = cp[q];
    }
    __syncthreads();
    float* out = pl.partials + (size_t)(pass & 1) * pl.part_total + sl.part_off;
    for (int j = threadIdx.x; j < D; j += kThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += sm_part[w * D + j];
        out[j] = t;
    }
    __syncthreads();
}

__device__ void backward_pass_dispatch(const Plan& pl, const Slab& sl, const JobDev& jd, int pass, const float* sm_v,
                                       float* sm_part) {
    const int npl = (jd.j.D + 31) / 32;
    if (npl <= 1) backward_pass_slab<1>(pl, sl, jd, pass, sm_v, sm_part);
    else if (npl <= 2) backward_pass_slab<2>(pl, sl, jd, pass, sm_v, sm_part);
    else if (npl <= 4) backward_pass_slab<4>(pl, sl, jd, pass, sm_v, sm_part);
    else if (npl <= 8) backward_pass_slab<8>(pl, sl, jd, pass, sm_v, sm_part);
    else if (npl <= 16) backward_pass_slab<16>(pl, sl, jd, pass, sm_v, sm_part);
    else if (npl <= 32) backward_pass_slab<32>(pl, sl, jd, pass, sm_v, sm_part);
    else if (npl <= 64) backward_pass_slab<64>(pl, sl, jd, pass, sm_v, sm_part);
    else backward_pass_slab<128>(pl, sl, jd, pass, sm_v, sm_part);
}

constexpr int kRowBlock = 4, kRowFactorIters = 32;

__global__ void __launch_bounds__(kThreads, 1) static_coeffs_bwd_kernel(const Plan pl) {
    extern __shared__ float sm[];
    __shared__ float s_rowf[kWarps][kRowBlock][2 * kRowFactorIters];
    float* sm_v = sm;
    float* sm_part = sm + pl.max_d;
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s_beg = pl.cta_slab[blockIdx.x], s_end = pl.cta_slab[blockIdx.x + 1];
    const int n = pl.iters;

    // ---- gate gradients: d raw = dH * gain * s (1 - s)
    for (int jb = 0; jb < pl.njobs; ++jb) {
        const JobDev& jd = pl.jobs[jb];
        const int64_t ne = (int64_t)jd.j.D * jd.j.H;
        if (jd.g.d_h_pre != nullptr && jd.g.d_h_pre_raw != nullptr)
            for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < ne; e += (int64_t)gridDim.x * kThreads) {
                const float s = sigmoid_ref(jd.j.h_pre_raw[e]);
                jd.g.d_h_pre_raw[e] = jd.g.d_h_pre[e] * s * (1.0f - s);
            }
        if (jd.g.d_h_post != nullptr && jd.g.d_h_post_raw != nullptr)
            for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < ne; e += (int64_t)gridDim.x * kThreads) {
                const float s = sigmoid_ref(jd.j.h_post_raw[e]);
                jd.g.d_h_post_raw[e] = jd.g.d_h_post[e] * 2.0f * s * (1.0f - s);
            }
    }
    // ---- K = softmax(raw) * m into the scratch matrix, then pass 0
    for (int s = s_beg; s < s_end; ++s) {
        const Slab sl = pl.slabs[s];
        const JobDev& jd = pl.jobs[sl.job];
        if (jd.g.d_h_res == nullptr || jd.g.d_h_res_raw == nullptr) continue;
        const int D = jd.j.D;
        for (int r = warp; r < sl.nrows; r += kWarps) {
            const float* src = jd.j.h_res_raw + (size_t)(sl.row0 + r) * D;
            float* dst = pl.kbuf + jd.mat_off + (size_t)(sl.row0 + r) * D;
            float mx = -INFINITY;
            for (int j = lane; j < D; j += 32) mx = fmaxf(mx, src[j]);
            mx = wmax(mx);
            float sum = 0.f;
            for (int j = lane; j < D; j += 32) sum += expf(src[j] - mx);
            sum = wsum(sum);
            for (int j = lane; j < D; j += 32) dst[j] = __fdiv_rn(expf(src[j] - mx), sum) * (float)D;
        }
        __syncthreads();
        backward_pass_dispatch(pl, sl, jd, 0, sm_v, sm_part);
    }
    grid.sync();
    for (int pass = 1; pass <= n; ++pass) {
        for (int s = s_beg; s < s_end; ++s) {
            const Slab& sl = pl.slabs[s];
            const JobDev& jd = pl.jobs[sl.job];
            if (jd.g.d_h_res == nullptr || jd.g.d_h_res_raw == nullptr) continue;
            load_t_backward(pl, jd, pass, sm_v, s == jd.first_slab);
            backward_pass_dispatch(pl, sl, jd, pass, sm_v, sm_part);
        }
        grid.sync();
    }
    // ---- final pass: Kbar row by row, softmax backward
    for (int s = s_beg; s < s_end; ++s) {
        const Slab sl = pl.slabs[s];
        const JobDev& jd = pl.jobs[sl.job];
        if (jd.g.d_h_res == nullptr || jd.g.d_h_res_raw == nullptr) continue;
        const int D = jd.j.D;
        const float* K = pl.kbuf + jd.mat_off;
        const float* hist = jd.j.uv_history;
        const float* bh = pl.bwd_hist + (size_t)2 * n * jd.vec_off;
        if (n <= kRowFactorIters) {
            // four rows per warp at a time: the 2n column vectors (bh / hist rows at column j) are fetched once for the four, the
            // rows' own 2n factors wait in shared memory.  (One row at a time fetched four L2 values per element and iteration:
            // 740 M loads for the model's 76 layers, most of the kernel's time.)  Kbar goes through the output row itself.
            for (int r = warp * kRowBlock; r < sl.nrows; r += kWarps * kRowBlock) {
                const int nr = sl.nrows - r < kRowBlock ? sl.nrows - r : kRowBlock;
                __syncwarp();
                for (int q = lane; q < kRowBlock * 2 * n; q += 32) {
                    const int rr = q / (2 * n), kk = q - rr * 2 * n;
                    const int i = sl.row0 + r + (rr < nr ? rr : nr - 1);
                    s_rowf[warp][rr][kk] = kk < n ? hist[(size_t)(2 * (kk + 1)) * D + i] : bh[(size_t)(2 * (kk - n) + 1) * D + i];
                }
                __syncwarp();
                float un[kRowBlock], dot[kRowBlock];
                const float* grow[kRowBlock];
                const float* krow[kRowBlock];
                float* orow[kRowBlock];
#pragma unroll
                for (int rr = 0; rr < kRowBlock; ++rr) {
                    const int i = sl.row0 + r + (rr < nr ? rr : nr - 1);
                    un[rr] = hist[(size_t)(2 * n) * D + i];
                    dot[rr] = 0.f;
                    grow[rr] = jd.g.d_h_res + (size_t)i * D;
                    krow[rr] = K + (size_t)i * D;
                    orow[rr] = jd.g.d_h_res_raw + (size_t)i * D;
                }
                for (int j = lane; j < D; j += 32) {
                    const float vfin = hist[(size_t)(2 * n + 1) * D + j];
                    float kb[kRowBlock];
#pragma unroll
                    for (int rr = 0; rr < kRowBlock; ++rr) kb[rr] = grow[rr][j] * un[rr] * vfin;
                    for (int k = 1; k <= n; ++k) {
                        const float bcol = bh[(size_t)(2 * (k - 1)) * D + j], hcol = hist[(size_t)(2 * (k - 1) + 1) * D + j];
#pragma unroll
                        for (int rr = 0; rr < kRowBlock; ++rr) kb[rr] += s_rowf[warp][rr][k - 1] * bcol + s_rowf[warp][rr][n + k - 1] * hcol;
                    }
#pragma unroll
                    for (int rr = 0; rr < kRowBlock; ++rr)
                        if (rr < nr) {
                            orow[rr][j] = kb[rr];
                            dot[rr] = fmaf(kb[rr], krow[rr][j], dot[rr]);
                        }
                }
#pragma unroll
                for (int rr = 0; rr < kRowBlock; ++rr) dot[rr] = wsum(dot[rr]) / (float)D;
                for (int j = lane; j < D; j += 32) {
#pragma unroll
                    for (int rr = 0; rr < kRowBlock; ++rr)
                        if (rr < nr) orow[rr][j] = krow[rr][j] * (orow[rr][j] - dot[rr]);      // (this lane wrote orow[rr][j] itself)
                }
            }
            __syncthreads();
            continue;
        }
        float* rowbuf = sm_part + warp * D;
        for (int r = warp; r < sl.nrows; r += kWarps) {
            const int i = sl.row0 + r;
            const float* krow = K + (size_t)i * D;
            const float* grow = jd.g.d_h_res + (size_t)i * D;
            const float un = hist[(size_t)(2 * n) * D + i];
            float dot = 0.f;
            for (int j = lane; j < D; j += 32) {
                float kb = grow[j] * un * hist[(size_t)(2 * n + 1) * D + j];
                for (int k = 1; k <= n; ++k)
                    kb += hist[(size_t)(2 * k) * D + i] * bh[(size_t)(2 * (k - 1)) * D + j] +
                          bh[(size_t)(2 * (k - 1) + 1) * D + i] * hist[(size_t)(2 * (k - 1) + 1) * D + j];
                rowbuf[j] = kb;
                dot = fmaf(kb, krow[j], dot);
            }
            dot = wsum(dot) / (float)D;
            float* out = jd.g.d_h_res_raw + (size_t)i * D;
            for (int j = lane; j < D; j += 32) out[j] = krow[j] * (rowbuf[j] - dot);
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------- host plan
struct HostPlan {
    std::vector<JobDev> jobs;
    std::vector<Slab> slabs;
    std::vector<int> cta_slab;
    std::vector<TransOp> ops;
    int vec_total = 0, part_total = 0, max_d = 0;
    size_t mat_total = 0;
    // workspace layout (byte offsets)
    size_t off_jobs = 0, off_slabs = 0, off_cta = 0, off_ops = 0, off_u = 0, off_v = 0, off_part = 0, off_rows = 0,
           off_kbuf = 0, off_bh = 0, total = 0, table_bytes = 0;
};

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

int build_plan(HostPlan& hp, const hvs_coeff_job* jobs, const hvs_coeff_grad* grads, int num_jobs, int grid, int iters,
               bool backward) {
    double total_elems = 0;
    for (int b = 0; b < num_jobs; ++b) {
        const hvs_coeff_job& j = jobs[b];
        if (j.D <= 0 || j.H <= 0 || j.D > kMaxD) return j.D > kMaxD ? HVS_ERR_UNSUPPORTED : HVS_ERR_BAD_ARG;
        if (!j.h_res_raw || !j.h_res) return HVS_ERR_BAD_ARG;
        if (backward && !j.uv_history) return HVS_ERR_BAD_ARG;
        total_elems += (double)j.D * j.D;
        if (j.D > hp.max_d) hp.max_d = j.D;
    }
    const double target = total_elems / grid > 2048.0 ? total_elems / grid : 2048.0;
    for (int b = 0; b < num_jobs; ++b) {
        const hvs_coeff_job& j = jobs[b];
        JobDev jd{};
        jd.j = j;
        if (grads) jd.g = grads[b];
        jd.first_slab = (int)hp.slabs.size();
        int rows = (int)(target / j.D);
        if (rows < 1) rows = 1;
        if (rows > j.D) rows = j.D;
        const int ns = (j.D + rows - 1) / rows;
        rows = (j.D + ns - 1) / ns;                               // even slabs
        for (int r0 = 0; r0 < j.D; r0 += rows) {
            Slab s{b, r0, r0 + rows <= j.D ? rows : j.D - r0, hp.part_total};
            hp.part_total += j.D;
            hp.slabs.push_back(s);
        }
        jd.nslabs = (int)hp.slabs.size() - jd.first_slab;
        jd.vec_off = hp.vec_total;
        hp.vec_total += j.D;
        jd.mat_off = (int)hp.mat_total;
        hp.mat_total += (size_t)j.D * j.D;
        hp.jobs.push_back(jd);
        if (!backward) {
            if (j.h_pre_raw && (j.h_pre || j.h_pre_t))
                hp.ops.push_back(TransOp{j.h_pre_raw, j.h_pre, (__nv_bfloat16*)j.h_pre_t, j.D, j.H, j.Dp, 1.0f, 0});
            if (j.h_post_raw && (j.h_post || j.h_post_t))
                hp.ops.push_back(TransOp{j.h_post_raw, j.h_post, (__nv_bfloat16*)j.h_post_t, j.H, j.D, j.H, 2.0f, 0});
            if (j.h_res_t) hp.ops.push_back(TransOp{j.h_res, j.h_res, (__nv_bfloat16*)j.h_res_t, j.D, j.D, j.Dp, 0.0f, 1});
        }
    }
    if (hp.mat_total >= ((size_t)1 << 31)) return HVS_ERR_UNSUPPORTED;
    // contiguous slab ranges of about equal element count per CTA
    hp.cta_slab.assign(grid + 1, (int)hp.slabs.size());
    hp.cta_slab[0] = 0;
    double acc = 0;
    int cta = 1;
    for (size_t s = 0; s < hp.slabs.size() && cta < grid; ++s) {
        acc += (double)hp.slabs[s].nrows * jobs[hp.slabs[s].job].D;
        while (cta < grid && acc >= total_elems * cta / grid) hp.cta_slab[cta++] = (int)s + 1;
    }
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
    hp.off_jobs = take(hp.jobs.size() * sizeof(JobDev));
    hp.off_slabs = take(hp.slabs.size() * sizeof(Slab));
    hp.off_cta = take(hp.cta_slab.size() * sizeof(int));
    hp.off_ops = take((hp.ops.size() + 1) * sizeof(TransOp));
    hp.table_bytes = off;
    hp.off_u = take((size_t)hp.vec_total * 4);
    hp.off_v = take((size_t)hp.vec_total * 4 * 2);
    hp.off_part = take((size_t)hp.part_total * 4 * 2);
    hp.off_rows = take(hp.slabs.size() * 4 * 2);
    if (backward) {
        hp.off_kbuf = take(hp.mat_total * 4);
        hp.off_bh = take((size_t)2 * (iters > 0 ? iters : 1) * hp.vec_total * 4);
    }
    hp.total = off;
    return HVS_OK;
}

int coop_grid(const void* kernel, int smem) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess || per_sm < 1) return 0;
    return sm_count();                                              // one CTA per SM (per_sm >= 1 guarantees co-residency)
}

int run_plan(const hvs_coeff_job* jobs, const hvs_coeff_grad* grads, int num_jobs, int iters, float eps, void* workspace,
             size_t workspace_bytes, cudaStream_t stream, bool backward, bool size_only, size_t* size_out) {
    if (num_jobs < 0 || iters < 0 || (num_jobs > 0 && !jobs)) return HVS_ERR_BAD_ARG;
    if (num_jobs == 0) { if (size_out) *size_out = 256; return HVS_OK; }
    // as many CTAs as the work can feed (>= 2048 matrix elements each): a single 256 x 256 layer (the per-layer backward of a
    // training step) runs on 32 CTAs, whose grid barrier is cheaper than one across all 148 SMs.  (8192 per CTA was tried:
    // the barriers got cheaper still but the first / last pass -- exp, the rank-2n update -- serialised: 233 -> 390 us.)
    double elems = 0;
    for (int b = 0; b < num_jobs; ++b) elems += (double)jobs[b].D * jobs[b].D;
    int grid = (int)(elems / 2048.0);
    if (grid < 1) grid = 1;
    if (grid > sm_count()) grid = sm_count();
    HostPlan hp;
    int rc = build_plan(hp, jobs, grads, num_jobs, grid, iters, backward);
    if (rc) return rc;
    if (size_out) *size_out = hp.total;
    if (size_only) return HVS_OK;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return HVS_ERR_ALIGNMENT;
    if (workspace_bytes < hp.total) return HVS_ERR_WORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    std::vector<uint8_t> tables(hp.table_bytes, 0);
    memcpy(tables.data() + hp.off_jobs, hp.jobs.data(), hp.jobs.size() * sizeof(JobDev));
    memcpy(tables.data() + hp.off_slabs, hp.slabs.data(), hp.slabs.size() * sizeof(Slab));
    memcpy(tables.data() + hp.off_cta, hp.cta_slab.data(), hp.cta_slab.size() * sizeof(int));
    if (!hp.ops.empty()) memcpy(tables.data() + hp.off_ops, hp.ops.data(), hp.ops.size() * sizeof(TransOp));
    {
        const int rcu = upload_table(ws, std::move(tables), stream);
        if (rcu) return rcu;
    }
    Plan pl{};
    pl.jobs = reinterpret_cast<const JobDev*>(ws + hp.off_jobs);
    pl.slabs = reinterpret_cast<const Slab*>(ws + hp.off_slabs);
    pl.cta_slab = reinterpret_cast<const int*>(ws + hp.off_cta);
    pl.ops = reinterpret_cast<const TransOp*>(ws + hp.off_ops);
    pl.vec_u = reinterpret_cast<float*>(ws + hp.off_u);
    pl.vec_v = reinterpret_cast<float*>(ws + hp.off_v);
    pl.partials = reinterpret_cast<float*>(ws + hp.off_part);
    pl.row_partials = reinterpret_cast<float*>(ws + hp.off_rows);
    pl.kbuf = backward ? reinterpret_cast<float*>(ws + hp.off_kbuf) : nullptr;
    pl.bwd_hist = backward ? reinterpret_cast<float*>(ws + hp.off_bh) : nullptr;
    pl.njobs = num_jobs; pl.nslabs = (int)hp.slabs.size(); pl.nops = (int)hp.ops.size();
    pl.vec_total = hp.vec_total; pl.part_total = hp.part_total;
    pl.iters = iters; pl.eps = eps; pl.max_d = hp.max_d;
    const int smem = (1 + kWarps) * hp.max_d * (int)sizeof(float);
    void* args[] = {&pl};
    if (backward) {
        HVS_SET_MAX_SMEM(static_coeffs_bwd_kernel, (1 + kWarps) * kMaxD * 4);
        if (coop_grid((const void*)static_coeffs_bwd_kernel, smem) == 0) return HVS_ERR_UNSUPPORTED;
        timer_begin(7, stream);
        HVS_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)static_coeffs_bwd_kernel, dim3(grid), dim3(kThreads), args, smem, stream));
        timer_end(7, stream);
    } else {
        HVS_SET_MAX_SMEM(static_coeffs_fwd_kernel, (1 + kWarps) * kMaxD * 4);
        if (coop_grid((const void*)static_coeffs_fwd_kernel, smem) == 0) return HVS_ERR_UNSUPPORTED;
        timer_begin(7, stream);
        HVS_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)static_coeffs_fwd_kernel, dim3(grid), dim3(kThreads), args, smem, stream));
        timer_end(7, stream);
    }
    count_launch();
    return launch_status();
}

}  // namespace
}  // namespace hvs

extern "C" size_t hvs_mhc_static_coeffs_workspace(const hvs_coeff_job* jobs_host, int num_jobs, int iters, int backward) {
    size_t n = 0;
    if (hvs::run_plan(jobs_host, nullptr, num_jobs, iters, 0.f, nullptr, 0, nullptr, backward != 0, true, &n) != HVS_OK) return 0;
    return n;
}

extern "C" int hvs_mhc_static_coeffs(const hvs_coeff_job* jobs_host, int num_jobs, int iters, float eps, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    return hvs::run_plan(jobs_host, nullptr, num_jobs, iters, eps, workspace, workspace_bytes, (cudaStream_t)stream, false, false, nullptr);
}

extern "C" int hvs_mhc_static_coeffs_bwd(const hvs_coeff_job* jobs_host, const hvs_coeff_grad* grads_host, int num_jobs,
                                         int iters, float eps, void* workspace, size_t workspace_bytes, void* stream) {
    if (num_jobs > 0 && !grads_host) return HVS_ERR_BAD_ARG;
    return hvs::run_plan(jobs_host, grads_host, num_jobs, iters, eps, workspace, workspace_bytes, (cudaStream_t)stream, true, false, nullptr);
}
