// Stand-alone Sinkhorn-Knopp projection and the static (parameter-only) coefficient path of the
// reference-literal module:
//   SinkhornKnoppProjection.forward      src/models/manifold_layers.py:32-93
//   ManifoldHyperConnection.constrained_matrices                      :205-221
// Two shapes of work: many tiny blocks (a warp per matrix, columns in lanes, rows in registers) and
// one large D x D matrix per layer (a CTA per matrix, iterating in place on the L2-resident output).
// Divisions are IEEE (the reference divides); the reductions that produce `out` have a fixed order, so `out` is
// bitwise reproducible run to run (the small-block convergence history, a diagnostic, sums with fp32 atomics).
#include <math.h>

#include "common.cuh"

namespace hvs {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------- small blocks
// One warp per matrix, lane j holds column j (m <= 32), rows unrolled up to NMAX.
template <int NMAX>
__global__ void __launch_bounds__(128)
sinkhorn_small_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t batch, int n, int m,
                      int iters, float eps, float inv_tau, float* __restrict__ hist_acc) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool act = lane < m;
    for (int64_t b = wid; b < batch; b += nw) {
        const float* src = in + b * n * m;
        float p[NMAX];
#pragma unroll
        for (int i = 0; i < NMAX; ++i) {
            float v = (i < n && act) ? src[i * m + lane] * inv_tau : -INFINITY;
            const float mx = warp_max(v);
            const float e = (i < n && act) ? expf(v - mx) : 0.f;
            const float s = warp_sum(e);
            p[i] = (i < n && act) ? __fdiv_rn(e, s) * (float)m : 0.f;
        }
        for (int it = 0; it < iters; ++it) {
            float rs_total = 0.f;
#pragma unroll
            for (int i = 0; i < NMAX; ++i) {
                if (i < n) {
                    const float rs = warp_sum(p[i]);
                    rs_total += rs;
                    p[i] = __fdiv_rn(p[i], rs + eps);
                }
            }
            float cs = 0.f;
#pragma unroll
            for (int i = 0; i < NMAX; ++i) cs += p[i];
            const float cd = cs + eps;
#pragma unroll
            for (int i = 0; i < NMAX; ++i) p[i] = __fdiv_rn(p[i], cd);
            if (hist_acc != nullptr && lane == 0) atomicAdd(hist_acc + it, rs_total);
        }
        float* dst = out + b * n * m;
#pragma unroll
        for (int i = 0; i < NMAX; ++i)
            if (i < n && act) dst[i * m + lane] = p[i];
    }
}

__global__ void hist_finalize_kernel(float* hist, int iters, float inv_rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < iters) hist[i] = fabsf(hist[i] * inv_rows - 1.0f);
}

// ---------------------------------------------------------------------------- one large matrix per CTA
constexpr int kLargeThreads = 1024;

struct LargeJob {
    const float* in;
    float* out;
    float* history;   // [iters] or null
    int n, m;
};

__device__ void sinkhorn_large_body(const LargeJob& job, int iters, float eps, float inv_tau, float* sm_red) {
    const int n = job.n, m = job.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kLargeThreads / 32;
    float* out = job.out;
    // softmax(M / tau, dim=-1) * m   (:56-57)
    for (int i = warp; i < n; i += kWarps) {
        const float* src = job.in + (size_t)i * m;
        float mx = -INFINITY;
        for (int j = lane; j < m; j += 32) mx = fmaxf(mx, src[j] * inv_tau);
        mx = warp_max(mx);
        float s = 0.f;
        for (int j = lane; j < m; j += 32) s += expf(src[j] * inv_tau - mx);
        s = warp_sum(s);
        for (int j = lane; j < m; j += 32) out[(size_t)i * m + j] = __fdiv_rn(expf(src[j] * inv_tau - mx), s) * (float)m;
    }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        // rows (:66-67)
        float hist_part = 0.f;
        for (int i = warp; i < n; i += kWarps) {
            float* row = out + (size_t)i * m;
            float s = 0.f;
            for (int j = lane; j < m; j += 32) s += row[j];
            s = warp_sum(s);
            hist_part += s;
            const float d = s + eps;
            for (int j = lane; j < m; j += 32) row[j] = __fdiv_rn(row[j], d);
        }
        if (job.history != nullptr) {
            if (lane == 0) sm_red[warp] = hist_part;
        }
        __syncthreads();
        if (job.history != nullptr && threadIdx.x == 0) {
            float tot = 0.f;
            for (int w2 = 0; w2 < kWarps; ++w2) tot += sm_red[w2];
            job.history[it] = fabsf(tot / (float)n - 1.0f);      // |mean(row_sum) - 1|  (:76-77)
        }
        // columns (:71-72): a thread walks one column, a warp reads 128 contiguous bytes per row
        for (int j = threadIdx.x; j < m; j += kLargeThreads) {
            float s = 0.f;
            for (int i = 0; i < n; ++i) s += out[(size_t)i * m + j];
            const float d = s + eps;
            for (int i = 0; i < n; ++i) out[(size_t)i * m + j] = __fdiv_rn(out[(size_t)i * m + j], d);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kLargeThreads)
sinkhorn_large_kernel(LargeJob job, int iters, float eps, float inv_tau) {
    __shared__ float sm_red[kLargeThreads / 32];
    sinkhorn_large_body(job, iters, eps, inv_tau, sm_red);
}

// sigmoid gates of constrained_matrices (:213, :216): out = gain * sigmoid(in)
__global__ void sigmoid_gate_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float gain) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = gain * __fdiv_rn(1.0f, 1.0f + expf(-in[i]));
}

}  // namespace
}  // namespace hvs

extern "C" int hvs_sinkhorn(const float* in, float* out, int64_t batch, int n, int m, int iters, float eps,
                            float tau, float* history, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!in || !out || batch < 0 || n <= 0 || m <= 0 || iters < 0 || !(tau > 0.f)) return HVS_ERR_BAD_ARG;
    if (batch == 0) return HVS_OK;
    const float inv_tau = 1.0f / tau;
    if (n <= 32 && m <= 32) {
        if (history != nullptr && iters > 0) HVS_CUDA_TRY(cudaMemsetAsync(history, 0, sizeof(float) * iters, stream));
        int64_t blocks = (batch + 3) / 4;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        const int nmax = n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32;
        switch (nmax) {
            case 4: sinkhorn_small_kernel<4><<<(int)blocks, 128, 0, stream>>>(in, out, batch, n, m, iters, eps, inv_tau, history); break;
            case 8: sinkhorn_small_kernel<8><<<(int)blocks, 128, 0, stream>>>(in, out, batch, n, m, iters, eps, inv_tau, history); break;
            case 16: sinkhorn_small_kernel<16><<<(int)blocks, 128, 0, stream>>>(in, out, batch, n, m, iters, eps, inv_tau, history); break;
            default: sinkhorn_small_kernel<32><<<(int)blocks, 128, 0, stream>>>(in, out, batch, n, m, iters, eps, inv_tau, history); break;
        }
        count_launch();
        if (history != nullptr && iters > 0) {
            hist_finalize_kernel<<<(iters + 63) / 64, 64, 0, stream>>>(history, iters, 1.0f / (float)(batch * n));
            count_launch();
        }
        return launch_status();
    }
    if (n > 4096 || m > 4096) return HVS_ERR_UNSUPPORTED;
    for (int64_t b = 0; b < batch; ++b) {
        if (b > 0 && history != nullptr) return HVS_ERR_UNSUPPORTED;   // history of a large batch: not defined here
        LargeJob job{in + b * (int64_t)n * m, out + b * (int64_t)n * m, history, n, m};
        sinkhorn_large_kernel<<<1, kLargeThreads, 0, stream>>>(job, iters, eps, inv_tau);
        count_launch();
    }
    return launch_status();
}

extern "C" int hvs_mhc_constrained_matrices(const float* h_pre_raw, const float* h_post_raw, const float* h_res_raw,
                                            float* h_pre, float* h_post, float* h_res, int D, int hidden, int iters,
                                            float eps, float* history, void* stream_) {
    using namespace hvs;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h_pre_raw || !h_post_raw || !h_res_raw || !h_pre || !h_post || !h_res || D <= 0 || hidden <= 0)
        return HVS_ERR_BAD_ARG;
    const int64_t ne = (int64_t)D * hidden;
    int blocks = (int)((ne + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    sigmoid_gate_kernel<<<blocks, 256, 0, stream>>>(h_pre_raw, h_pre, ne, 1.0f);
    sigmoid_gate_kernel<<<blocks, 256, 0, stream>>>(h_post_raw, h_post, ne, 2.0f);
    count_launch(2);
    return hvs_sinkhorn(h_res_raw, h_res, 1, D, D, iters, eps, 1.0f, history, stream_);
}
