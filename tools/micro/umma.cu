// Issue / completion cost of small tcgen05.mma (kind::f16, bf16, SS mode, cta_group::1) in the shapes the fused
// backward uses, one CTA, one issuing thread.  nvcc -gencode arch=compute_100a,code=sm_100a -I ../../hvs_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "umma_sm100.cuh"
using namespace hvs;

struct Cfg { int m, n, a_mn, b_mn; uint32_t a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout; int n_acc; uint32_t acc_stride; uint32_t lane_alt; int a_step, b_step; int hmma; };

__global__ void __launch_bounds__(160, 1) k_umma(Cfg c, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 0) {
        const uint32_t s0 = smem_u32(smem);
        const uint32_t idesc = umma_idesc_bf16(c.m, c.n, c.a_mn, c.b_mn);
        const uint64_t a0 = umma_smem_desc(s0, c.a_lbo, c.a_sbo, c.a_layout);
        const uint64_t b0 = umma_smem_desc(s0 + 100 * 1024, c.b_lbo, c.b_sbo, c.b_layout);
        long long t_issue = 0, t_done = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t d = tb + (uint32_t)((i & (c.n_acc - 1)) >> (c.lane_alt ? 1 : 0)) * c.acc_stride + ((c.lane_alt && (i & 1)) ? (16u << 16) : 0u);
                if (lane == 0) umma_bf16_ss(d, a0 + (uint64_t)(i * c.a_step), b0 + (uint64_t)(i * c.b_step), idesc, 1u);
            }
            const long long t1 = clock64();
            if (lane == 0) umma_commit(&bar);
            mbar_wait(&bar, (uint32_t)r & 1u);
            tc_fence_after();
            const long long t2 = clock64();
            t_issue += t1 - t0; t_done += t2 - t0;
        }
        if (lane == 0) { out[0] = t_issue / reps; out[1] = t_done / reps; }
    } else if (c.hmma) {
        // legacy mma.sync on the four other warps (one per SM sub-partition) while the tcgen05 MMAs run
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t a = 0x3c003c00u + lane, b = 0x3c003c00u;
        const long long t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < c.hmma; ++i) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a), "r"(a), "r"(b), "r"(b), "r"(b), "r"(a));
        }
        const long long t1 = clock64();
        if (lane == 0) out[2 + warp] = (t1 - t0) / (c.hmma * 16);
        if (d[0] == 123.f) out[7] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Named { const char* name; Cfg c; };
    const uint32_t SW = kUmmaLayoutSw128, NO = kUmmaLayoutNone;
    Named cfgs[] = {
        //                      m   n  amn bmn a_lbo a_sbo  a_lay b_lbo b_sbo b_lay nacc stride alt a_step b_step hmma
        {"GS  64x32 same D    ", {64, 32, 0, 0, 16, 8192, SW, 16, 8192, SW, 1, 32, 0, 2, 2, 0}},
        {"GS  64x32 2 acc     ", {64, 32, 0, 0, 16, 8192, SW, 16, 8192, SW, 2, 32, 0, 2, 2, 0}},
        {"GS  64x32 4 acc     ", {64, 32, 0, 0, 16, 8192, SW, 16, 8192, SW, 4, 32, 0, 2, 2, 0}},
        {"GS  64x32 2 acc lane", {64, 32, 0, 0, 16, 8192, SW, 16, 8192, SW, 2, 32, 1, 2, 2, 0}},
        {"GS  64x16 same D    ", {64, 16, 0, 0, 16, 8192, SW, 16, 8192, SW, 1, 32, 0, 2, 2, 0}},
        {"GS  64x8  same D    ", {64, 8, 0, 0, 16, 8192, SW, 16, 8192, SW, 1, 32, 0, 2, 2, 0}},
        {"GS  64x64 same D    ", {64, 64, 0, 0, 16, 8192, SW, 16, 8192, SW, 1, 64, 0, 2, 2, 0}},
        {"GS  64x128 same D   ", {64, 128, 0, 0, 16, 8192, SW, 16, 1024, SW, 1, 128, 0, 2, 2, 0}},
        {"GS 128x32 same D    ", {128, 32, 0, 0, 16, 8192, SW, 16, 8192, SW, 1, 32, 0, 2, 2, 0}},
        {"GS 128x32 4 acc     ", {128, 32, 0, 0, 16, 8192, SW, 16, 8192, SW, 4, 32, 0, 2, 2, 0}},
        {"GS 128x256 same D   ", {128, 256, 0, 0, 16, 1024, SW, 16, 1024, SW, 1, 256, 0, 2, 2, 0}},
        {"dW  64x24 alias 32ac", {64, 24, 1, 0, 1024, 0, SW, 512, 128, NO, 32, 24, 1, 64, 0, 0}},
        {"dW  64x24 alias 1acc", {64, 24, 1, 0, 1024, 0, SW, 512, 128, NO, 1, 24, 0, 64, 0, 0}},
        {"dW  64x24 K16  32acc", {64, 24, 1, 0, 1024, 1024, SW, 512, 128, NO, 32, 24, 1, 64, 0, 0}},
        {"dW 128x32 alias 16ac", {128, 32, 1, 0, 1024, 0, SW, 512, 128, NO, 16, 32, 0, 128, 0, 0}},
        {"dW  64x24 Bsw128    ", {64, 24, 1, 0, 1024, 0, SW, 16, 1024, SW, 32, 24, 1, 64, 0, 0}},
        {"GS  64x32 + hmma    ", {64, 32, 0, 0, 16, 8192, SW, 16, 8192, SW, 1, 32, 0, 2, 2, 64}},
        {"dW  64x24 + hmma    ", {64, 24, 1, 0, 1024, 0, SW, 512, 128, NO, 32, 24, 1, 64, 0, 64}},
    };
    for (auto& nc : cfgs) {
        cudaMemset(d, 0, 64);
        k_umma<<<1, 160, 200 * 1024>>>(nc.c, 20, d);
        long long h[8];
        cudaError_t e = cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("%s issue %6.1f  done %6.1f cycles/MMA   hmma/warp %lld %lld %lld %lld  %s\n", nc.name, h[0] / 32.0, h[1] / 32.0, h[3], h[4], h[5], h[6],
               e == cudaSuccess ? "" : cudaGetErrorString(e));
        if (e != cudaSuccess) break;
    }
    // hmma alone
    {
        Cfg c = {64, 32, 0, 0, 16, 8192, SW, 16, 8192, SW, 1, 32, 0, 2, 2, 64};
        cudaMemset(d, 0, 64);
        k_umma<<<1, 160, 200 * 1024>>>(c, 0, d);
        long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("hmma alone: cycles per mma.sync per warp %lld %lld %lld %lld\n", h[3], h[4], h[5], h[6]);
    }
    return 0;
}
