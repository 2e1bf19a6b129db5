// Dependent-chain latencies of the instructions on the Sinkhorn critical path (one warp, sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
#define CHAIN(NAME, BODY)                                                          \
    __global__ void NAME(float* out, long long* cyc, float seed) {                 \
        float v = seed + threadIdx.x * 1e-3f, w = seed * 0.5f;                     \
        u64 p = pk2(v, w), q = pk2(w, v);                                          \
        long long t0 = clock64();                                                  \
        _Pragma("unroll 1") for (int i = 0; i < 64; ++i) {                         \
            BODY BODY BODY BODY BODY BODY BODY BODY BODY BODY BODY BODY BODY BODY BODY BODY \
        }                                                                          \
        long long t1 = clock64();                                                  \
        float a, b; upk2(p, a, b);                                                 \
        out[threadIdx.x] = v + w + a + b;                                          \
        if (threadIdx.x == 0) *cyc = t1 - t0;                                      \
    }
CHAIN(k_fadd, v = v + w;)
CHAIN(k_ffma, v = fmaf(v, w, w);)
CHAIN(k_shfl, v = __shfl_xor_sync(0xffffffffu, v, 1);)
CHAIN(k_shfl_add, v = v + __shfl_xor_sync(0xffffffffu, v, 1);)
CHAIN(k_rcp, asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v));)
CHAIN(k_ex2, asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v));)
CHAIN(k_fadd2, asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));)
CHAIN(k_fmul2, asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));)
CHAIN(k_ffma2, asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(q));)
#define XBODY { float a; float b; asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q)); upk2(p, a, b); p = pk2(b, a); }
CHAIN(k_fadd2_cross, XBODY)
#define RBODY { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); v = r * w; }
CHAIN(k_rcp_mul, RBODY)
CHAIN(k_div, v = __fdividef(w, v);)
__global__ void k_lds(float* out, long long* cyc, float seed) {
    __shared__ int s[64];
    s[threadIdx.x] = (threadIdx.x + 1) & 31; s[32 + threadIdx.x] = threadIdx.x;
    __syncwarp();
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) idx = s[idx];
    }
    long long t1 = clock64();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1024); cudaMalloc(&cyc, 8);
    long long h;
#define RUN(K) K<<<1, 32>>>(out, cyc, 1.0001f); K<<<1, 32>>>(out, cyc, 1.0001f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-14s %.1f cycles/op\n", #K, h / 1024.0);
    RUN(k_fadd) RUN(k_ffma) RUN(k_shfl) RUN(k_shfl_add) RUN(k_rcp) RUN(k_ex2) RUN(k_fadd2) RUN(k_fmul2) RUN(k_ffma2) RUN(k_fadd2_cross) RUN(k_rcp_mul) RUN(k_div) RUN(k_lds)
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
