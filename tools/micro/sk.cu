// Sinkhorn forward-iteration loop (lane = token, packed fp32x2 rows) timed alone, and MUFU throughput.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rcpa(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__global__ void k_fwd(float* out, long long* cyc, int iters, float eps) {
    __shared__ float sk[24 * 64];
    const int tk = threadIdx.x & 7, part = (threadIdx.x & 31) >> 3;
    u64 P[4][2];
    for (int i = 0; i < 4; ++i) { P[i][0] = pk2(1.f + 0.01f * i + tk * 0.001f, 0.9f); P[i][1] = pk2(1.1f, 1.f - 0.02f * i); }
    const u64 eps2 = pk2(eps, eps);
    float* skl = sk + tk * 8;
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
        float dr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { float sa, sb; upk2(add2(P[i][0], P[i][1]), sa, sb); dr[i] = (sa + sb) + eps; }
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float rr = rcpa(dr[i]); const u64 rr2 = pk2(rr, rr); P[i][0] = mul2(P[i][0], rr2); P[i][1] = mul2(P[i][1], rr2); }
        const u64 c01 = add2(add2(add2(P[0][0], P[1][0]), add2(P[2][0], P[3][0])), eps2);
        const u64 c23 = add2(add2(add2(P[0][1], P[1][1]), add2(P[2][1], P[3][1])), eps2);
        float c0, c1, c2, c3; upk2(c01, c0, c1); upk2(c23, c2, c3);
        const u64 rc01 = pk2(rcpa(c0), rcpa(c1)), rc23 = pk2(rcpa(c2), rcpa(c3));
#pragma unroll
        for (int i = 0; i < 4; ++i) { P[i][0] = mul2(P[i][0], rc01); P[i][1] = mul2(P[i][1], rc23); }
        if (part == 0) { float4* o = reinterpret_cast<float4*>(skl + (k % 24) * 64); o[0] = make_float4(dr[0], dr[1], dr[2], dr[3]); o[1] = make_float4(c0, c1, c2, c3); }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 4; ++i) { float a, b; upk2(P[i][0], a, b); s += a + b; upk2(P[i][1], a, b); s += a + b; }
    out[threadIdx.x] = s + sk[threadIdx.x];
    if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}
// 8 independent rcp chains: XU throughput seen by one warp
__global__ void k_mufu(float* out, long long* cyc, float seed) {
    float v[8]; for (int j = 0; j < 8; ++j) v[j] = seed + j + threadIdx.x * 1e-3f;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = rcpa(v[j]) + 0.25f;
    }
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 8; ++j) s += v[j]; out[threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 256);
    long long h[8];
    for (int nw : {1, 4, 8}) {
        k_mufu<<<1, 32 * nw>>>(out, cyc, 1.0001f); k_mufu<<<1, 32 * nw>>>(out, cyc, 1.0001f);
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("mufu+fadd  warps %d: %.2f cycles per (rcp,add) pair, one warp's view\n", nw, h[0] / 2048.0);
    }
    for (int nw : {1, 4}) {
        k_fwd<<<1, 32 * nw>>>(out, cyc, 2000, 1e-8f); k_fwd<<<1, 32 * nw>>>(out, cyc, 2000, 1e-8f);
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("fwd loop   warps %d: %.1f cycles per iteration\n", nw, h[0] / 2000.0);
    }
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
