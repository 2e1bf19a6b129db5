// Single-warp issue throughput (cycles per warp-instruction) with 8 independent chains.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
#define KERNEL(NAME, DECL, OP, FIN)                                              \
    __global__ void NAME(float* out, long long* cyc, float seed) {               \
        DECL                                                                     \
        long long t0 = clock64();                                                \
        _Pragma("unroll 1") for (int i = 0; i < 64; ++i) {                       \
            _Pragma("unroll") for (int r = 0; r < 4; ++r) {                      \
                _Pragma("unroll") for (int j = 0; j < 8; ++j) { OP }             \
            }                                                                    \
        }                                                                        \
        long long t1 = clock64();                                                \
        FIN                                                                      \
        if (threadIdx.x == 0) cyc[blockIdx.x * 0] = t1 - t0;                     \
    }
#define DECLF float v[8]; for (int j = 0; j < 8; ++j) v[j] = seed + j + threadIdx.x * 1e-3f; float w = seed * 0.999f;
#define FINF float s = 0; for (int j = 0; j < 8; ++j) s += v[j]; out[threadIdx.x] = s;
#define DECLP u64 v[8]; for (int j = 0; j < 8; ++j) v[j] = pk2(seed + j, seed - j + threadIdx.x * 1e-3f); u64 w = pk2(seed * 0.999f, seed * 1.001f);
#define FINP float s = 0; for (int j = 0; j < 8; ++j) { float a, b; upk2(v[j], a, b); s += a + b; } out[threadIdx.x] = s;
KERNEL(t_ffma, DECLF, v[j] = fmaf(v[j], w, w);, FINF)
KERNEL(t_fadd, DECLF, v[j] = v[j] + w;, FINF)
KERNEL(t_ffma2, DECLP, asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[j]) : "l"(w));, FINP)
KERNEL(t_fmul2, DECLP, asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[j]) : "l"(w));, FINP)
KERNEL(t_rcp, DECLF, asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[j]));, FINF)
KERNEL(t_shfl, DECLF, v[j] = __shfl_xor_sync(0xffffffffu, v[j], 1);, FINF)
KERNEL(t_mix, DECLF, if (j & 1) v[j] = fmaf(v[j], w, w); else v[j] = __int_as_float(__float_as_int(v[j]) ^ 0x3);, FINF)
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    long long h;
#define RUN(K, NW) K<<<1, 32 * NW>>>(out, cyc, 1.0001f); K<<<1, 32 * NW>>>(out, cyc, 1.0001f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-10s warps %d: %.2f cycles/instr (one warp's view)\n", #K, NW, h / 2048.0);
    RUN(t_ffma, 1) RUN(t_fadd, 1) RUN(t_ffma2, 1) RUN(t_fmul2, 1) RUN(t_rcp, 1) RUN(t_shfl, 1) RUN(t_mix, 1)
    RUN(t_ffma, 4) RUN(t_ffma2, 4) RUN(t_rcp, 4) RUN(t_ffma, 8) RUN(t_ffma2, 8)
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
