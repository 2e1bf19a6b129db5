// Development probe for the tcgen05 pieces of the fused backward (not part of the public ABI): one CTA loads one
// 8-token tile of x and dy with the 3-D tensor maps, runs the G = [x; dy] x^T MMA (K-major operands) and the
// dW = x^T E MMA (MN-major A straight from the token tile, E as two bf16 terms along K) and dumps tensor memory.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "umma_sm100.cuh"

namespace hvs {
namespace {

constexpr int kPStage = 65536;
constexpr int kPOffE = kPStage;            // two E tiles of 768 B (mode 1 uses both)
constexpr int kPOffBar = kPOffE + 2048;
constexpr int kPSmem = kPOffBar + 64;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                  const float* __restrict__ e_in, float* __restrict__ out_gs, float* __restrict__ out_dw, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023u) __trap();
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kPOffBar);
    uint64_t* bar_mma = bar_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kPOffBar + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar_full, kPStage);
        for (int cb = 0; cb < 8; ++cb) {
            tma_load_3d(smem + cb * 8192, &tmap_x, bar_full, cb * 64, 0, 0);
            tma_load_3d(smem + cb * 8192 + 4096, &tmap_dy, bar_full, cb * 64, 0, 0);
        }
    }
    // E tile(s): row = logit (24), K = 16: mode 0 -> [hi(8 tokens) | lo(8 tokens)]; mode 1 -> tile0 [hi | 0], tile1 [lo | 0]
    for (int i = threadIdx.x; i < 1024; i += 128) reinterpret_cast<uint16_t*>(smem + kPOffE)[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 24; i += 128) {
        const int tk = i / 24, r = i % 24;
        const float v = e_in[tk * 24 + r];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        uint8_t* t0 = smem + kPOffE;
        const int off = (r >> 3) * 128 + (r & 7) * 16 + tk * 2;
        *reinterpret_cast<__nv_bfloat16*>(t0 + off) = h;
        if (mode == 0) *reinterpret_cast<__nv_bfloat16*>(t0 + 384 + off) = l;
        else *reinterpret_cast<__nv_bfloat16*>(t0 + 768 + off) = l;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_wait(bar_full, 0);
        tc_fence_after();
        const uint32_t s0 = smem_u32(smem);
        // G / x x^T: A = 64 rows [x box | dy box] of one 64-channel block, B = its x box (32 rows), K-major
        const uint32_t id_gs = umma_idesc_bf16(64, 32, 0, 0);
        for (int cb = 0; cb < 8; ++cb)
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t a = s0 + cb * 8192 + ks * 32;
                umma_bf16_ss(tmem_base + 384, umma_smem_desc(a, 16, 1024, kUmmaLayoutSw128),
                             umma_smem_desc(a, 16, 1024, kUmmaLayoutSw128), id_gs, (cb | ks) != 0);
            }
        // dW: block b = (cb, j): A = x atom [8 tokens x 64 channels] read MN-major, D rows = channels
        const uint32_t id_dw = umma_idesc_bf16(64, 24, 1, 0);
        const uint32_t e0 = s0 + kPOffE;
        for (int b = 0; b < 32; ++b) {
            const uint32_t a = s0 + (b >> 2) * 8192 + (b & 3) * 1024;
            const uint32_t d = tmem_base + ((uint32_t)((b & 1) * 16) << 16) + (uint32_t)((b >> 1) * 24);
            if (mode == 0) {
                umma_bf16_ss(d, umma_smem_desc(a, 1024, 0, kUmmaLayoutSw128), umma_smem_desc(e0, 384, 128, kUmmaLayoutNone), id_dw, 0);
            } else {
                umma_bf16_ss(d, umma_smem_desc(a, 1024, 4096, kUmmaLayoutSw128), umma_smem_desc(e0, 384, 128, kUmmaLayoutNone), id_dw, 0);
                umma_bf16_ss(d, umma_smem_desc(a, 1024, 4096, kUmmaLayoutSw128), umma_smem_desc(e0 + 768, 384, 128, kUmmaLayoutNone), id_dw, 1);
            }
        }
        umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    const uint32_t tq = tmem_base + ((uint32_t)(32 * warp) << 16);
    const int gl = 32 * warp + lane;
    uint32_t v[32];
    tmem_ld32(tq + 384, v);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) out_gs[gl * 32 + i] = __uint_as_float(v[i]);
    for (int c = 0; c < 12; ++c) {
        tmem_ld32(tq + c * 32, v);
        tmem_wait_ld();
        for (int i = 0; i < 32; ++i) out_dw[gl * 384 + c * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace hvs

extern "C" int hvs_debug_umma_probe(const void* x, const void* dy, const float* e, float* out_gs, float* out_dw,
                                    int64_t T, int mode, void* stream) {
    using namespace hvs;
    CUtensorMap tx, tdy;
    int rc = make_tmap_bf16_streams3d(&tx, x, (uint64_t)T, 8);
    if (rc) return rc;
    rc = make_tmap_bf16_streams3d(&tdy, dy, (uint64_t)T, 8);
    if (rc) return rc;
    HVS_CUDA_TRY(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
    umma_probe_kernel<<<1, 128, kPSmem, (cudaStream_t)stream>>>(tx, tdy, e, out_gs, out_dw, mode);
    return launch_status();
}
