// ABI bookkeeping + tensor-map encode through the driver entry point (no -lcuda link).
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace hvs {

std::atomic<uint64_t> g_launches{0};
KernelTimer g_timer;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The encode is a driver-API call and needs the primary context current on THIS thread; a thread that has only been
// handed a device ordinal (e.g. PyTorch's autograd workers) has not bound it yet.  cudaFree(nullptr) binds it -- once
// per thread, because it is not allowed while a stream of the thread is being captured into a CUDA graph (the warm-up
// run every capture needs has then already done it).
static void bind_primary_context_once() {
    thread_local bool bound = false;
    if (!bound) {
        cudaFree(nullptr);
        bound = true;
    }
}

static EncodeTiledFn encode_fn() {
    // function-local static: initialised exactly once, thread-safe (C++11), so a second thread (e.g. an autograd
    // worker) can never observe a half-initialised lookup
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(p);
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return HVS_ERR_DRIVER;
    bind_primary_context_once();
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS && getenv("HVS_DEBUG"))
        fprintf(stderr, "[hvs_b200] cuTensorMapEncodeTiled -> CUresult %d (ptr %p rows %llu cols %llu box_rows %u)\n", (int)r,
                gptr, (unsigned long long)rows, (unsigned long long)cols, box_rows);
    return r == CUDA_SUCCESS ? HVS_OK : HVS_ERR_DRIVER;
}

int make_tmap_bf16_2d_ld(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return HVS_ERR_DRIVER;
    bind_primary_context_once();
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS && getenv("HVS_DEBUG"))
        fprintf(stderr, "[hvs_b200] cuTensorMapEncodeTiled(ld) -> CUresult %d (ptr %p rows %llu cols %llu ld %llu box_rows %u)\n",
                (int)r, gptr, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return r == CUDA_SUCCESS ? HVS_OK : HVS_ERR_DRIVER;
}

int upload_table(void* dst, std::vector<uint8_t>&& table, cudaStream_t stream) {
    if (table.empty()) return HVS_OK;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) st = cudaStreamCaptureStatusNone;
    if (st == cudaStreamCaptureStatusActive) {
        static std::mutex mu;
        static std::vector<std::vector<uint8_t>>* kept = new std::vector<std::vector<uint8_t>>();   // outlives every graph
        std::lock_guard<std::mutex> lock(mu);
        kept->push_back(std::move(table));
        const std::vector<uint8_t>& t = kept->back();
        return (int)cudaMemcpyAsync(dst, t.data(), t.size(), cudaMemcpyHostToDevice, stream);
    }
    return (int)cudaMemcpyAsync(dst, table.data(), table.size(), cudaMemcpyHostToDevice, stream);
}

// 3-D view of a [T, 4, 512] bf16 stream tensor as (channel, token, stream): a box of 64 channels x 8 tokens
// x 4 streams lands in shared memory as four 1 KB swizzle atoms [stream][token][64 channels].
int make_tmap_bf16_streams3d(CUtensorMap* out, const void* gptr, uint64_t tokens, uint32_t box_tokens) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return HVS_ERR_DRIVER;
    bind_primary_context_once();
    cuuint64_t gdim[3] = {512, tokens, 4};
    cuuint64_t gstride[2] = {4096, 1024};       // bytes: token stride, stream stride
    cuuint32_t box[3] = {64, box_tokens, 4};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS && getenv("HVS_DEBUG"))
        fprintf(stderr, "[hvs_b200] cuTensorMapEncodeTiled(3d) -> CUresult %d (ptr %p tokens %llu)\n", (int)r, gptr,
                (unsigned long long)tokens);
    return r == CUDA_SUCCESS ? HVS_OK : HVS_ERR_DRIVER;
}

// 4-D view of [T, 4, 512] bf16 as (channel in a 64-channel block, token, block, stream): ONE box of
// 64 x box_tokens x 8 x 4 lands as 32 swizzle atoms [stream][block][token][64 channels] (32 KB for 8 tokens).
int make_tmap_bf16_streams4d(CUtensorMap* out, const void* gptr, uint64_t tokens, uint32_t box_tokens, uint32_t box_streams) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return HVS_ERR_DRIVER;
    bind_primary_context_once();
    cuuint64_t gdim[4] = {64, tokens, 8, 4};
    cuuint64_t gstride[3] = {4096, 128, 1024};  // bytes: token, 64-channel block, stream
    cuuint32_t box[4] = {64, box_tokens, 8, box_streams};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(gptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS && getenv("HVS_DEBUG"))
        fprintf(stderr, "[hvs_b200] cuTensorMapEncodeTiled(4d) -> CUresult %d (ptr %p tokens %llu)\n", (int)r, gptr,
                (unsigned long long)tokens);
    return r == CUDA_SUCCESS ? HVS_OK : HVS_ERR_DRIVER;
}

}  // namespace hvs

extern "C" {

int hvs_abi_version(void) { return 1; }

uint64_t hvs_launch_count(void) { return hvs::g_launches.load(std::memory_order_relaxed); }

int hvs_mhc_stream_profile(int enable) {
    std::lock_guard<std::mutex> lock(hvs::g_timer.mu);
    hvs::g_timer.enabled.store(enable != 0);
    if (enable)
        for (int i = 0; i < hvs::KernelTimer::kSlots; ++i) hvs::g_timer.count[i] = 0;
    return HVS_OK;
}

static int kernel_ms(float* out_host, int slots);
int hvs_mhc_stream_kernel_ms(float* out4_host) { return kernel_ms(out4_host, 4); }
int hvs_profile_kernel_ms(float* out8_host) { return kernel_ms(out8_host, hvs::KernelTimer::kSlots); }

static int kernel_ms(float* out4_host, int slots) {
    if (!out4_host) return HVS_ERR_BAD_ARG;
    using hvs::KernelTimer;
    std::lock_guard<std::mutex> lock(hvs::g_timer.mu);
    for (int i = 0; i < slots; ++i) {
        out4_host[i] = -1.0f;
        const int n = hvs::g_timer.count[i] < KernelTimer::kRing ? hvs::g_timer.count[i] : KernelTimer::kRing;
        if (n == 0) continue;
        double sum = 0.0;
        for (int k = 0; k < n; ++k) {
            cudaError_t e = cudaEventSynchronize(hvs::g_timer.end[i][k]);
            if (e != cudaSuccess) return (int)e;
            float ms = 0.f;
            e = cudaEventElapsedTime(&ms, hvs::g_timer.beg[i][k], hvs::g_timer.end[i][k]);
            if (e != cudaSuccess) return (int)e;
            sum += ms;
        }
        out4_host[i] = (float)(sum / n);
    }
    return HVS_OK;
}

const char* hvs_error_string(int code) {
    switch (code) {
        case HVS_OK: return "ok";
        case HVS_ERR_BAD_ARG: return "bad argument (null pointer or negative size)";
        case HVS_ERR_UNSUPPORTED: return "shape or option not supported by the sm_100a kernels";
        case HVS_ERR_ALIGNMENT: return "pointer or stride alignment";
        case HVS_ERR_WORKSPACE: return "workspace too small";
        case HVS_ERR_DRIVER: return "CUDA driver entry point unavailable or tensor-map encode failed";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

}  // extern "C"
