// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/hvs_b200.h"

namespace hvs {

extern std::atomic<uint64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

inline int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

#define HVS_CUDA_TRY(expr)                      \
    do {                                        \
        cudaError_t _e = (expr);                \
        if (_e != cudaSuccess) return (int)_e;  \
    } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory ONCE PER DEVICE (the attribute is per device, and a
// process may drive several GPUs).  `done` is the call site's own static flag array.
template <class Kernel>
inline int set_max_smem_once(Kernel kernel, int bytes, std::atomic<bool> (&done)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
    if (dev >= 0 && dev < 64 && done[dev].load(std::memory_order_acquire)) return HVS_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) done[dev].store(true, std::memory_order_release);
    return HVS_OK;
}
#define HVS_SET_MAX_SMEM(kernel, bytes)                                   \
    do {                                                                  \
        static std::atomic<bool> _hvs_done[64];                           \
        const int _rc = hvs::set_max_smem_once(kernel, bytes, _hvs_done); \
        if (_rc) return _rc;                                              \
    } while (0)

// Returns the launch status without clearing a sticky error of an earlier kernel.
inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HVS_OK : (int)e;
}

// Optional event bracketing of individual launches (hvs_mhc_stream_profile): a ring of event pairs per kernel slot,
// read back (one synchronisation) by hvs_mhc_stream_kernel_ms -- the host never waits inside the profiled region.
struct KernelTimer {
    static constexpr int kRing = 128;
    static constexpr int kSlots = 8;
    cudaEvent_t beg[kSlots][kRing] = {}, end[kSlots][kRing] = {};
    int count[kSlots] = {};                            // launches recorded since profiling was switched on
    std::atomic<bool> enabled{false};
    std::mutex mu;                                     // launches may come from several host threads / streams
};
extern KernelTimer g_timer;
// begin reserves a ring entry under the lock and remembers it for this thread's matching timer_end
inline int& timer_slot_index(int slot) {
    thread_local int idx[KernelTimer::kSlots];
    return idx[slot];
}
inline void timer_begin(int slot, cudaStream_t s) {
    if (!g_timer.enabled.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_timer.mu);
    const int i = g_timer.count[slot]++ % KernelTimer::kRing;
    if (!g_timer.beg[slot][i]) { cudaEventCreate(&g_timer.beg[slot][i]); cudaEventCreate(&g_timer.end[slot][i]); }
    timer_slot_index(slot) = i;
    cudaEventRecord(g_timer.beg[slot][i], s);
}
inline void timer_end(int slot, cudaStream_t s) {
    if (!g_timer.enabled.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_timer.mu);
    cudaEventRecord(g_timer.end[slot][timer_slot_index(slot)], s);
}

// Host -> device upload of a small parameter table on `stream`.  Eagerly the (pageable) source may be freed as soon as the
// call returns; while the stream is being captured into a CUDA graph the copy becomes a graph node that reads the HOST
// buffer again at every replay, so the table is moved into a process-lifetime registry instead of dying with the caller.
int upload_table(void* dst, std::vector<uint8_t>&& table, cudaStream_t stream);

// 2-D bf16 tensor map: [rows][cols] row-major, box [box_rows][64 cols] (=128 B inner), 128-byte swizzle.
int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
// same with an explicit row stride `ld` (elements) and any box height up to 256
int make_tmap_bf16_2d_ld(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
// 3-D (channel, token, stream) view of [T,4,512] bf16: box = 64 channels x box_tokens x 4 streams, 128-byte swizzle.
int make_tmap_bf16_streams3d(CUtensorMap* out, const void* gptr, uint64_t tokens, uint32_t box_tokens);
// 4-D (channel-in-block, token, block, stream) view: one box = 64 x box_tokens x 8 blocks x box_streams streams.
int make_tmap_bf16_streams4d(CUtensorMap* out, const void* gptr, uint64_t tokens, uint32_t box_tokens, uint32_t box_streams = 4);

}  // namespace hvs
