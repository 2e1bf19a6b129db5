// K1 backward (placeholder until the fused kernel lands in this file).
#include "common.cuh"

extern "C" size_t hvs_mhc_stream_bwd_workspace(int64_t T, int n, int C) {
    (void)T; (void)n; (void)C;
    return 0;
}

extern "C" int hvs_mhc_stream_bwd(const void*, const void*, const float*, const float*, const float*, const float*,
                                  void*, float*, float*, float*, float*, int64_t, int, int, int, float, float,
                                  uint32_t, void*, size_t, void*) {
    return HVS_ERR_UNSUPPORTED;
}
