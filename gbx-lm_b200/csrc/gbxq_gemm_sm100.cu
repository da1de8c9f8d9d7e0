// placeholder until the tcgen05 kernel lands
#include "gbxq_common.cuh"
namespace gbxq {
bool gemm_supported(int64_t, int64_t, int64_t, int, int, int, const void*, const void*, const void*) { return false; }
int launch_gemm(const void*, const uint32_t*, const void*, const void*, const void*, void*, int64_t, int64_t, int64_t,
                int, int, cudaStream_t) { return GBXQ_EUNSUPPORTED; }
}  // namespace gbxq
