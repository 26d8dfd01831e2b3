// tcgen05 tensor-core path (placeholder until the kernel lands in this round).
#include "common.cuh"
using namespace dexnerf;
extern "C" DEXNERF_API int64_t dexnerf_tc_packed_bytes(const dexnerf_flexible_spec*) { return 0; }
extern "C" DEXNERF_API int dexnerf_tc_pack(const dexnerf_flexible_spec*, const dexnerf_mlp_program*, const float*,
                               void*, void*) {
  set_error("tc_pack: not built");
  return DEXNERF_E_UNSUPPORTED;
}
extern "C" DEXNERF_API int dexnerf_tc_query(const dexnerf_flexible_spec*, const void*, const float*, const float*,
                                const float*, const float*, int64_t, int, float*, void*) {
  set_error("tc_query: not built");
  return DEXNERF_E_UNSUPPORTED;
}
