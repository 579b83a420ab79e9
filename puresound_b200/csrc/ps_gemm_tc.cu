// tcgen05 / TMEM GEMM back end (placeholder until the tensor-core kernel lands).
#include "ps_common.cuh"

namespace ps {
bool gemm_tc_eligible(const ps_gemm_t&) { return false; }
int gemm_tc_launch(const ps_gemm_t&, cudaStream_t) { return PS_ERR_UNSUPPORTED; }
}  // namespace ps

extern "C" int64_t ps_gemm_packed_bytes(int64_t, int64_t) { return 0; }
extern "C" int ps_gemm_pack_weights(const float*, int64_t, int64_t, int64_t, void*, void*) { return PS_ERR_UNSUPPORTED; }
