// C-ABI glue: status strings, device check, GEMM dispatch between the exact-fp32
// CUDA-core kernel and the tcgen05/TMEM kernel.
#include <stdio.h>
#include <string.h>

#include "ps_common.cuh"

namespace ps {

static thread_local char g_err[256] = "";

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int gemm_simt_launch(const ps_gemm_t& d, cudaStream_t s);
bool gemm_tc_eligible(const ps_gemm_t& d);
int gemm_tc_launch(const ps_gemm_t& d, cudaStream_t s);

}  // namespace ps

extern "C" const char* ps_error_string(int status) {
  switch (status) {
    case PS_OK: return "ok";
    case PS_ERR_INVALID_ARG: return "invalid argument";
    case PS_ERR_UNSUPPORTED: return "configuration not supported by the B200 engine";
    case PS_ERR_CUDA: return "CUDA error";
    case PS_ERR_NO_DEVICE: return "no sm_100 device";
    default: return "unknown status";
  }
}

extern "C" const char* ps_last_cuda_error(void) { return ps::g_err; }

extern "C" int ps_version(void) { return PS_ABI_VERSION; }

extern "C" int ps_device_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { ps::set_cuda_error(e, "cudaGetDevice"); return PS_ERR_NO_DEVICE; }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) { ps::set_cuda_error(e, "cudaDeviceGetAttribute"); return PS_ERR_NO_DEVICE; }
  return major == 10 ? 1 : 0;
}

extern "C" int64_t ps_struct_size(int which) {
  switch (which) {
    case 0: return (int64_t)sizeof(ps_gemm_t);
    case 1: return (int64_t)sizeof(ps_dwconv_t);
    case 2: return (int64_t)sizeof(ps_lstm_t);
    case 3: return (int64_t)sizeof(ps_stream_dw_t);
    default: return -1;
  }
}

extern "C" int64_t ps_gemm_stats_slots(int64_t rows, int64_t M) {
  if (rows <= 0 || M <= 0) return 0;
  return ps::cdiv(rows, 128) * ps::cdiv(M, 128);
}

extern "C" int ps_gemm(const ps_gemm_t* dp, void* stream) {
  PS_REQUIRE(dp != nullptr);
  const ps_gemm_t& d = *dp;
  PS_REQUIRE(d.X && d.W && d.Y && d.batch > 0 && d.rows > 0 && d.M > 0 && d.K > 0);
  PS_REQUIRE(d.x_row_stride > 0 && d.w_row_stride >= d.K && d.y_row_stride >= d.M);
  PS_REQUIRE(d.pro_mode >= PS_PRO_NONE && d.pro_mode <= PS_PRO_MASK);
  if (d.pro_mode == PS_PRO_AFFINE || d.pro_mode == PS_PRO_ROWNORM) PS_REQUIRE(d.pro_a && d.pro_b);
  if (d.pro_mode == PS_PRO_ROWNORM) PS_REQUIRE(d.pro_rowstats);
  if (d.pro_mode == PS_PRO_MASK) PS_REQUIRE(d.X2);
  else PS_REQUIRE(d.X2 == nullptr);
  if (d.pro_mode != PS_PRO_NONE && d.pro_act == PS_ACT_PRELU) PS_REQUIRE(d.pro_slope);
  if (d.epi_act == PS_ACT_PRELU) PS_REQUIRE(d.epi_slope);
  cudaStream_t s = (cudaStream_t)stream;
  if (d.backend == PS_GEMM_TCGEN05) {
    if (!ps::gemm_tc_eligible(d)) return PS_ERR_UNSUPPORTED;
    return ps::gemm_tc_launch(d, s);
  }
  if (d.backend == PS_GEMM_AUTO && ps::gemm_tc_eligible(d)) return ps::gemm_tc_launch(d, s);
  return ps::gemm_simt_launch(d, s);
}
