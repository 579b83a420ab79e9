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

bool pdl_enabled() {
  static EnvInt env;
  return env.get("PS_PDL", 0) != 0;
}

bool order_reversed() {
  static EnvInt env;
  static std::atomic<unsigned> flip{0};
  if (env.get("PS_ORDER_ALT", 0) == 0) return false;
  return (flip.fetch_add(1, std::memory_order_relaxed) & 1u) != 0;
}

int gemm_simt_launch(const ps_gemm_t& d, cudaStream_t s);
bool gemm_tc_eligible(const ps_gemm_t& d);
int gemm_tc_launch(const ps_gemm_t& d, cudaStream_t s);
bool tc_pair();
bool gemm_pair_ln_eligible(const ps_gemm_t& d);
bool gemm_rows_eligible(const ps_gemm_t& d);

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
    case 4: return (int64_t)sizeof(ps_gated_t);
    case 5: return (int64_t)sizeof(ps_stream_hop_block_t);
    case 6: return (int64_t)sizeof(ps_stream_hop_t);
    default: return -1;
  }
}

extern "C" int64_t ps_gemm_stats_slots(int64_t rows, int64_t M) {
  if (rows <= 0 || M <= 0) return 0;
  return ps::cdiv(rows, 128) * ps::cdiv(M, 128);
}

namespace ps {
bool gemm_wide_eligible(const ps_gemm_t& d, int sms);
bool gemm_pair_few_tiles(const ps_gemm_t& d, int sms);
}

extern "C" int ps_gemm_path(const ps_gemm_t* dp) {
  PS_REQUIRE(dp != nullptr);
  const ps_gemm_t& d = *dp;
  const bool tc = (d.backend == PS_GEMM_TCGEN05 || d.backend == PS_GEMM_AUTO) && ps::gemm_tc_eligible(d);
  if (d.ln_eps > 0.f && !(tc && ps::tc_pair() && (ps::gemm_pair_ln_eligible(d) || ps::gemm_rows_eligible(d)))) return 0;
  if (!tc) return 0;
  if (!ps::tc_pair()) return 1;
  if (ps::gemm_rows_eligible(d)) return 4;
  int dev = 0, sms = 0;
  if (int rc = ps::current_device(&dev)) return rc;
  if (int rc = ps::sm_count_of(dev, &sms)) return rc;
  if (ps::gemm_pair_few_tiles(d, sms)) return 2;
  return ps::gemm_wide_eligible(d, sms) ? 3 : 2;
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
  if (d.fin_scale) PS_REQUIRE(d.stats_partials && d.fin_shift && d.fin_counter);
  cudaStream_t s = (cudaStream_t)stream;
  if (d.ln_eps > 0.f) {
    // Linear -> LayerNorm (-> + residual): fused epilogue on the CTA-pair kernel, else GEMM then ps_rownorm in place
    PS_REQUIRE(!d.stats_partials && d.epi_act == PS_ACT_NONE && d.y_row_stride == d.M && d.y_batch_stride == d.rows * d.M);
    const bool fuse = (d.backend == PS_GEMM_TCGEN05 || d.backend == PS_GEMM_AUTO) && ps::tc_pair() && ps::gemm_tc_eligible(d) &&
                      (ps::gemm_pair_ln_eligible(d) || ps::gemm_rows_eligible(d));
    if (fuse) return ps::gemm_tc_launch(d, s);
    if (d.backend == PS_GEMM_TCGEN05) return PS_ERR_UNSUPPORTED;
    PS_REQUIRE(!d.residual || (d.res_row_stride == d.M && d.res_batch_stride == d.rows * d.M && d.residual != d.Y));
    ps_gemm_t dd = d;
    dd.ln_eps = 0.f; dd.ln_gamma = nullptr; dd.ln_beta = nullptr; dd.residual = nullptr;
    const int rc = ps_gemm(&dd, stream);
    if (rc != PS_OK) return rc;
    return ps_rownorm(d.Y, d.residual, d.Y, d.batch * d.rows, d.M, d.ln_gamma, d.ln_beta, d.ln_eps, PS_ACT_NONE, nullptr, stream);
  }
  const bool tc = (d.backend == PS_GEMM_TCGEN05 || d.backend == PS_GEMM_AUTO) && ps::gemm_tc_eligible(d);
  if (d.backend == PS_GEMM_TCGEN05 && !tc) return PS_ERR_UNSUPPORTED;
  // the CTA-pair kernels fuse the statistics finalize (the few-channel kernel of ps_gemm_rows.cu does not)
  if (tc && ps::tc_pair() && !ps::gemm_rows_eligible(d)) return ps::gemm_tc_launch(d, s);
  // the other kernels leave the finalize to a follow-up launch
  ps_gemm_t dd = d;
  dd.fin_scale = nullptr;
  const int rc = tc ? ps::gemm_tc_launch(dd, s) : ps::gemm_simt_launch(dd, s);
  if (rc != PS_OK || !d.fin_scale) return rc;
  return ps_stats_finalize(d.stats_partials, d.batch, ps_gemm_stats_slots(d.rows, d.M), d.fin_gamma, d.fin_beta, d.fin_eps, d.M,
                           d.fin_scale, d.fin_shift, nullptr, stream);
}
