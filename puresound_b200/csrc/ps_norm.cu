// Normalisation kernels: gLN/gGN statistics merge, bN1d fold, cLN / LayerNorm rows.
#include "ps_common.cuh"

namespace ps {

// One CTA per batch item: Chan-merge the (count, mean, M2) partials in a fixed
// (deterministic) order in fp64, then emit the folded per-channel affine.
__global__ void __launch_bounds__(256) stats_finalize_kernel(const float* __restrict__ partials, int64_t slots,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int64_t C,
                                                             float* __restrict__ scale, float* __restrict__ shift,
                                                             float* __restrict__ meanvar) {
  __shared__ double sn[256], sm[256], s2[256];
  const int64_t b = blockIdx.x;
  pdl_trigger();
  pdl_wait();  // the producer kernel's partials are complete and visible
  stats_finalize_item<0>(partials + b * slots * 3, slots, gamma, beta, eps, C, scale + b * C, shift + b * C,
                         meanvar ? meanvar + b * 2 : nullptr, (int)threadIdx.x, sn, sm, s2);
}

__global__ void bn_fold_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ rm,
                               const float* __restrict__ rv, float eps, int64_t C, float* __restrict__ scale,
                               float* __restrict__ shift) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sc = (w ? w[c] : 1.f) / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = fmaf(-rm[c], sc, b ? b[c] : 0.f);
}

// one warp per row: two-pass mean / biased variance, exactly as torch computes them
__device__ __forceinline__ void row_mean_rstd(const float* __restrict__ x, int64_t C, float eps, int lane, float& mean,
                                              float& rstd) {
  float s = 0.f;
  for (int64_t c = lane; c < C; c += 32) s += x[c];
  mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int64_t c = lane; c < C; c += 32) {
    float dlt = x[c] - mean;
    q = fmaf(dlt, dlt, q);
  }
  rstd = 1.f / sqrtf(warp_sum(q) / (float)C + eps);
}

__global__ void __launch_bounds__(256) rowstats_kernel(const float* __restrict__ x, int64_t rows, int64_t C,
                                                       int64_t row_stride, float eps, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float mean, rstd;
  row_mean_rstd(x + r * row_stride, C, eps, lane, mean, rstd);
  if (lane == 0) {
    out[r * 2] = mean;
    out[r * 2 + 1] = rstd;
  }
}

__global__ void __launch_bounds__(256) rownorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                      float* __restrict__ y, int64_t rows, int64_t C,
                                                      const float* __restrict__ w, const float* __restrict__ b,
                                                      float eps, int act, const float* __restrict__ slope_p) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float slope = slope_p ? __ldg(slope_p) : 0.f;
  const float* xr = x + r * C;
  float mean, rstd;
  row_mean_rstd(xr, C, eps, lane, mean, rstd);
  for (int64_t c = lane; c < C; c += 32) {
    float v = (xr[c] - mean) * rstd;
    v = fmaf(v, w ? w[c] : 1.f, b ? b[c] : 0.f);
    v = apply_act(v, act, slope);
    if (res) v += res[r * C + c];
    y[r * C + c] = v;
  }
}

// Welford partials of a strided region: item n, element (m, r, c) at x + n*bs + m*ms + r*rs + c, m < mid, r < rows, c < C.
// blockIdx.y = item, blockIdx.x = slot: the slot's CTA takes a contiguous share of the item's mid*rows rows.  Used for
// the gLN statistics of U-Net layers, whose GEMM outputs carry rows that are not part of the tensor (unet.py: the
// 2-D convs run as flat framed GEMMs over frequency-padded frames).
__global__ void __launch_bounds__(256) stats_region_kernel(const float* __restrict__ x, int64_t bs, int64_t ms, int64_t rs,
                                                           int64_t mid, int64_t rows, int C, float* __restrict__ partials) {
  __shared__ Wf red[8];
  const int64_t n = blockIdx.y, slots = gridDim.x;
  const int64_t total = mid * rows;
  const int64_t per = (total + slots - 1) / slots;
  const int64_t r0 = blockIdx.x * per, r1 = (r0 + per < total) ? r0 + per : total;
  const float* xb = x + n * bs;
  WfAcc acc;
  acc.init();
  const bool vec = (C & 3) == 0 && ((bs | ms | rs) & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (vec && (C >> 2) <= (int)blockDim.x && total < (1LL << 31)) {
    // a thread keeps its 16-byte column and walks the rows of the slot: one 32-bit division per element instead of four
    // 64-bit ones (the kernel was 5 % of a tse_unet_tcn_v0 forward, run 53)
    const uint32_t C4 = (uint32_t)C >> 2, rpb = blockDim.x / C4, rl = threadIdx.x / C4;
    const int c = (int)(threadIdx.x - rl * C4) * 4;
    const uint32_t rows32 = (uint32_t)rows;
    if (rl < rpb) {
      for (uint32_t row = (uint32_t)r0 + rl; row < (uint32_t)r1; row += rpb) {
        const uint32_t m = row / rows32, r = row - m * rows32;
        const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)m * ms + (int64_t)r * rs + c));
        acc.add(v.x); acc.add(v.y); acc.add(v.z); acc.add(v.w);
      }
    }
  } else if (vec) {
    const int C4 = C >> 2;
    const int64_t cnt = (r1 > r0 ? r1 - r0 : 0) * C4;
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
      const int64_t row = r0 + i / C4;
      const int c = (int)(i % C4) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (row / rows) * ms + (row % rows) * rs + c));
      acc.add(v.x); acc.add(v.y); acc.add(v.z); acc.add(v.w);
    }
  } else {
    const int64_t cnt = (r1 > r0 ? r1 - r0 : 0) * C;
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
      const int64_t row = r0 + i / C;
      const int c = (int)(i % C);
      acc.add(__ldg(xb + (row / rows) * ms + (row % rows) * rs + c));
    }
  }
  const Wf w = wf_block_reduce(acc.finish(), red);
  if (threadIdx.x == 0) {
    float* o = partials + (n * slots + blockIdx.x) * 3;
    o[0] = w.n; o[1] = w.mean; o[2] = w.m2;
  }
}

}  // namespace ps

extern "C" int ps_stats_region(const float* x, int64_t batch, int64_t mid, int64_t rows, int64_t C, int64_t batch_stride,
                               int64_t mid_stride, int64_t row_stride, int64_t slots, float* partials, void* stream) {
  PS_REQUIRE(x && partials && batch > 0 && mid > 0 && rows > 0 && C > 0 && slots > 0 && row_stride >= C);
  if (batch > 65535 || slots > 2147483647LL || C > 2147483647LL) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)slots, (unsigned)batch);
  ps::stats_region_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, batch_stride, mid_stride, row_stride, mid, rows, (int)C, partials);
  PS_CHECK_LAUNCH("stats_region_kernel");
  return PS_OK;
}

extern "C" int ps_stats_finalize(const float* partials, int64_t batch, int64_t slots, const float* gamma,
                                 const float* beta, float eps, int64_t C, float* scale, float* shift, float* meanvar,
                                 void* stream) {
  PS_REQUIRE(partials && scale && shift && batch > 0 && slots > 0 && C > 0);
  cudaError_t le = ps::launch_pdl(ps::stats_finalize_kernel, dim3((unsigned)batch), dim3(256), 0, (cudaStream_t)stream, partials, slots, gamma, beta, eps, C,
                                  scale, shift, meanvar);
  if (le != cudaSuccess) { ps::set_cuda_error(le, "stats_finalize_kernel"); return PS_ERR_CUDA; }
  return PS_OK;
}

extern "C" int ps_bn_fold(const float* weight, const float* bias, const float* running_mean, const float* running_var,
                          float eps, int64_t C, float* scale, float* shift, void* stream) {
  PS_REQUIRE(running_mean && running_var && scale && shift && C > 0);
  ps::bn_fold_kernel<<<(unsigned)ps::cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(weight, bias, running_mean,
                                                                                  running_var, eps, C, scale, shift);
  PS_CHECK_LAUNCH("bn_fold_kernel");
  return PS_OK;
}

extern "C" int ps_rowstats(const float* x, int64_t rows, int64_t C, int64_t row_stride, float eps, float* out,
                           void* stream) {
  PS_REQUIRE(x && out && rows > 0 && C > 0 && row_stride >= C);
  ps::rowstats_kernel<<<(unsigned)ps::cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, C, row_stride, eps, out);
  PS_CHECK_LAUNCH("rowstats_kernel");
  return PS_OK;
}

extern "C" int ps_rownorm(const float* x, const float* res, float* y, int64_t rows, int64_t C, const float* w,
                          const float* b, float eps, int32_t act, const float* slope, void* stream) {
  PS_REQUIRE(x && y && rows > 0 && C > 0);
  ps::rownorm_kernel<<<(unsigned)ps::cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, res, y, rows, C, w, b, eps, act,
                                                                                   slope);
  PS_CHECK_LAUNCH("rownorm_kernel");
  return PS_OK;
}
