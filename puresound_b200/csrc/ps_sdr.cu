// On-device SDR / SI-SNR scoring (reference loss/sdr.py:7-185 `SDRLoss.forward`, :263-299 `si_snr`): one pass over the two
// waveforms accumulates the five moments  sum s1, sum s2, sum s1^2, sum s2^2, sum s1 s2  per row in fp64, from which the
// zero-mean inner products, the scaled target  alpha s2  (alpha = <s1,s2> / (<s2,s2> + eps)), the noise energy and
// 10 log10(target / (noise + eps) + eps) follow in closed form - the reference's five elementwise passes and four reductions
// become one read of each signal.  fp64 accumulation keeps the closed form accurate up to ~120 dB.
#include "ps_common.cuh"

namespace ps {

__global__ void __launch_bounds__(1024) sdr_kernel(const float* __restrict__ s1, const float* __restrict__ s2, int64_t L,
                                                   int64_t stride1, int64_t stride2, int scaled, int scale_dependent, int zero_mean,
                                                   float tau, float eps, float* __restrict__ out) {
  __shared__ double red[32][5];
  const int64_t r = blockIdx.x;
  const float* a = s1 + r * stride1;
  const float* b = s2 + r * stride2;
  double m[5] = {0, 0, 0, 0, 0};
  for (int64_t i = threadIdx.x; i < L; i += blockDim.x) {
    const double x = a[i], y = b[i];
    m[0] += x; m[1] += y; m[2] += x * x; m[3] += y * y; m[4] += x * y;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k)
    for (int o = 16; o > 0; o >>= 1) m[k] += __shfl_xor_sync(0xffffffffu, m[k], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < 5; ++k) red[warp][k] = m[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[5] = {0, 0, 0, 0, 0};
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w)
      for (int k = 0; k < 5; ++k) t[k] += red[w][k];
    const double n = (double)L;
    double a11 = t[2], a22 = t[3], a12 = t[4];
    if (zero_mean) {
      const double m1 = t[0] / n, m2 = t[1] / n;
      a11 -= n * m1 * m1; a22 -= n * m2 * m2; a12 -= n * m1 * m2;
    }
    const double alpha = scaled ? a12 / (a22 + (double)eps) : 1.0;
    const double target = alpha * alpha * a22;
    double noise = scale_dependent ? (a11 - 2.0 * a12 + a22) : (a11 - 2.0 * alpha * a12 + alpha * alpha * a22);
    if (noise < 0.0) noise = 0.0;
    noise += (double)tau * target;
    out[r] = (float)(10.0 * log10(target / (noise + (double)eps) + (double)eps));
  }
}

}  // namespace ps

extern "C" int ps_sdr(const float* s1, const float* s2, int64_t rows, int64_t L, int64_t stride1, int64_t stride2, int32_t scaled,
                      int32_t scale_dependent, int32_t zero_mean, float tau, float eps, float* out, void* stream) {
  PS_REQUIRE(s1 && s2 && out && rows > 0 && L > 0 && stride1 >= 0 && stride2 >= 0);
  if (rows > 2147483647LL) return PS_ERR_UNSUPPORTED;
  ps::sdr_kernel<<<(unsigned)rows, 1024, 0, (cudaStream_t)stream>>>(s1, s2, L, stride1, stride2, scaled, scale_dependent, zero_mean, tau,
                                                                 eps, out);
  PS_CHECK_LAUNCH("sdr_kernel");
  return PS_OK;
}
