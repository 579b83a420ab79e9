// Frame-by-frame streaming kernels.  All per-stream state (dilation-history ring
// buffers, encoder sample history, overlap-add accumulator, step counter) lives in
// device memory so that one hop of every concurrent stream is a fixed kernel chain
// that can be captured once in a CUDA graph and replayed.
#include "ps_common.cuh"

namespace ps {

constexpr int SD_THREADS = 128;
constexpr int SD_MAXV = 16;  // channels per thread -> C <= 2048

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect red from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < SD_THREADS / 32; ++i) t += red[i];
  return t;
}

// row-wise norm + PReLU on register-resident channels. kind 0: cLN, 1: per-channel affine
__device__ __forceinline__ void row_norm_act(float (&v)[SD_MAXV], int nv, int64_t C, int kind, float eps,
                                             const float* __restrict__ a, const float* __restrict__ b, float slope,
                                             float* red) {
  float mean = 0.f, rstd = 1.f;
  if (kind == 0) {
    float s = 0.f;
    for (int i = 0; i < nv; ++i) s += v[i];
    mean = block_sum(s, red) / (float)C;
    float q = 0.f;
    for (int i = 0; i < nv; ++i) { float dl = v[i] - mean; q = fmaf(dl, dl, q); }
    rstd = 1.f / sqrtf(block_sum(q, red) / (float)C + eps);
  }
  for (int i = 0; i < nv; ++i) {
    const int64_t c = threadIdx.x + (int64_t)i * SD_THREADS;
    float x = (kind == 0) ? (v[i] - mean) * rstd : v[i];
    x = fmaf(x, a[c], b[c]);
    v[i] = x > 0.f ? x : x * slope;
  }
}

__global__ void __launch_bounds__(SD_THREADS) stream_dwconv_kernel(const ps_stream_dw_t d) {
  __shared__ float red[SD_THREADS / 32];
  const int64_t s = blockIdx.x;
  const int64_t C = d.C;
  const int64_t RL = (int64_t)(d.P - 1) * d.dilation + 1;
  const int64_t step = *d.step;
  float v[SD_MAXV];
  int nv = 0;
  for (int64_t c = threadIdx.x; c < C; c += SD_THREADS) v[nv++] = d.u[s * C + c];
  row_norm_act(v, nv, C, d.norm_kind, d.eps, d.n1_a, d.n1_b, __ldg(d.slope1), red);
  float* ring = d.ring + s * RL * C;
  const int64_t slot = step % RL;
  float acc[SD_MAXV];
  for (int i = 0; i < nv; ++i) {
    const int64_t c = threadIdx.x + (int64_t)i * SD_THREADS;
    ring[slot * C + c] = v[i];
    float a = d.bias ? d.bias[c] : 0.f;
    for (int p = 0; p < d.P; ++p) {
      const int64_t back = (int64_t)(d.P - 1 - p) * d.dilation;
      float x;
      if (back == 0) x = v[i];
      else if (back > step) x = 0.f;  // causal zero padding before the stream started
      else x = ring[((step - back) % RL) * C + c];
      a = fmaf(d.w[c * d.P + p], x, a);
    }
    acc[i] = a;
  }
  row_norm_act(acc, nv, C, d.norm_kind, d.eps, d.n2_a, d.n2_b, __ldg(d.slope2), red);
  for (int i = 0; i < nv; ++i) d.y[s * C + threadIdx.x + (int64_t)i * SD_THREADS] = acc[i];
}

__global__ void stream_push_kernel(const float* __restrict__ chunk, float* __restrict__ hist, float* __restrict__ frame,
                                   int64_t win, int64_t hop) {
  const int64_t s = blockIdx.x;
  const int64_t keep = win - hop;
  float* f = frame + s * win;
  float* h = hist + s * keep;
  const float* c = chunk + s * hop;
  for (int64_t i = threadIdx.x; i < win; i += blockDim.x) f[i] = (i < keep) ? h[i] : c[i - keep];
  __syncthreads();
  for (int64_t i = threadIdx.x; i < keep; i += blockDim.x) h[i] = f[i + hop];
}

__global__ void stream_ola_kernel(const float* __restrict__ frame, float* __restrict__ acc, float* __restrict__ out,
                                  int64_t win, int64_t hop, int constraint) {
  extern __shared__ float tmp[];  // win floats
  const int64_t s = blockIdx.x;
  float* a = acc + s * win;
  for (int64_t i = threadIdx.x; i < win; i += blockDim.x) tmp[i] = a[i] + frame[s * win + i];
  __syncthreads();
  for (int64_t i = threadIdx.x; i < win; i += blockDim.x) {
    if (i < hop) {
      float v = tmp[i];
      if (constraint == 1) v = (v != v) ? v : fminf(fmaxf(v, -1.f), 1.f);
      else if (constraint == 2) v = 1.f / (1.f + expf(-v));
      out[s * hop + i] = v;
    }
    a[i] = (i + hop < win) ? tmp[i + hop] : 0.f;
  }
}

__global__ void stream_advance_kernel(int64_t* step) { *step += 1; }

}  // namespace ps

extern "C" int ps_stream_dwconv_step(const ps_stream_dw_t* dp, void* stream) {
  PS_REQUIRE(dp != nullptr);
  const ps_stream_dw_t& d = *dp;
  PS_REQUIRE(d.u && d.y && d.ring && d.step && d.w && d.streams > 0 && d.C > 0 && d.P >= 1 && d.dilation >= 1);
  PS_REQUIRE(d.n1_a && d.n1_b && d.n2_a && d.n2_b && d.slope1 && d.slope2 && (d.norm_kind == 0 || d.norm_kind == 1));
  if (d.C > (int64_t)ps::SD_THREADS * ps::SD_MAXV) return PS_ERR_UNSUPPORTED;
  ps::stream_dwconv_kernel<<<(unsigned)d.streams, ps::SD_THREADS, 0, (cudaStream_t)stream>>>(d);
  PS_CHECK_LAUNCH("stream_dwconv_kernel");
  return PS_OK;
}

extern "C" int ps_stream_push(const float* chunk, float* hist, float* frame, int64_t streams, int64_t win, int64_t hop,
                              void* stream) {
  PS_REQUIRE(chunk && frame && streams > 0 && hop > 0 && win >= hop && (win == hop || hist));
  ps::stream_push_kernel<<<(unsigned)streams, 128, 0, (cudaStream_t)stream>>>(chunk, hist, frame, win, hop);
  PS_CHECK_LAUNCH("stream_push_kernel");
  return PS_OK;
}

extern "C" int ps_stream_ola(const float* frame, float* acc, float* out, int64_t streams, int64_t win, int64_t hop,
                             int32_t constraint, void* stream) {
  PS_REQUIRE(frame && acc && out && streams > 0 && hop > 0 && win >= hop && win * sizeof(float) <= 48 * 1024);
  ps::stream_ola_kernel<<<(unsigned)streams, 128, (size_t)win * sizeof(float), (cudaStream_t)stream>>>(
      frame, acc, out, win, hop, constraint);
  PS_CHECK_LAUNCH("stream_ola_kernel");
  return PS_OK;
}

extern "C" int ps_stream_advance(int64_t* step, void* stream) {
  PS_REQUIRE(step != nullptr);
  ps::stream_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
  PS_CHECK_LAUNCH("stream_advance_kernel");
  return PS_OK;
}
