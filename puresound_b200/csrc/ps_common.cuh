// Shared device/host helpers for the puresound_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/puresound_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "puresound_b200 kernels are written for sm_100a (B200) only"
#endif

namespace ps {

void set_cuda_error(cudaError_t e, const char* where);

#define PS_REQUIRE(cond)                 \
  do {                                   \
    if (!(cond)) return PS_ERR_INVALID_ARG; \
  } while (0)

#define PS_CHECK_LAUNCH(where)                      \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    if (e__ != cudaSuccess) {                       \
      ps::set_cuda_error(e__, where);               \
      return PS_ERR_CUDA;                           \
    }                                               \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// if-chain with the hot cases first (PReLU, none): a switch here becomes an indirect branch (BRX) per element
// once it is inlined into unrolled loops, which measured 4x more instructions in the GEMM producers
__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == PS_ACT_PRELU) return v > 0.f ? v : v * slope;  // NaN stays NaN: (NaN > 0) is false, NaN*slope = NaN
  if (act == PS_ACT_NONE) return v;
  if (act == PS_ACT_RELU) return (v != v) ? v : fmaxf(v, 0.f);  // torch.relu propagates NaN
  if (act == PS_ACT_TANH) return tanhf(v);
  if (act == PS_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

// ---- Welford / Chan partials: (count, mean, M2) --------------------------------
struct Wf {
  float n, mean, m2;
};

__device__ __forceinline__ Wf wf_merge(Wf a, Wf b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  float n = a.n + b.n;
  float d = b.mean - a.mean;
  float f = b.n / n;
  Wf r;
  r.n = n;
  r.mean = a.mean + d * f;
  r.m2 = a.m2 + b.m2 + d * d * a.n * f;
  return r;
}

// per-thread accumulation of a few values as (n, sum, sumsq about a pivot)
struct WfAcc {
  float n, pivot, s, ss;
  __device__ __forceinline__ void init() { n = 0.f; pivot = 0.f; s = 0.f; ss = 0.f; }
  __device__ __forceinline__ void add(float v) {
    if (n == 0.f) pivot = v;
    float d = v - pivot;
    s += d;
    ss += d * d;
    n += 1.f;
  }
  __device__ __forceinline__ Wf finish() const {
    Wf r;
    r.n = n;
    if (n == 0.f) { r.mean = 0.f; r.m2 = 0.f; return r; }
    float md = s / n;
    r.mean = pivot + md;
    r.m2 = ss - s * md;
    if (r.m2 < 0.f) r.m2 = 0.f;
    return r;
  }
};

__device__ __forceinline__ Wf wf_warp_reduce(Wf v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Wf b;
    b.n = __shfl_xor_sync(0xffffffffu, v.n, o);
    b.mean = __shfl_xor_sync(0xffffffffu, v.mean, o);
    b.m2 = __shfl_xor_sync(0xffffffffu, v.m2, o);
    // merge in a lane-symmetric order so every lane ends with the same bits
    Wf lo = (threadIdx.x & o) ? b : v;
    Wf hi = (threadIdx.x & o) ? v : b;
    v = wf_merge(lo, hi);
  }
  return v;
}

// block reduce; result valid in thread 0.  smem must hold (blockDim.x/32) Wf.
__device__ __forceinline__ Wf wf_block_reduce(Wf v, Wf* smem) {
  v = wf_warp_reduce(v);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int nw = (blockDim.x + 31) >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  Wf r;
  r.n = 0.f; r.mean = 0.f; r.m2 = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nw; ++i) r = wf_merge(r, smem[i]);
  }
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ps
