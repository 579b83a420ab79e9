// Shared device/host helpers for the puresound_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>

#include <atomic>

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

// ---- host-side caches of the launchers ------------------------------------------------------------------------
// The C-ABI promises re-entrancy across host threads, streams and devices.  The launchers keep three kinds of cached
// facts, all process-wide, monotonic and idempotent (a lost race only repeats an idempotent call): the SM count of a
// device, "this kernel's dynamic shared-memory limit has been raised on this device", and integer A/B switches read
// from the environment.  All are std::atomic; a "done" flag is stored only AFTER the call it stands for succeeded.
constexpr int PS_MAX_DEVICES = 64;

inline int current_device(int* dev) {
  cudaError_t e = cudaGetDevice(dev);
  if (e != cudaSuccess || *dev < 0 || *dev >= PS_MAX_DEVICES) { set_cuda_error(e, "cudaGetDevice"); return PS_ERR_CUDA; }
  return PS_OK;
}

inline int sm_count_of(int dev, int* out) {
  static std::atomic<int> cache[PS_MAX_DEVICES];
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess || v <= 0) { set_cuda_error(e, "cudaDeviceGetAttribute(multiProcessorCount)"); return PS_ERR_CUDA; }
    cache[dev].store(v, std::memory_order_relaxed);
  }
  *out = v;
  return PS_OK;
}

// one flag per (device, kernel variant)
template <int N>
struct SmemOnce {
  std::atomic<bool> done[PS_MAX_DEVICES][N];
  template <typename Fn>
  int ensure(int dev, int variant, Fn* fn, int bytes, const char* where) {
    std::atomic<bool>& f = done[dev][variant];
    if (f.load(std::memory_order_acquire)) return PS_OK;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { set_cuda_error(e, where); return PS_ERR_CUDA; }
    f.store(true, std::memory_order_release);
    return PS_OK;
  }
};

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// The separator forward is 170-650 short kernels in stream order (replayed from a CUDA graph); at batch 1 a kernel's launch
// latency, barrier / TMEM setup and first loads were longer than its work.  Kernels on the hot path (GEMMs, depthwise conv,
// statistics merge, overlap-add) are launched with cudaLaunchAttributeProgrammaticStreamSerialization and
//   * call pdl_trigger() at entry: the NEXT kernel of the stream may be scheduled on SMs that are idle or become free,
//   * call pdl_wait() after their on-chip prologue and BEFORE their first access to global memory that another kernel of
//     the stream writes or reads: it returns when every earlier kernel has completed and flushed (griddepcontrol.wait is
//     transitive through the chain), so from there on plain stream order holds - only the prologue overlaps.
// A kernel launched this way after one that never triggers simply starts when that one ends.  Stream capture records the
// programmatic edges in the graph.  MEASURED (round 2, run 14, CUDA-graph replay): no gain - cfg1 2.02 -> 2.12 ms, cfg4 6.13 ->
// 6.07 ms, cfg2 37.7 -> 37.8 ms: graph replay already issues the nodes back to back and the kernels' cost at batch 1 is
// inside them (weight stream per tile), not between them.  So the attribute is OFF by default; PS_PDL=1 turns it on.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct EnvInt;
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// integer switch read once from the environment (A/B runs only; product defaults never depend on it being set)
struct EnvInt {
  std::atomic<int> v{INT32_MIN};
  int get(const char* name, int dflt) {
    int x = v.load(std::memory_order_relaxed);
    if (x == INT32_MIN) {
      const char* e = getenv(name);
      x = e ? atoi(e) : dflt;
      v.store(x, std::memory_order_relaxed);
    }
    return x;
  }
};

// EXPERIMENT (PS_ORDER_ALT=1): consecutive big launches of a stream walk their tensors in alternating directions, so that a
// kernel starts with the part of its input the previous kernel wrote LAST - the part most likely still in the 126 MB L2
// (a [64, 3999, 512] tensor is 524 MB).  One process-wide toggle, flipped by every launch that honours it.
bool order_reversed();

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

// ---- gLN / gGN statistics merge of ONE batch item by a group of 256 threads (tid in [0,256)) -----------------
// The (count, mean, M2) partials are turned into raw moments and summed in fp64 in a fixed order,
//   S0 = sum n_i,  S1 = sum n_i mean_i,  S2 = sum (M2_i + n_i mean_i^2),   mean = S1 / S0,  var = (S2 - S1^2 / S0) / S0,
// -> folded per-channel affine  scale[c] = gamma[c] * rstd,  shift[c] = beta[c] - mean * rstd * gamma[c].
// (fp64 additions of fp32 inputs: the cancellation in var leaves 53 - log2(1 + mean^2 / var) bits, far beyond the fp32
// result.  The first version Chan-merged pairwise with two fp64 DIVISIONS per merge through an 8-level barrier tree: 5.7 us
// per launch, 20 % of a batch-1 Conv-TasNet forward - run 31.)
// BAR = barrier id the 256 threads share (0 = the whole 256-thread CTA).  Partials are read with ld.cg: they may have
// been written by other CTAs of the same launch (fused finalize), so a stale L1 line must not be used.
// Used by stats_finalize_kernel and, fused, by the CTA that completes an item's last tile in the producer kernels.
template <int BAR>
__device__ __forceinline__ void stats_finalize_item(const float* p, int64_t slots, const float* gamma, const float* beta, float eps,
                                                    int64_t C, float* scale_b, float* shift_b, float* meanvar_b, int tid,
                                                    double* sn, double* sm, double* s2) {
  double n = 0.0, a1 = 0.0, a2 = 0.0;
  for (int64_t i = tid; i < slots; i += 256) {
    const double bn = __ldcg(p + i * 3), bm = __ldcg(p + i * 3 + 1), b2 = __ldcg(p + i * 3 + 2);
    if (bn > 0.0) {
      const double t = bn * bm;
      n += bn;
      a1 += t;
      a2 += b2 + t * bm;
    }
  }
  // xor butterfly: both lanes of a pair add the same two values, so every lane ends with the same bits
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n += __shfl_xor_sync(0xffffffffu, n, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((tid & 31) == 0) { sn[tid >> 5] = n; sm[tid >> 5] = a1; s2[tid >> 5] = a2; }
  asm volatile("bar.sync %0, 256;" ::"n"(BAR) : "memory");
  double cnt = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { cnt += sn[w]; t1 += sm[w]; t2 += s2[w]; }  // fixed order, every thread the same
  const double mu = cnt > 0.0 ? t1 / cnt : 0.0;
  double var = cnt > 0.0 ? (t2 - t1 * mu) / cnt : 0.0;
  var = var > 0.0 ? var : 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float muf = (float)mu;
  for (int64_t c = tid; c < C; c += 256) {
    const float g = gamma ? gamma[c] : 1.f;
    const float bt = beta ? beta[c] : 0.f;
    const float sc = g * rstd;
    scale_b[c] = sc;
    shift_b[c] = fmaf(-muf, sc, bt);
  }
  if (meanvar_b && tid == 0) {
    meanvar_b[0] = muf;
    meanvar_b[1] = (float)var;
  }
  asm volatile("bar.sync %0, 256;" ::"n"(BAR) : "memory");  // scratch may be reused
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
