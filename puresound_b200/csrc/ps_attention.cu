// Multi-head self-attention core of nn.MultiheadAttention (reference lobe/attention.py:37-113 -> DPARNblock2D's
// intra-chunk attention over the frequency rows of a frame, dparn.py:12-108):
//
//     out[b, t, h*dh : (h+1)*dh] = sum_j softmax_j( q_t . k_j / sqrt(dh) ) v_j        (optionally only j <= t)
//
// qkv [B, L, 3E] is the fused in-projection (q | k | v, each E = heads*dh wide); the in/out projections are ps_gemm.
// One CTA per (sequence, head): K and V of the head sit in shared memory (L*dh floats each; every lane reads the same
// k_j / v_j: broadcast, conflict-free), a thread owns one query row: pass 1 computes its L scores (kept in shared memory,
// one column per thread) and their maximum, pass 2 the exponentials, the normaliser and the weighted sum of V with the dh
// accumulators in registers.  Exact fp32 (exp2f on log2-domain scores).  Sequences here are short (64 frequency rows, dh = 16).
#include "ps_common.cuh"

namespace ps {

// 16-byte shared-memory reads of a key / value row (all lanes of a warp read the same row: one broadcast wavefront per
// float4 instead of four).
// Packed fp32 FMA of sm_100 (FFMA2, fma.rn.f32x2): the per-key work is 2 x DH FMAs in ~50 issue slots and the kernel is
// issue-bound (run 88), so both loops run two lanes per instruction: the dot product as an (even, odd) pair of partial sums
// (a dependent chain of DH / 2 instead of DH), the weighted sum of V as DH / 2 accumulator pairs.
__device__ __forceinline__ unsigned long long att_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void att_unpack2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void att_ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

template <int DH>
__device__ __forceinline__ float att_dot(const float (&q)[DH], const float* __restrict__ k) {
  unsigned long long s2 = att_pack2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < DH; c += 4) {
    const float4 k4 = *reinterpret_cast<const float4*>(k + c);
    att_ffma2(s2, att_pack2(q[c], q[c + 1]), att_pack2(k4.x, k4.y));
    att_ffma2(s2, att_pack2(q[c + 2], q[c + 3]), att_pack2(k4.z, k4.w));
  }
  float a, b;
  att_unpack2(s2, a, b);
  return a + b;
}
template <int DH>
__device__ __forceinline__ void att_axpy(float (&acc)[DH], float p, const float* __restrict__ v) {
  const unsigned long long pp = att_pack2(p, p);
#pragma unroll
  for (int c = 0; c < DH; c += 4) {
    const float4 v4 = *reinterpret_cast<const float4*>(v + c);
    unsigned long long a0 = att_pack2(acc[c], acc[c + 1]), a1 = att_pack2(acc[c + 2], acc[c + 3]);
    att_ffma2(a0, pp, att_pack2(v4.x, v4.y));
    att_ffma2(a1, pp, att_pack2(v4.z, v4.w));
    att_unpack2(a0, acc[c], acc[c + 1]);
    att_unpack2(a1, acc[c + 2], acc[c + 3]);
  }
}

template <int DH>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int64_t L, int E,
                                                        int causal, float scale_log2e) {
  extern __shared__ __align__(16) float kv[];  // K [L][DH] | V [L][DH] | scores [L][blockDim.x]
  float* ks = kv;
  float* vs = kv + L * DH;
  float* sc = vs + L * DH + threadIdx.x;  // this thread's score of key j at sc[j * blockDim.x]: conflict-free
  const int nt = blockDim.x;
  const int64_t b = blockIdx.x;
  const int h = blockIdx.y;
  const float* base = qkv + b * L * 3 * E + h * DH;
  for (int64_t i = threadIdx.x; i < L * DH; i += nt) {
    const int64_t j = i / DH;
    const int c = (int)(i % DH);
    ks[i] = __ldg(base + j * 3 * E + E + c);
    vs[i] = __ldg(base + j * 3 * E + 2 * E + c);
  }
  __syncthreads();
  for (int64_t t = threadIdx.x; t < L; t += nt) {
    // pass 1: scores in the log2 domain (q pre-scaled by log2(e) / sqrt(dh)) and their maximum
    float q[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) q[c] = __ldg(base + t * 3 * E + c) * scale_log2e;
    const int64_t jend = causal ? t + 1 : L;
    float m = -INFINITY;
    for (int64_t j = 0; j < jend; ++j) {
      const float s = att_dot<DH>(q, ks + j * DH);
      sc[j * nt] = s;
      m = fmaxf(m, s);
    }
    // pass 2: p = 2^(s - m), normaliser and weighted sum of V
    float acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
    float l = 0.f;
    for (int64_t j = 0; j < jend; ++j) {
      const float p = exp2f(sc[j * nt] - m);
      l += p;
      att_axpy<DH>(acc, p, vs + j * DH);
    }
    const float inv = 1.f / l;
    float* o = out + (b * L + t) * E + h * DH;
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] = acc[c] * inv;
  }
}

// One CTA per SEQUENCE, all heads: the [L, 3E] tile of the fused projection is contiguous in memory, so it is staged with
// fully coalesced loads (the per-head kernel above fetches 64-byte slices 1.5 KB apart and launches heads x more CTAs);
// a thread then owns (head, query) items with the online max-rescaled softmax (no score buffer: the tile is the shared
// memory budget).  Used whenever the tile fits.
template <int DH>
__global__ void __launch_bounds__(256) attention_seq_kernel(const float* __restrict__ qkv, float* __restrict__ out, int L, int E,
                                                            int heads, int causal, float scale_log2e) {
  extern __shared__ __align__(16) float tile[];  // [L][3E]
  const int64_t b = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(qkv + b * (int64_t)L * 3 * E);
  const int n4 = L * 3 * E / 4;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<float4*>(tile)[i] = __ldg(src + i);
  __syncthreads();
  const int E3 = 3 * E;
  for (int item = threadIdx.x; item < heads * L; item += blockDim.x) {
    const int h = item / L, t = item % L;
    const float* kb = tile + E + h * DH;
    const float* vb = tile + 2 * E + h * DH;
    float q[DH], acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) { q[c] = tile[t * E3 + h * DH + c] * scale_log2e; acc[c] = 0.f; }
    float m = -INFINITY, l = 0.f;
    const int jend = causal ? t + 1 : L;
    for (int j = 0; j < jend; ++j) {
      const float s = att_dot<DH>(q, kb + j * E3);
      if (s > m) {  // rare after the first few keys: rescale the running sums
        const float corr = exp2f(m - s);
        l *= corr;
#pragma unroll
        for (int c = 0; c < DH; ++c) acc[c] *= corr;
        m = s;
      }
      const float p = exp2f(s - m);
      l += p;
      att_axpy<DH>(acc, p, vb + j * E3);
    }
    const float inv = 1.f / l;
    float* o = out + (b * L + t) * E + h * DH;
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] = acc[c] * inv;
  }
}

template <int DH>
static int attention_seq_launch(const float* qkv, float* out, int64_t batch, int64_t L, int64_t E, int heads, int causal, float scale,
                                size_t smem, cudaStream_t s) {
  static SmemOnce<1> once;  // per instantiation and device
  int dev = 0;
  if (int rc = current_device(&dev)) return rc;
  if (int rc = once.ensure(dev, 0, attention_seq_kernel<DH>, 200 * 1024, "cudaFuncSetAttribute(attention_seq_kernel)")) return rc;
  attention_seq_kernel<DH><<<(unsigned)batch, 256, smem, s>>>(qkv, out, (int)L, (int)E, heads, causal, scale);
  return PS_OK;
}

}  // namespace ps

extern "C" int ps_attention(const float* qkv, float* out, int64_t batch, int64_t L, int64_t E, int32_t heads, int32_t causal,
                            void* stream) {
  PS_REQUIRE(qkv && out && batch > 0 && L > 0 && E > 0 && heads > 0 && E % heads == 0);
  const int dh = (int)(E / heads);
  if (batch > 2147483647LL || heads > 65535) return PS_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  // whole-sequence kernel when the [L, 3E] tile fits in shared memory (and is 16-byte friendly)
  const size_t tile_bytes = (size_t)L * 3 * E * sizeof(float);
  if (tile_bytes <= 200 * 1024 && (E % 4) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0) {
    const float sc2 = 1.4426950408889634f / sqrtf((float)dh);
    int rc = PS_ERR_UNSUPPORTED;
    switch (dh) {
      case 4: rc = ps::attention_seq_launch<4>(qkv, out, batch, L, E, heads, causal, sc2, tile_bytes, s); break;
      case 8: rc = ps::attention_seq_launch<8>(qkv, out, batch, L, E, heads, causal, sc2, tile_bytes, s); break;
      case 16: rc = ps::attention_seq_launch<16>(qkv, out, batch, L, E, heads, causal, sc2, tile_bytes, s); break;
      case 32: rc = ps::attention_seq_launch<32>(qkv, out, batch, L, E, heads, causal, sc2, tile_bytes, s); break;
      case 64: rc = ps::attention_seq_launch<64>(qkv, out, batch, L, E, heads, causal, sc2, tile_bytes, s); break;
      default: break;
    }
    if (rc != PS_OK) return rc;
    PS_CHECK_LAUNCH("attention_seq_kernel");
    return PS_OK;
  }
  const int threads = L >= 128 ? 128 : (int)((L + 31) / 32 * 32);
  const size_t smem = ((size_t)2 * L * dh + (size_t)L * threads) * sizeof(float);
  constexpr size_t kMaxSmem = 200 * 1024;            // long sequences would need a key-blocked variant
  if (smem > kMaxSmem) return PS_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {  // opt in to large dynamic shared memory (idempotent, cheap)
    cudaError_t e = cudaSuccess;
    switch (dh) {
      case 4: e = cudaFuncSetAttribute(ps::attention_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem); break;
      case 8: e = cudaFuncSetAttribute(ps::attention_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem); break;
      case 16: e = cudaFuncSetAttribute(ps::attention_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem); break;
      case 32: e = cudaFuncSetAttribute(ps::attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem); break;
      case 64: e = cudaFuncSetAttribute(ps::attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem); break;
      default: break;
    }
    if (e != cudaSuccess) { ps::set_cuda_error(e, "cudaFuncSetAttribute(attention_kernel)"); return PS_ERR_CUDA; }
  }
  const float scale = 1.4426950408889634f / sqrtf((float)dh);  // scores in the log2 domain: softmax through exp2
  dim3 grid((unsigned)batch, (unsigned)heads);
  switch (dh) {
    case 4: ps::attention_kernel<4><<<grid, threads, smem, s>>>(qkv, out, L, (int)E, causal, scale); break;
    case 8: ps::attention_kernel<8><<<grid, threads, smem, s>>>(qkv, out, L, (int)E, causal, scale); break;
    case 16: ps::attention_kernel<16><<<grid, threads, smem, s>>>(qkv, out, L, (int)E, causal, scale); break;
    case 32: ps::attention_kernel<32><<<grid, threads, smem, s>>>(qkv, out, L, (int)E, causal, scale); break;
    case 64: ps::attention_kernel<64><<<grid, threads, smem, s>>>(qkv, out, L, (int)E, causal, scale); break;
    default: return PS_ERR_UNSUPPORTED;
  }
  PS_CHECK_LAUNCH("attention_kernel");
  return PS_OK;
}
