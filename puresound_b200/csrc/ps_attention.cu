// Multi-head self-attention core of nn.MultiheadAttention (reference lobe/attention.py:37-113 -> DPARNblock2D's
// intra-chunk attention over the frequency rows of a frame, dparn.py:12-108):
//
//     out[b, t, h*dh : (h+1)*dh] = sum_j softmax_j( q_t . k_j / sqrt(dh) ) v_j        (optionally only j <= t)
//
// qkv [B, L, 3E] is the fused in-projection (q | k | v, each E = heads*dh wide); the in/out projections are ps_gemm.
// One CTA per (sequence, head): K and V of the head sit in shared memory (L*dh floats each; every lane reads the same
// k_j / v_j: broadcast, conflict-free), a thread owns one query row with its dh accumulators in registers and runs the
// online (single-pass, max-rescaled) softmax in exact fp32.  Sequences here are short (64 frequency rows, dh = 16).
#include "ps_common.cuh"

namespace ps {

template <int DH>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int64_t L, int E,
                                                        int causal, float scale) {
  extern __shared__ __align__(16) float kv[];  // K [L][DH] | V [L][DH]
  float* ks = kv;
  float* vs = kv + L * DH;
  const int64_t b = blockIdx.x;
  const int h = blockIdx.y;
  const float* base = qkv + b * L * 3 * E + h * DH;
  for (int64_t i = threadIdx.x; i < L * DH; i += blockDim.x) {
    const int64_t j = i / DH;
    const int c = (int)(i % DH);
    ks[i] = __ldg(base + j * 3 * E + E + c);
    vs[i] = __ldg(base + j * 3 * E + 2 * E + c);
  }
  __syncthreads();
  for (int64_t t = threadIdx.x; t < L; t += blockDim.x) {
    float q[DH], acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) { q[c] = __ldg(base + t * 3 * E + c) * scale; acc[c] = 0.f; }
    float m = -INFINITY, l = 0.f;
    const int64_t jend = causal ? t + 1 : L;
    for (int64_t j = 0; j < jend; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DH; ++c) s = fmaf(q[c], ks[j * DH + c], s);
      const float mn = fmaxf(m, s);
      const float corr = (m == -INFINITY) ? 0.f : expf(m - mn);
      const float p = expf(s - mn);
      l = fmaf(l, corr, p);
#pragma unroll
      for (int c = 0; c < DH; ++c) acc[c] = fmaf(acc[c], corr, p * vs[j * DH + c]);
      m = mn;
    }
    const float inv = 1.f / l;
    float* o = out + (b * L + t) * E + h * DH;
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] = acc[c] * inv;
  }
}

}  // namespace ps

extern "C" int ps_attention(const float* qkv, float* out, int64_t batch, int64_t L, int64_t E, int32_t heads, int32_t causal,
                            void* stream) {
  PS_REQUIRE(qkv && out && batch > 0 && L > 0 && E > 0 && heads > 0 && E % heads == 0);
  const int dh = (int)(E / heads);
  if (batch > 2147483647LL || heads > 65535) return PS_ERR_UNSUPPORTED;
  const size_t smem = (size_t)2 * L * dh * sizeof(float);
  if (smem > 48 * 1024) return PS_ERR_UNSUPPORTED;  // long sequences would need a key-blocked variant
  const float scale = 1.f / sqrtf((float)dh);
  dim3 grid((unsigned)batch, (unsigned)heads);
  const int threads = L >= 128 ? 128 : (int)((L + 31) / 32 * 32);
  cudaStream_t s = (cudaStream_t)stream;
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
