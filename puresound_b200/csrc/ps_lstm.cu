// LSTM recurrence (gate order i,f,g,o as torch.nn.LSTM), exact fp32.
//
// The input projections W_ih x + b_ih + b_hh for every position are produced by
// ps_gemm; this kernel runs the sequential part.  One CTA owns a group of BS
// sequences for all L steps: the hidden state lives in shared memory
// (double-buffered, [k][seq] so the matvec reads it as broadcast float4), the cell
// state in registers.  A thread owns one hidden unit for SPT sequences and
// computes all four gates of that unit, so the cell update is thread-local and a
// step needs a single __syncthreads().  W_hh is read k-major ([H][4H], coalesced
// over units) from L1/L2 every step — it is 256 KB at H=128, resident in the
// 126 MB L2 and shared by every CTA.
//
// The strided position function lets the same kernel run the intra-chunk pass
// ([N*S] sequences over K) and the inter-chunk pass ([N*K] sequences over S) of
// DPRNN on one [N,S,K,*] tensor, i.e. the permutes of dprnn.py:165-178 are folded
// into addressing.
#include "ps_common.cuh"

namespace ps {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// SPT: sequences per thread (8 or 16).  kPacked: W_hh comes as the gate-minor image [D][k][unit][4 gates] built by
// ps_lstm_pack_weights for the sizes the tensor-core kernel does not serve (SkiM's H = 256): ONE 16-byte load per k and
// thread instead of four 4-byte loads H apart, issued two chunks (16 k) ahead of the FMAs that use them.  The first
// version of this kernel (4 x LDG.32 per k, 16 loads in flight per thread) ran a step of H = 256 in ~45 us = the L2 latency
// times 64 dependent rounds; W_hh is 1 MB per direction there and is streamed from L2 by every CTA at every step, so the
// kernel needs ~64 KB in flight per SM to reach the L2 rate.
template <int SPT, bool kPacked>
__global__ void __launch_bounds__(256, 1) lstm_kernel(const ps_lstm_t d, const int BG) {
  extern __shared__ __align__(16) float hs[];  // [2][H][BS] hidden state, then [BS] int64 position bases
  const int H = (int)d.H;
  const int BS = BG * SPT;
  const int u = threadIdx.x % H;
  const int g = threadIdx.x / H;
  const int dir = blockIdx.y;
  const int64_t q0 = (int64_t)blockIdx.x * BS + (int64_t)g * SPT;
  const int64_t G = (int64_t)d.D * 4 * H;   // gx row width
  const int64_t OW = (int64_t)d.D * H;      // out row width
  const float* __restrict__ W = (kPacked ? reinterpret_cast<const float*>(d.w_packed) : d.w_hh_t) + (int64_t)dir * H * 4 * H;
  const bool gxi = d.gx_interleaved != 0;   // gx rows [dir][unit][gate]: the four gates of a unit are one 16-byte load
  int64_t* base_s = reinterpret_cast<int64_t*>(hs + (size_t)2 * H * BS);

  float c[SPT];
  unsigned valid = 0;
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int64_t q = q0 + i;
    const bool ok = q < d.n_seq;
    valid |= ok ? (1u << i) : 0u;
    if (u == 0) base_s[g * SPT + i] = ok ? (q / d.inner) * d.outer_stride + (q % d.inner) * d.inner_stride : 0;
    const int64_t so = ((int64_t)dir * d.n_seq + q) * H + u;
    c[i] = (ok && d.c0) ? d.c0[so] : 0.f;
    hs[(0 * H + u) * BS + g * SPT + i] = (ok && d.h0) ? d.h0[so] : 0.f;
  }
  __syncthreads();

  constexpr int UK = 8;  // k per chunk of the weight pipeline
  for (int64_t step = 0; step < d.L; ++step) {
    const int64_t t = dir ? d.L - 1 - step : step;
    const int cur = (int)(step & 1);
    float4 wq[2][UK];  // packed path: weights of the next two chunks, in flight while the current chunk is used
    if constexpr (kPacked) {
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int j = 0; j < UK; ++j) {
          const int k = b * UK + j;
          wq[b][j] = k < H ? __ldg(reinterpret_cast<const float4*>(W) + (int64_t)k * H + u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float acc[4][SPT];
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      if (valid & (1u << i)) {
        const float* gp = d.gx + (base_s[g * SPT + i] + t * d.step_stride) * G + (int64_t)dir * 4 * H;
        if (gxi) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp + 4 * u));
          acc[0][i] = g4.x; acc[1][i] = g4.y; acc[2][i] = g4.z; acc[3][i] = g4.w;
        } else {
#pragma unroll
          for (int gt = 0; gt < 4; ++gt) acc[gt][i] = __ldg(gp + gt * H + u);
        }
      } else {
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) acc[gt][i] = 0.f;
      }
    }
    const float* hcur = hs + (int64_t)cur * H * BS + g * SPT;
    auto fma_k = [&](int k, float w0, float w1, float w2, float w3) {
      float hv[SPT];
#pragma unroll
      for (int v = 0; v < SPT / 4; ++v) {
        const float4 h4 = *reinterpret_cast<const float4*>(hcur + k * BS + 4 * v);
        hv[4 * v] = h4.x; hv[4 * v + 1] = h4.y; hv[4 * v + 2] = h4.z; hv[4 * v + 3] = h4.w;
      }
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        acc[0][i] = fmaf(w0, hv[i], acc[0][i]);
        acc[1][i] = fmaf(w1, hv[i], acc[1][i]);
        acc[2][i] = fmaf(w2, hv[i], acc[2][i]);
        acc[3][i] = fmaf(w3, hv[i], acc[3][i]);
      }
    };
    if constexpr (kPacked) {
#pragma unroll 1
      for (int k0 = 0; k0 < H; k0 += 2 * UK) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float4 wc[UK];
#pragma unroll
          for (int j = 0; j < UK; ++j) wc[j] = wq[b][j];
#pragma unroll
          for (int j = 0; j < UK; ++j) {  // refill this buffer with the chunk two ahead
            const int k = k0 + (b + 2) * UK + j;
            wq[b][j] = k < H ? __ldg(reinterpret_cast<const float4*>(W) + (int64_t)k * H + u) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < UK; ++j) {
            const int k = k0 + b * UK + j;
            if (k < H) fma_k(k, wc[j].x, wc[j].y, wc[j].z, wc[j].w);
          }
        }
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float* wr = W + (int64_t)k * 4 * H + u;
        fma_k(k, __ldg(wr), __ldg(wr + H), __ldg(wr + 2 * H), __ldg(wr + 3 * H));
      }
    }
    float* hnext = hs + (int64_t)(cur ^ 1) * H * BS + u * BS + g * SPT;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const float ig = sigmoidf_(acc[0][i]);
      const float fg = sigmoidf_(acc[1][i]);
      const float gg = tanhf(acc[2][i]);
      const float og = sigmoidf_(acc[3][i]);
      c[i] = fmaf(fg, c[i], ig * gg);
      const float h = og * tanhf(c[i]);
      hnext[i] = h;
      if (valid & (1u << i)) d.out[(base_s[g * SPT + i] + t * d.step_stride) * OW + (int64_t)dir * H + u] = h;
    }
    __syncthreads();
  }
  const float* hfin = hs + (int64_t)(d.L & 1) * H * BS + u * BS + g * SPT;  // the buffer the last step wrote
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    if (!(valid & (1u << i))) continue;
    const int64_t so = ((int64_t)dir * d.n_seq + q0 + i) * H + u;
    if (d.hn) d.hn[so] = hfin[i];
    if (d.cn) d.cn[so] = c[i];
  }
}

// W_hh^T [D][k][4 gates][unit] -> gate-minor [D][k][unit][4 gates]
__global__ void lstm_pack_simt_kernel(const float* __restrict__ w, int64_t H, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t gt = i & 3, uu = (i >> 2) % H, dk = (i >> 2) / H;  // dk = dir * H + k
  out[i] = w[dk * 4 * H + gt * H + uu];
}

int lstm_simt_pack(const float* w_hh_t, int64_t H, int32_t D, void* packed, cudaStream_t s) {
  const int64_t n = (int64_t)D * H * 4 * H;
  lstm_pack_simt_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(w_hh_t, H, n, reinterpret_cast<float*>(packed));
  PS_CHECK_LAUNCH("lstm_pack_simt_kernel");
  return PS_OK;
}

bool lstm_tc_eligible(const ps_lstm_t& d);
int lstm_tc_launch(const ps_lstm_t& d, cudaStream_t s);

}  // namespace ps

// true when ps_lstm_pack_weights builds the CUDA-core kernel's gate-minor fp32 image for this H (defined in ps_lstm_tc.cu)
bool ps_lstm_packed_is_simt(int64_t H);

extern "C" int ps_lstm(const ps_lstm_t* dp, void* stream) {
  PS_REQUIRE(dp != nullptr);
  const ps_lstm_t& d = *dp;
  PS_REQUIRE(d.gx && d.w_hh_t && d.out && d.n_seq > 0 && d.L > 0 && d.H > 0 && (d.D == 1 || d.D == 2));
  PS_REQUIRE(d.inner > 0 && (d.h0 == nullptr) == (d.c0 == nullptr));
  if (ps::lstm_tc_eligible(d)) return ps::lstm_tc_launch(d, (cudaStream_t)stream);
  if (d.H > 256) return PS_ERR_UNSUPPORTED;  // one thread per hidden unit, <= 256 threads per CTA
  cudaStream_t s = (cudaStream_t)stream;
  int BG = (int)(256 / d.H);
  if (BG < 1) BG = 1;
  // the gate-minor image exists only for the sizes the tensor-core kernel does not serve (for the others w_packed is that
  // kernel's image, which this one must not read)
  const bool packed = d.w_packed != nullptr && ps_lstm_packed_is_simt(d.H) && (reinterpret_cast<uintptr_t>(d.w_packed) & 15) == 0;
  if (d.gx_interleaved) PS_REQUIRE((reinterpret_cast<uintptr_t>(d.gx) & 15) == 0);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  // 16 sequences per thread when 8 would need more than one wave of CTAs (SkiM segments: 2144 sequences -> 134 CTAs);
  // 4 when even that leaves most SMs idle (SkiM's memory LSTMs: 32 sequences x 67 steps - a step is then 1024 k-FMA
  // rounds per warp instead of 2048, on twice as many SMs)
  const bool wide = packed && ps::cdiv(d.n_seq, (int64_t)BG * 8) * d.D > sms;
  const bool narrow = packed && !wide && ps::cdiv(d.n_seq, (int64_t)BG * 4) * d.D <= sms;
  const int SPT = wide ? 16 : (narrow ? 4 : 8);
  const int BS = BG * SPT;
  const int threads = (int)d.H * BG;
  const size_t smem = (size_t)2 * d.H * BS * sizeof(float) + (size_t)BS * sizeof(int64_t);
  const int64_t nblk = ps::cdiv(d.n_seq, BS);
  if (nblk > 2147483647LL) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)nblk, (unsigned)d.D);
  if (smem > 48 * 1024) {
    static bool attr[2] = {false, false};
    if (!attr[wide]) {
      cudaError_t e = wide ? cudaFuncSetAttribute(ps::lstm_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)
                           : cudaFuncSetAttribute(packed ? ps::lstm_kernel<8, true> : ps::lstm_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) { ps::set_cuda_error(e, "cudaFuncSetAttribute(lstm_kernel)"); return PS_ERR_CUDA; }
      attr[wide] = true;
    }
  }
  if (wide) ps::lstm_kernel<16, true><<<grid, threads, smem, s>>>(d, BG);
  else if (narrow) ps::lstm_kernel<4, true><<<grid, threads, smem, s>>>(d, BG);
  else if (packed) ps::lstm_kernel<8, true><<<grid, threads, smem, s>>>(d, BG);
  else ps::lstm_kernel<8, false><<<grid, threads, smem, s>>>(d, BG);
  PS_CHECK_LAUNCH("lstm_kernel");
  return PS_OK;
}
