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
#include <stdlib.h>

#include "ps_common.cuh"

namespace ps {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// SPT: sequences per thread (4, 8 or 16).  kPacked: W_hh comes as the gate-minor image [D][k][unit][4 gates] built by
// ps_lstm_pack_weights for the sizes the tensor-core kernel does not serve (SkiM's H = 256).  W_hh is 1 MB per direction
// there and every CTA streams it from L2 at every step; a thread needs one 16-byte weight vector per k, and fetches it with
// cp.async into a PRIVATE shared-memory ring (no barrier: the thread that copies is the thread that reads) that runs
// 16 - 32 k ahead and straight across step boundaries, so the L2 latency never reaches the FMA chain and no registers
// hold weights in flight.  History (run 98/100, H = 256, 2144 sequences x 150 steps): 4 x LDG.32 per k, 16 loads in
// flight: ~45 us per step-wave pair; LDG.128 prefetched 16 k ahead through registers (254 registers, 2 warps per
// scheduler stalled on the h loads): 45 us per step at 16 sequences per thread.
constexpr int LSTM_CK = 4;     // k per cp.async group
// groups in flight: 8 (32 k, 128 KB at 256 threads) - or 4 in the fixed-size build with <= 8 sequences per thread, which is
// meant to run TWO CTAs per SM (2 x (16 KB state + 64 KB ring), <= 128 registers): four warps per scheduler instead of two
__host__ __device__ constexpr int lstm_nch(int SPT, int HC) { return (HC && SPT <= 8) ? 4 : 8; }
__host__ __device__ constexpr int lstm_minb(int SPT, int HC) { return (HC && SPT <= 8) ? 2 : 1; }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// Packed fp32 FMA of sm_100 (FFMA2): (d.lo, d.hi) += (a.lo, a.hi) * (b.lo, b.hi), each half rounded exactly like fmaf.
// The k loop of the recurrence is issue-bound (ncu run 102: issue slots 68 % busy, fp32 FMA pipe 54 %): two sequences per
// instruction halve its FMA instruction count.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// HC: compile-time hidden size (256: SkiM's recipes; one thread per unit, BG = 1) or 0 = run-time sizes.  With HC every
// shared-memory address of the k loop is an immediate offset from one induction register: the run-time-size build spent 48
// of its 112 instructions per k on address arithmetic and `k < H` guards (ncu, run 101), the fixed-size one is 64 FFMA + 5
// LDS + 1 cp.async.
template <int SPT, bool kPacked, int HC = 0>
__global__ void __launch_bounds__(256, lstm_minb(SPT, HC)) lstm_kernel(const ps_lstm_t d, const int BG) {
  constexpr int LSTM_NCH = lstm_nch(SPT, HC);
  extern __shared__ __align__(16) float hs[];  // [2][H][BS] hidden state | [BS] int64 position bases | weight ring (packed)
  const int H = HC ? HC : (int)d.H;
  const int BS = HC ? SPT : BG * SPT;
  const int u = HC ? (int)threadIdx.x : (int)(threadIdx.x % H);
  const int g = HC ? 0 : (int)(threadIdx.x / H);
  const int dir = blockIdx.y;
  const int nthr = HC ? HC : (int)blockDim.x;
  const int64_t q0 = (int64_t)blockIdx.x * BS + (int64_t)g * SPT;
  const int64_t G = (int64_t)d.D * 4 * H;   // gx row width
  const int64_t OW = (int64_t)d.D * H;      // out row width
  const float* __restrict__ W = (kPacked ? reinterpret_cast<const float*>(d.w_packed) : d.w_hh_t) + (int64_t)dir * H * 4 * H;
  const bool gxi = d.gx_interleaved != 0;   // gx rows [dir][unit][gate]: the four gates of a unit are one 16-byte load
  int64_t* base_s = reinterpret_cast<int64_t*>(hs + (size_t)2 * H * BS);
  float4* ring = reinterpret_cast<float4*>(base_s + ((BS + 1) & ~1)) + threadIdx.x;  // this thread's slots: ring[slot * nthr]

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

  // weight stream (packed): chunk n of the whole launch covers k = (n % KC) * CK .. + CK of step n / KC
  const int KC = (H + LSTM_CK - 1) / LSTM_CK;
  int64_t to_issue = (int64_t)d.L * KC;  // chunks not yet requested
  int issue_kc = 0, issue_slot = 0;      // k chunk / ring slot of the next request
  const float4* wsrc = reinterpret_cast<const float4*>(W) + u;
  auto issue_chunk = [&]() {
    if (to_issue > 0) {
#pragma unroll
      for (int j = 0; j < LSTM_CK; ++j) {
        const int k = issue_kc * LSTM_CK + j;
        if (HC || k < H) cp_async16(ring + (issue_slot * LSTM_CK + j) * nthr, wsrc + (int64_t)k * H);
      }
    }
    cp_async_commit();  // (an empty group past the end keeps the group count in step with the consumer)
    --to_issue;
    if (++issue_kc == KC) issue_kc = 0;
    if (++issue_slot == LSTM_NCH) issue_slot = 0;
  };
  if constexpr (kPacked) {
#pragma unroll 1
    for (int n = 0; n < LSTM_NCH; ++n) issue_chunk();
  }
  int use_slot = 0;  // ring slot of the next chunk to consume

  for (int64_t step = 0; step < d.L; ++step) {
    const int64_t t = dir ? d.L - 1 - step : step;
    const int cur = (int)(step & 1);
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
    unsigned long long acc2[4][SPT / 2];  // (sequence 2j, sequence 2j + 1) per gate
#pragma unroll
    for (int gt = 0; gt < 4; ++gt)
#pragma unroll
      for (int j = 0; j < SPT / 2; ++j) acc2[gt][j] = pack2(acc[gt][2 * j], acc[gt][2 * j + 1]);
    const float* hcur = hs + (int64_t)cur * H * BS + g * SPT;
    auto fma_k = [&](int k, float w0, float w1, float w2, float w3) {
      unsigned long long hv2[SPT / 2];
#pragma unroll
      for (int v = 0; v < SPT / 4; ++v) {
        const float4 h4 = *reinterpret_cast<const float4*>(hcur + k * BS + 4 * v);
        hv2[2 * v] = pack2(h4.x, h4.y);
        hv2[2 * v + 1] = pack2(h4.z, h4.w);
      }
      const unsigned long long ww[4] = {pack2(w0, w0), pack2(w1, w1), pack2(w2, w2), pack2(w3, w3)};
#pragma unroll
      for (int j = 0; j < SPT / 2; ++j) {
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) ffma2(acc2[gt][j], ww[gt], hv2[j]);
      }
    };
    if constexpr (kPacked) {
#pragma unroll 1
      for (int kc = 0; kc < KC; ++kc) {
        cp_async_wait<LSTM_NCH - 1>();  // the oldest group in flight (this chunk) has landed
        float4 wc[LSTM_CK];
#pragma unroll
        for (int j = 0; j < LSTM_CK; ++j) wc[j] = ring[(use_slot * LSTM_CK + j) * nthr];
        issue_chunk();  // refill the slot just read (same thread, program order)
        if (++use_slot == LSTM_NCH) use_slot = 0;
#pragma unroll
        for (int j = 0; j < LSTM_CK; ++j) {
          const int k = kc * LSTM_CK + j;
          if (HC || k < H) fma_k(k, wc[j].x, wc[j].y, wc[j].z, wc[j].w);
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
    for (int gt = 0; gt < 4; ++gt)
#pragma unroll
      for (int j = 0; j < SPT / 2; ++j) unpack2(acc2[gt][j], acc[gt][2 * j], acc[gt][2 * j + 1]);
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
  if constexpr (kPacked) cp_async_wait<0>();
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
  int dev = 0, sms = 0;
  if (int rc = ps::current_device(&dev)) return rc;
  if (int rc = ps::sm_count_of(dev, &sms)) return rc;
  // 16 sequences per thread when 8 would need more than one wave of CTAs (SkiM segments: 2144 sequences -> 134 CTAs);
  // 4 when even that leaves most SMs idle (SkiM's memory LSTMs: 32 sequences x 67 steps - a step is then 1024 k-FMA
  // rounds per warp instead of 2048, on twice as many SMs)
  const bool h256 = packed && d.H == 256;  // the fixed-size build (BG == 1)
  static ps::EnvInt spt16_env;  // PS_LSTM_SPT16=1: 16 sequences per thread, one CTA per SM (A/B against 8 per thread, two CTAs per SM)
  const int spt16 = spt16_env.get("PS_LSTM_SPT16", 0) == 1;
  const bool wide = spt16 && h256 && ps::cdiv(d.n_seq, (int64_t)BG * 8) * d.D > sms;
  const bool narrow = packed && !wide && ps::cdiv(d.n_seq, (int64_t)BG * 4) * d.D <= sms;
  const int SPT = wide ? 16 : (narrow ? 4 : 8);
  const int BS = BG * SPT;
  const int threads = (int)d.H * BG;
  const size_t smem = (size_t)2 * d.H * BS * sizeof(float) + (size_t)((BS + 1) & ~1) * sizeof(int64_t) +
                      (packed ? (size_t)ps::LSTM_CK * ps::lstm_nch(SPT, h256 ? 256 : 0) * threads * sizeof(float4) : 0);
  const int64_t nblk = ps::cdiv(d.n_seq, BS);
  if (nblk > 2147483647LL) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)nblk, (unsigned)d.D);
  if (packed) {
    static ps::SmemOnce<5> once;
    const int smax = 200 * 1024;
    const char* where = "cudaFuncSetAttribute(lstm_kernel)";
    int rc = PS_OK;
    if (wide) rc = once.ensure(dev, 0, ps::lstm_kernel<16, true, 256>, smax, where);
    else if (narrow) {
      rc = once.ensure(dev, 1, ps::lstm_kernel<4, true>, smax, where);
      if (rc == PS_OK) rc = once.ensure(dev, 2, ps::lstm_kernel<4, true, 256>, smax, where);
    } else {
      rc = once.ensure(dev, 3, ps::lstm_kernel<8, true>, smax, where);
      if (rc == PS_OK) rc = once.ensure(dev, 4, ps::lstm_kernel<8, true, 256>, smax, where);
    }
    if (rc != PS_OK) return rc;
  }
  if (wide) ps::lstm_kernel<16, true, 256><<<grid, threads, smem, s>>>(d, BG);
  else if (narrow && h256) ps::lstm_kernel<4, true, 256><<<grid, threads, smem, s>>>(d, BG);
  else if (narrow) ps::lstm_kernel<4, true><<<grid, threads, smem, s>>>(d, BG);
  else if (h256) ps::lstm_kernel<8, true, 256><<<grid, threads, smem, s>>>(d, BG);
  else if (packed) ps::lstm_kernel<8, true><<<grid, threads, smem, s>>>(d, BG);
  else ps::lstm_kernel<8, false><<<grid, threads, smem, s>>>(d, BG);
  PS_CHECK_LAUNCH("lstm_kernel");
  return PS_OK;
}
