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

constexpr int LSTM_SPT = 8;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void lstm_kernel(const ps_lstm_t d, const int BG) {
  extern __shared__ __align__(16) float hs[];  // [2][H][BS]
  constexpr int SPT = LSTM_SPT;
  const int H = (int)d.H;
  const int BS = BG * SPT;
  const int u = threadIdx.x % H;
  const int g = threadIdx.x / H;
  const int dir = blockIdx.y;
  const int64_t q0 = (int64_t)blockIdx.x * BS + (int64_t)g * SPT;
  const int64_t G = (int64_t)d.D * 4 * H;   // gx row width
  const int64_t OW = (int64_t)d.D * H;      // out row width
  const float* __restrict__ W = d.w_hh_t + (int64_t)dir * H * 4 * H;

  float c[SPT], hlast[SPT];
  int64_t base[SPT];
  bool valid[SPT];
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int64_t q = q0 + i;
    valid[i] = q < d.n_seq;
    base[i] = valid[i] ? (q / d.inner) * d.outer_stride + (q % d.inner) * d.inner_stride : 0;
    const int64_t so = ((int64_t)dir * d.n_seq + q) * H + u;
    c[i] = (valid[i] && d.c0) ? d.c0[so] : 0.f;
    hlast[i] = (valid[i] && d.h0) ? d.h0[so] : 0.f;
    hs[(0 * H + u) * BS + g * SPT + i] = hlast[i];
  }
  __syncthreads();

  for (int64_t step = 0; step < d.L; ++step) {
    const int64_t t = dir ? d.L - 1 - step : step;
    const int cur = (int)(step & 1);
    float acc[4][SPT];
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      if (valid[i]) {
        const float* gp = d.gx + (base[i] + t * d.step_stride) * G + (int64_t)dir * 4 * H + u;
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) acc[gt][i] = __ldg(gp + gt * H);
      } else {
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) acc[gt][i] = 0.f;
      }
    }
    const float* hcur = hs + (int64_t)cur * H * BS + g * SPT;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float* wr = W + (int64_t)k * 4 * H + u;
      const float w0 = __ldg(wr), w1 = __ldg(wr + H), w2 = __ldg(wr + 2 * H), w3 = __ldg(wr + 3 * H);
      const float4 ha = *reinterpret_cast<const float4*>(hcur + k * BS);
      const float4 hb = *reinterpret_cast<const float4*>(hcur + k * BS + 4);
      const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        acc[0][i] = fmaf(w0, hv[i], acc[0][i]);
        acc[1][i] = fmaf(w1, hv[i], acc[1][i]);
        acc[2][i] = fmaf(w2, hv[i], acc[2][i]);
        acc[3][i] = fmaf(w3, hv[i], acc[3][i]);
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
      hlast[i] = h;
      hnext[i] = h;
      if (valid[i]) d.out[(base[i] + t * d.step_stride) * OW + (int64_t)dir * H + u] = h;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    if (!valid[i]) continue;
    const int64_t so = ((int64_t)dir * d.n_seq + q0 + i) * H + u;
    if (d.hn) d.hn[so] = hlast[i];
    if (d.cn) d.cn[so] = c[i];
  }
}

bool lstm_tc_eligible(const ps_lstm_t& d);
int lstm_tc_launch(const ps_lstm_t& d, cudaStream_t s);

}  // namespace ps

extern "C" int ps_lstm(const ps_lstm_t* dp, void* stream) {
  PS_REQUIRE(dp != nullptr);
  const ps_lstm_t& d = *dp;
  PS_REQUIRE(d.gx && d.w_hh_t && d.out && d.n_seq > 0 && d.L > 0 && d.H > 0 && (d.D == 1 || d.D == 2));
  PS_REQUIRE(d.inner > 0 && (d.h0 == nullptr) == (d.c0 == nullptr));
  if (ps::lstm_tc_eligible(d)) return ps::lstm_tc_launch(d, (cudaStream_t)stream);
  if (d.H > 256) return PS_ERR_UNSUPPORTED;  // 140 regs x (H*BG) threads must fit the register file
  int BG = (int)(256 / d.H);
  if (BG < 1) BG = 1;
  const int BS = BG * ps::LSTM_SPT;
  const int threads = (int)d.H * BG;
  const size_t smem = (size_t)2 * d.H * BS * sizeof(float);
  const int64_t nblk = ps::cdiv(d.n_seq, BS);
  if (nblk > 2147483647LL) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)nblk, (unsigned)d.D);
  ps::lstm_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(d, BG);
  PS_CHECK_LAUNCH("lstm_kernel");
  return PS_OK;
}
