// Exact-fp32 CUDA-core GEMM with fused prologue/epilogue.
//
//   Y[b,r,m] = epi( sum_k pro(X[b,r,k]) * W[m,k] )
//
// This is the reference-accuracy back end (true fp32 FMA accumulation): it serves
// every shape the tcgen05 back end does not take (ragged K/M, framed encoder views
// with row stride < K, the iSTFT synthesis GEMM that must stay fp32 because of the
// 2.6e4x edge amplification of the window-sumsquare division, the per-hop streaming
// steps) and is the on-device cross-check for the tensor-core kernel.
//
// Two tile shapes, 256 threads (16 x 16) each:
//   128 rows x 128 channels x 16 k, 8x8 outputs per thread — throughput shape;
//    32 rows x  64 channels x 16 k, 2x4 outputs per thread — latency shape for skinny problems (a streaming hop is
//    256 x 512 x 512: the big tile would run on 8 of the 148 SMs and take 55 us; measured round 1 run 6).
// Register prefetch of the next k-slab while the current one is consumed from shared memory.  The prologue (norm
// affine + activation, cLN row-norm, or mask product; compile-time selected) is applied on the global->register leg
// so shared memory already holds the transformed operand; the epilogue adds bias / per-item bias / activation /
// residual and emits one Welford partial (count, mean, M2) per CTA (big tile only: the slot layout is 128 x 128).
#include <stdlib.h>

#include "ps_common.cuh"

namespace ps {

constexpr int BK = 16, NT = 256;

__device__ __forceinline__ unsigned long long simt_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void simt_unpack2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void simt_ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

struct XLoadCtx {
  float mean, rstd, slope;
};

// PRO is a template parameter: dispatching on d.pro_mode per loaded element costs an indirect branch each time
template <int PRO>
__device__ __forceinline__ float pro_apply(const ps_gemm_t& d, float x, float x2, int64_t b, int64_t k, const XLoadCtx& c) {
  if constexpr (PRO == PS_PRO_AFFINE) {
    int64_t o = b * d.pro_batch_stride + k;
    return apply_act(fmaf(x, __ldg(d.pro_a + o), __ldg(d.pro_b + o)), d.pro_act, c.slope);
  } else if constexpr (PRO == PS_PRO_ROWNORM) {
    return apply_act(fmaf((x - c.mean) * c.rstd, __ldg(d.pro_a + k), __ldg(d.pro_b + k)), d.pro_act, c.slope);
  } else if constexpr (PRO == PS_PRO_MASK) {
    return x * apply_act(x2, d.pro_act, c.slope);
  } else {
    return x;
  }
}

// load PT consecutive floats (PT in {2,4,8}) with bounds k < K; vector path when aligned and fully in range
template <int PT>
__device__ __forceinline__ void load_run(const float* p, int64_t k, int64_t K, bool vec, float (&v)[PT]) {
#pragma unroll
  for (int i = 0; i < PT; ++i) v[i] = 0.f;
  if constexpr (PT >= 4) {
#pragma unroll
    for (int h = 0; h < PT / 4; ++h) {
      const int64_t kk = k + h * 4;
      if (vec && kk + 3 < K) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p + kk));
        v[h * 4] = t.x; v[h * 4 + 1] = t.y; v[h * 4 + 2] = t.z; v[h * 4 + 3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (kk + i < K) v[h * 4 + i] = __ldg(p + kk + i);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < PT; ++i)
      if (k + i < K) v[i] = __ldg(p + k + i);
  }
}

template <int PRO, int BR, int BC>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const ps_gemm_t d, const int x_vec, const int w_vec) {
  constexpr int TR = BR / 16;            // rows per thread
  constexpr int NG = BC / 64;            // groups of 4 output columns per thread (column = g*64 + tx*4 + j)
  constexpr int TCc = NG * 4;            // columns per thread
  constexpr int XPT = BR / 16;           // k-run a thread loads from the X tile  (BR*BK/NT)
  constexpr int WPT = BC / 16;           // k-run a thread loads from the W tile
  constexpr int LDX = BR + 4, LDW = BC + 4;
  __shared__ __align__(16) float Xs[BK][LDX];
  __shared__ __align__(16) float Ws[BK][LDW];
  __shared__ Wf red[NT / 32];

  const int tid = threadIdx.x;
  const int64_t b = blockIdx.z;
  const int64_t row0 = (int64_t)blockIdx.y * BR;
  const int64_t m0 = (int64_t)blockIdx.x * BC;

  // loader mapping: (16 / PT) threads per tile row, PT consecutive k each
  const int xl_row = tid / (16 / XPT), xl_k0 = (tid % (16 / XPT)) * XPT;
  const int wl_row = tid / (16 / WPT), wl_k0 = (tid % (16 / WPT)) * WPT;
  const int64_t xr = row0 + xl_row;
  const int64_t wm = m0 + wl_row;
  const bool xr_ok = xr < d.rows;
  const bool wm_ok = wm < d.M;
  const float* xp = d.X + b * d.x_batch_stride + xr * d.x_row_stride;
  const float* x2p = d.X2 ? d.X2 + b * d.x_batch_stride + xr * d.x_row_stride : nullptr;
  const float* wp = d.W + wm * d.w_row_stride;

  XLoadCtx ctx;
  ctx.mean = 0.f; ctx.rstd = 1.f;
  ctx.slope = (d.pro_slope != nullptr) ? __ldg(d.pro_slope) : 0.f;
  if (PRO == PS_PRO_ROWNORM && xr_ok) {
    const float* rs = d.pro_rowstats + (b * d.rows + xr) * 2;
    ctx.mean = __ldg(rs);
    ctx.rstd = __ldg(rs + 1);
  }

  float xreg[XPT], wreg[WPT];
  auto load_g = [&](int64_t k0) {
    float xv[XPT], x2v[XPT];
#pragma unroll
    for (int i = 0; i < XPT; ++i) { xv[i] = 0.f; x2v[i] = 0.f; }
    if (xr_ok) {
      load_run<XPT>(xp, k0 + xl_k0, d.K, x_vec, xv);
      if (PRO == PS_PRO_MASK) load_run<XPT>(x2p, k0 + xl_k0, d.K, x_vec, x2v);
    }
#pragma unroll
    for (int i = 0; i < XPT; ++i) {
      const int64_t k = k0 + xl_k0 + i;
      xreg[i] = (xr_ok && k < d.K) ? pro_apply<PRO>(d, xv[i], x2v[i], b, k, ctx) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < WPT; ++i) wreg[i] = 0.f;
    if (wm_ok) load_run<WPT>(wp, k0 + wl_k0, d.K, w_vec, wreg);
  };
  auto store_s = [&]() {
#pragma unroll
    for (int i = 0; i < XPT; ++i) Xs[xl_k0 + i][xl_row] = xreg[i];
#pragma unroll
    for (int i = 0; i < WPT; ++i) Ws[wl_k0 + i][wl_row] = wreg[i];
  };

  const int ty = tid >> 4, tx = tid & 15;
  static_assert(TCc % 2 == 0, "column pairs");
  unsigned long long acc2[TR][TCc / 2];  // (column j, column j + 1) pairs
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < TCc / 2; ++j) acc2[i][j] = simt_pack2(0.f, 0.f);

  const int64_t nkt = (d.K + BK - 1) / BK;
  load_g(0);
  store_s();
  __syncthreads();
  for (int64_t kt = 0; kt < nkt; ++kt) {
    if (kt + 1 < nkt) load_g((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TR], w[TCc];
      if constexpr (TR % 4 == 0) {
#pragma unroll
        for (int h = 0; h < TR / 4; ++h) {
          const float4 t = *reinterpret_cast<const float4*>(&Xs[k][ty * TR + h * 4]);
          a[h * 4] = t.x; a[h * 4 + 1] = t.y; a[h * 4 + 2] = t.z; a[h * 4 + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TR; ++i) a[i] = Xs[k][ty * TR + i];
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(&Ws[k][g * 64 + tx * 4]);
        w[g * 4] = t.x; w[g * 4 + 1] = t.y; w[g * 4 + 2] = t.z; w[g * 4 + 3] = t.w;
      }
      // packed fp32 FMA (sm_100 FFMA2): two adjacent output columns per instruction, the row operand as a broadcast
      // scalar; each half rounds like fmaf, so results are unchanged
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const unsigned long long aa = simt_pack2(a[i], a[i]);
#pragma unroll
        for (int j = 0; j < TCc; j += 2) simt_ffma2(acc2[i][j / 2], aa, simt_pack2(w[j], w[j + 1]));
      }
    }
    __syncthreads();
    if (kt + 1 < nkt) {
      store_s();
      __syncthreads();
    }
  }

  // ---- epilogue ----
  float acc[TR][TCc];
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < TCc / 2; ++j) simt_unpack2(acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);
  const float eslope = (d.epi_slope != nullptr) ? __ldg(d.epi_slope) : 0.f;
  WfAcc st;
  st.init();
  float bj[TCc];
#pragma unroll
  for (int j = 0; j < TCc; ++j) {
    const int64_t m = m0 + (j / 4) * 64 + tx * 4 + (j % 4);
    float v = 0.f;
    if (m < d.M) {
      if (d.bias) v += __ldg(d.bias + m);
      if (d.bias_batch) v += __ldg(d.bias_batch + b * d.M + m);
    }
    bj[j] = v;
  }
  const bool y_vec = ((d.y_row_stride & 3) == 0) && ((d.y_batch_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.Y) & 15) == 0);
  const bool r_vec = d.residual && ((d.res_row_stride & 3) == 0) && ((d.res_batch_stride & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(d.residual) & 15) == 0);
#pragma unroll
  for (int i = 0; i < TR; ++i) {
    const int64_t r = row0 + ty * TR + i;
    if (r >= d.rows) continue;
    float* yp = d.Y + b * d.y_batch_stride + r * d.y_row_stride;
    const float* rp = d.residual ? d.residual + b * d.res_batch_stride + r * d.res_row_stride : nullptr;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int64_t mb = m0 + g * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = apply_act(acc[i][g * 4 + j] + bj[g * 4 + j], d.epi_act, eslope);
      if (mb + 3 < d.M) {
        if (rp) {
          if (r_vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(rp + mb));
            v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += __ldg(rp + mb + j);
          }
        }
        if (y_vec) {
          *reinterpret_cast<float4*>(yp + mb) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) yp[mb + j] = v[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) st.add(v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (mb + j < d.M) {
            const float o = v[j] + (rp ? __ldg(rp + mb + j) : 0.f);
            yp[mb + j] = o;
            st.add(o);
          }
      }
    }
  }
  if (d.stats_partials) {
    Wf tot = wf_block_reduce(st.finish(), red);
    if (tid == 0) {
      const int64_t slot = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
      const int64_t slots = (int64_t)gridDim.x * gridDim.y;
      float* o = d.stats_partials + (b * slots + slot) * 3;
      o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// Short-K framed filterbank:  Y[b,r,m] = act( sum_{k<K} x[b, r*stride + k] * W[m,k] ),  K <= 64.
// The learned encoder of FreeEncDec (lobe/encoder.py:50-56,71-83) at win = 32, hop = 16 is 0.06 FLOP per output byte:
// it is a pure write stream (cfg2: 524 MB), which the 128 x 128 x 16 tile kernel above ran at 15 % of HBM (two k-slabs
// per tile, then a long epilogue).  Here a lane owns two output channels with their filter taps in REGISTERS, the CTA
// stages the contiguous piece of waveform its FR frames cover in shared memory once, and every frame is K broadcast
// loads + 2K FMAs + two coalesced 128-byte stores per warp.  Exact fp32, same summation order over k as the tile kernel.
// ---------------------------------------------------------------------------------------------------
constexpr int FB_FR = 256;       // frames per CTA (the per-CTA cost of fetching 2 x K filter taps per lane is amortised over them)
constexpr int FB_MAXSEG = 12288; // floats of waveform a CTA may stage (48 KB)

template <int KT>
__global__ void __launch_bounds__(256, KT == 32 ? 2 : 1) filterbank_kernel(const ps_gemm_t d, const int seg_len) {
  extern __shared__ __align__(16) float seg[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.x * FB_FR;
  const int nfr = (int)((d.rows - r0) < FB_FR ? (d.rows - r0) : FB_FR);
  const int64_t m0 = (int64_t)blockIdx.y * 512 + warp * 64 + lane;  // this lane's channels: m0 and m0 + 32
  const int K = (int)d.K, stride = (int)d.x_row_stride;
  const float* xb = d.X + b * d.x_batch_stride + r0 * stride;
  const int need = (nfr - 1) * stride + K;
  for (int i = tid; i < seg_len; i += blockDim.x) seg[i] = i < need ? __ldg(xb + i) : 0.f;
  float w0[KT], w1[KT];
  const bool wvec = ((d.w_row_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.W) & 15) == 0);
#pragma unroll
  for (int k = 0; k < KT; k += 4) {
    // 16-byte loads: a lane walks its own 128-byte filter row, so every fetched line is used completely
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
    if (wvec && k + 3 < K) {
      if (m0 < d.M) t0 = __ldg(reinterpret_cast<const float4*>(d.W + m0 * d.w_row_stride + k));
      if (m0 + 32 < d.M) t1 = __ldg(reinterpret_cast<const float4*>(d.W + (m0 + 32) * d.w_row_stride + k));
    } else {
      float e0[4] = {0.f, 0.f, 0.f, 0.f}, e1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (k + i < K && m0 < d.M) e0[i] = __ldg(d.W + m0 * d.w_row_stride + k + i);
        if (k + i < K && m0 + 32 < d.M) e1[i] = __ldg(d.W + (m0 + 32) * d.w_row_stride + k + i);
      }
      t0 = make_float4(e0[0], e0[1], e0[2], e0[3]);
      t1 = make_float4(e1[0], e1[1], e1[2], e1[3]);
    }
    w0[k] = t0.x; w0[k + 1] = t0.y; w0[k + 2] = t0.z; w0[k + 3] = t0.w;
    w1[k] = t1.x; w1[k + 1] = t1.y; w1[k + 2] = t1.z; w1[k + 3] = t1.w;
  }
  float bias0 = 0.f, bias1 = 0.f;
  if (d.bias) {
    if (m0 < d.M) bias0 = __ldg(d.bias + m0);
    if (m0 + 32 < d.M) bias1 = __ldg(d.bias + m0 + 32);
  }
  const float eslope = d.epi_slope ? __ldg(d.epi_slope) : 0.f;
  const int act = d.epi_act;
  __syncthreads();
  if (warp * 64 + blockIdx.y * 512 >= d.M) return;
  float* yp = d.Y + b * d.y_batch_stride + r0 * d.y_row_stride + m0;
  const bool vec = (stride & 3) == 0;
  // four frames at a time: eight independent accumulator chains per lane hide the FMA latency (one frame at a time
  // measured 0.51 ms at cfg2, latency-bound with 8 warps per SM)
  for (int f = 0; f < nfr; f += 4) {
    float acc[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; }
    const float* xs = seg + f * stride;  // frames f..f+3 lie inside the staged piece (FB_FR is a multiple of 4)
    if (vec) {
#pragma unroll
      for (int k = 0; k < KT; k += 4) {
        float4 x4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x4[j] = *reinterpret_cast<const float4*>(xs + j * stride + k);  // broadcast loads
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[j][0] = fmaf(x4[j].x, w0[k], acc[j][0]); acc[j][1] = fmaf(x4[j].x, w1[k], acc[j][1]);
          acc[j][0] = fmaf(x4[j].y, w0[k + 1], acc[j][0]); acc[j][1] = fmaf(x4[j].y, w1[k + 1], acc[j][1]);
          acc[j][0] = fmaf(x4[j].z, w0[k + 2], acc[j][0]); acc[j][1] = fmaf(x4[j].z, w1[k + 2], acc[j][1]);
          acc[j][0] = fmaf(x4[j].w, w0[k + 3], acc[j][0]); acc[j][1] = fmaf(x4[j].w, w1[k + 3], acc[j][1]);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < KT; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xv = xs[j * stride + k];
          acc[j][0] = fmaf(xv, w0[k], acc[j][0]);
          acc[j][1] = fmaf(xv, w1[k], acc[j][1]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (f + j < nfr) {
        if (m0 < d.M) yp[0] = apply_act(acc[j][0] + bias0, act, eslope);
        if (m0 + 32 < d.M) yp[32] = apply_act(acc[j][1] + bias1, act, eslope);
      }
      yp += d.y_row_stride;
    }
  }
}

static bool filterbank_eligible(const ps_gemm_t& d) {
  if (d.K > 64 || d.K % 4 != 0 || d.pro_mode != PS_PRO_NONE || d.residual || d.stats_partials || d.bias_batch) return false;
  if (d.rows < 256 || d.batch > 65535) return false;                          // skinny problems keep the latency tile
  if ((FB_FR + 2) * d.x_row_stride + 64 > FB_MAXSEG) return false;             // the staged waveform piece must fit
  return true;
}

static int filterbank_launch(const ps_gemm_t& d, cudaStream_t s) {
  const int KT = d.K <= 32 ? 32 : 64;
  int seg_len = (int)((FB_FR + 2) * d.x_row_stride + KT);  // frames are processed four at a time: a little slack past the last one
  seg_len = (seg_len + 3) & ~3;
  dim3 grid((unsigned)cdiv(d.rows, FB_FR), (unsigned)cdiv(d.M, 512), (unsigned)d.batch);
  const int threads = (int)(d.M >= 512 ? 256 : ((cdiv(d.M, 64) * 32 + 31) / 32 * 32));
  if (KT == 32) filterbank_kernel<32><<<grid, threads, seg_len * sizeof(float), s>>>(d, seg_len);
  else filterbank_kernel<64><<<grid, threads, seg_len * sizeof(float), s>>>(d, seg_len);
  PS_CHECK_LAUNCH("filterbank_kernel");
  return PS_OK;
}

// ---------------------------------------------------------------------------------------------------
// Thin outputs (M <= 8): Y[b,r,m] = act(bias[m] + sum_k x[b, r*stride + k] * W[m,k]).  The output layer of the U-Net shell
// (ConvTranspose2d to 2 channels, unet.py:154-170: M = 2, K = 256 / 384) and similar projections: a 128 x 128 tile would
// spend 98 % of its FMAs on padding.  One warp per row: lanes stride K with 16-byte loads (rows overlap, so consecutive
// rows of a CTA hit L1), the filter taps sit in shared memory, M butterfly reductions, lane m stores output m.
// ---------------------------------------------------------------------------------------------------
constexpr int THIN_MAXM = 8;
constexpr int THIN_MAXWK = 8192;  // floats of weights in shared memory (32 KB)

template <int TM>
__global__ void __launch_bounds__(256) thin_gemm_kernel(const ps_gemm_t d) {
  extern __shared__ __align__(16) float wsm[];  // [TM][K]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = (int)d.K, M = (int)d.M;
  for (int i = tid; i < TM * K; i += blockDim.x) wsm[i] = (i / K < M) ? __ldg(d.W + (int64_t)(i / K) * d.w_row_stride + i % K) : 0.f;
  __syncthreads();
  const float eslope = d.epi_slope ? __ldg(d.epi_slope) : 0.f;
  const float bias = (d.bias && lane < M) ? __ldg(d.bias + lane) : 0.f;
  const int64_t total = d.batch * d.rows;
  const int64_t per = (total + gridDim.x - 1) / gridDim.x;  // a CTA walks a contiguous run of rows (L1 reuse of the overlap)
  const int64_t g0 = (int64_t)blockIdx.x * per, g1 = (g0 + per < total) ? g0 + per : total;
  for (int64_t g = g0 + warp; g < g1; g += 8) {
    const int64_t b = g / d.rows, r = g % d.rows;
    const float* xr = d.X + b * d.x_batch_stride + r * d.x_row_stride;
    float acc[TM];
#pragma unroll
    for (int m = 0; m < TM; ++m) acc[m] = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 x4 = __ldg(reinterpret_cast<const float4*>(xr + k));
#pragma unroll
      for (int m = 0; m < TM; ++m) {
        const float4 w4 = *reinterpret_cast<const float4*>(wsm + m * K + k);
        acc[m] = fmaf(x4.x, w4.x, acc[m]); acc[m] = fmaf(x4.y, w4.y, acc[m]);
        acc[m] = fmaf(x4.z, w4.z, acc[m]); acc[m] = fmaf(x4.w, w4.w, acc[m]);
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int m = 0; m < TM; ++m) {
      const float v = warp_sum(acc[m]);
      if (lane == m) mine = v;
    }
    if (lane < M) d.Y[b * d.y_batch_stride + r * d.y_row_stride + lane] = apply_act(mine + bias, d.epi_act, eslope);
  }
}

static bool thin_eligible(const ps_gemm_t& d, int x_vec) {
  if (d.M > THIN_MAXM || !x_vec || d.K % 4 != 0 || d.K * THIN_MAXM > THIN_MAXWK * 1) return false;
  if (d.pro_mode != PS_PRO_NONE || d.residual || d.stats_partials || d.bias_batch || d.ln_eps > 0.f) return false;
  return d.batch * d.rows >= 4096;  // small problems keep the latency tile
}

static int thin_launch(const ps_gemm_t& d, cudaStream_t s) {
  const int64_t total = d.batch * d.rows;
  int64_t blocks = cdiv(total, 64);
  if (blocks > 148 * 8) blocks = 148 * 8;
  const int TM = d.M <= 2 ? 2 : (d.M <= 4 ? 4 : 8);
  const size_t smem = (size_t)TM * d.K * sizeof(float);
  if (TM == 2) thin_gemm_kernel<2><<<(unsigned)blocks, 256, smem, s>>>(d);
  else if (TM == 4) thin_gemm_kernel<4><<<(unsigned)blocks, 256, smem, s>>>(d);
  else thin_gemm_kernel<8><<<(unsigned)blocks, 256, smem, s>>>(d);
  PS_CHECK_LAUNCH("thin_gemm_kernel");
  return PS_OK;
}

// ---------------------------------------------------------------------------------------------------
// A handful of rows (batch * rows <= 32, eight at a time) against many output channels: the per-hop 1x1 convs of a single stream (cfg5 at
// S = 1: 72 GEMVs of 512 x 512 per 10 ms hop) and the heads applied to pooled embeddings.  The 32 x 64 latency tile ran such
// a product on 8 CTAs with 32 synchronised k-slabs (~13 us); here the transformed operand rows sit in shared memory, a WARP
// owns one output channel (its weight row is one coalesced 16-byte-per-lane stream), and 8 channels share a CTA: 64 CTAs,
// four load rounds per warp at K = 512.
// ---------------------------------------------------------------------------------------------------
constexpr int FEW_MAXR = 8;             // rows per pass over a weight row (accumulators per lane)
constexpr int FEW_MAXROWS = 32;         // rows per launch: up to four passes, the weight row re-read from L1
constexpr int FEW_MAXSMEM = 128 * 1024;

template <int PRO>
__global__ void __launch_bounds__(256) few_rows_kernel(const ps_gemm_t d, const int R) {
  extern __shared__ __align__(16) float xs[];  // [R][K], prologue applied
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = (int)d.K;
  XLoadCtx ctx;
  ctx.mean = 0.f; ctx.rstd = 1.f;
  ctx.slope = (d.pro_slope != nullptr) ? __ldg(d.pro_slope) : 0.f;
  for (int i = tid; i < R * K; i += 256) {
    const int g = i / K, k = i - g * K;
    const int64_t b = g / d.rows, r = g % d.rows;
    const int64_t off = b * d.x_batch_stride + r * d.x_row_stride + k;
    const float x2 = (PRO == PS_PRO_MASK) ? __ldg(d.X2 + off) : 0.f;
    xs[i] = pro_apply<PRO>(d, __ldg(d.X + off), x2, b, k, ctx);
  }
  __syncthreads();
  const int64_t m = (int64_t)blockIdx.x * 8 + warp;
  if (m >= d.M) return;
  const float* wr = d.W + m * d.w_row_stride;
  const float eslope = d.epi_slope ? __ldg(d.epi_slope) : 0.f;
  float bsum = d.bias ? __ldg(d.bias + m) : 0.f;
  for (int g0 = 0; g0 < R; g0 += FEW_MAXR) {
    const int Rg = (R - g0) < FEW_MAXR ? (R - g0) : FEW_MAXR;
    const float* xg = xs + g0 * K;
    float acc[FEW_MAXR];
#pragma unroll
    for (int g = 0; g < FEW_MAXR; ++g) acc[g] = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
      for (int g = 0; g < FEW_MAXR; ++g) {
        if (g < Rg) {
          const float4 x4 = *reinterpret_cast<const float4*>(xg + g * K + k);
          acc[g] = fmaf(x4.x, w4.x, acc[g]); acc[g] = fmaf(x4.y, w4.y, acc[g]);
          acc[g] = fmaf(x4.z, w4.z, acc[g]); acc[g] = fmaf(x4.w, w4.w, acc[g]);
        }
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int g = 0; g < FEW_MAXR; ++g) {
      if (g < Rg) {
        const float v = warp_sum(acc[g]);
        if (lane == g) mine = v;
      }
    }
    if (lane < Rg) {
      const int64_t row = g0 + lane, b = row / d.rows, r = row % d.rows;
      float v = mine + bsum;
      if (d.bias_batch) v += __ldg(d.bias_batch + b * d.M + m);
      v = apply_act(v, d.epi_act, eslope);
      if (d.residual) v += __ldg(d.residual + b * d.res_batch_stride + r * d.res_row_stride + m);
      d.Y[b * d.y_batch_stride + r * d.y_row_stride + m] = v;
    }
  }
}

static bool few_rows_eligible(const ps_gemm_t& d, int x_vec, int w_vec) {
  const int64_t R = d.batch * d.rows;
  if (R > FEW_MAXROWS || !x_vec || !w_vec || d.K % 4 != 0 || R * d.K * (int64_t)sizeof(float) > FEW_MAXSMEM) return false;
  if (!(d.pro_mode == PS_PRO_NONE || d.pro_mode == PS_PRO_AFFINE || d.pro_mode == PS_PRO_MASK)) return false;
  if (d.stats_partials || d.ln_eps > 0.f || d.M < 32) return false;
  return getenv("PS_GEMM_NO_FEW_ROWS") == nullptr;  // (A/B switch)
}

static int few_rows_launch(const ps_gemm_t& d, cudaStream_t s) {
  const int R = (int)(d.batch * d.rows);
  const size_t smem = (size_t)R * d.K * sizeof(float);
  const unsigned blocks = (unsigned)cdiv(d.M, 8);
  if (smem > 48 * 1024) {
    static SmemOnce<3> once;
    int dev = 0;
    if (int rc = current_device(&dev)) return rc;
    if (int rc = once.ensure(dev, 0, few_rows_kernel<PS_PRO_NONE>, FEW_MAXSMEM, "cudaFuncSetAttribute(few_rows_kernel)")) return rc;
    if (int rc = once.ensure(dev, 1, few_rows_kernel<PS_PRO_AFFINE>, FEW_MAXSMEM, "cudaFuncSetAttribute(few_rows_kernel)")) return rc;
    if (int rc = once.ensure(dev, 2, few_rows_kernel<PS_PRO_MASK>, FEW_MAXSMEM, "cudaFuncSetAttribute(few_rows_kernel)")) return rc;
  }
  switch (d.pro_mode) {
    case PS_PRO_AFFINE: few_rows_kernel<PS_PRO_AFFINE><<<blocks, 256, smem, s>>>(d, R); break;
    case PS_PRO_MASK: few_rows_kernel<PS_PRO_MASK><<<blocks, 256, smem, s>>>(d, R); break;
    default: few_rows_kernel<PS_PRO_NONE><<<blocks, 256, smem, s>>>(d, R); break;
  }
  PS_CHECK_LAUNCH("few_rows_kernel");
  return PS_OK;
}

template <int BR, int BC>
static int launch_shape(const ps_gemm_t& d, cudaStream_t s, int x_vec, int w_vec) {
  const int64_t nrt = cdiv(d.rows, BR), nmt = cdiv(d.M, BC);
  if (nrt > 65535 || d.batch > 65535) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)nmt, (unsigned)nrt, (unsigned)d.batch);
  switch (d.pro_mode) {
    case PS_PRO_AFFINE: gemm_simt_kernel<PS_PRO_AFFINE, BR, BC><<<grid, NT, 0, s>>>(d, x_vec, w_vec); break;
    case PS_PRO_ROWNORM: gemm_simt_kernel<PS_PRO_ROWNORM, BR, BC><<<grid, NT, 0, s>>>(d, x_vec, w_vec); break;
    case PS_PRO_MASK: gemm_simt_kernel<PS_PRO_MASK, BR, BC><<<grid, NT, 0, s>>>(d, x_vec, w_vec); break;
    default: gemm_simt_kernel<PS_PRO_NONE, BR, BC><<<grid, NT, 0, s>>>(d, x_vec, w_vec); break;
  }
  PS_CHECK_LAUNCH("gemm_simt_kernel");
  return PS_OK;
}

int gemm_simt_launch(const ps_gemm_t& d, cudaStream_t s) {
  const int x_vec = ((d.x_row_stride & 3) == 0) && ((d.x_batch_stride & 3) == 0) &&
                    ((reinterpret_cast<uintptr_t>(d.X) & 15) == 0) &&
                    (!d.X2 || (reinterpret_cast<uintptr_t>(d.X2) & 15) == 0);
  const int w_vec = ((d.w_row_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.W) & 15) == 0);
  if (filterbank_eligible(d)) return filterbank_launch(d, s);
  if (few_rows_eligible(d, x_vec, w_vec)) return few_rows_launch(d, s);
  if (thin_eligible(d, x_vec)) return thin_launch(d, s);
  // latency shape when the throughput shape would leave most SMs idle (and no statistics are requested: the partial
  // slot layout is defined on 128 x 128 tiles)
  const int64_t big_ctas = d.batch * cdiv(d.rows, 128) * cdiv(d.M, 128);
  if (!d.stats_partials && big_ctas < 96) return launch_shape<32, 64>(d, s, x_vec, w_vec);
  return launch_shape<128, 128>(d, s, x_vec, w_vec);
}

}  // namespace ps
