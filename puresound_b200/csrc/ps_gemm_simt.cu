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
#include "ps_common.cuh"

namespace ps {

constexpr int BK = 16, NT = 256;

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
  float acc[TR][TCc];
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < TCc; ++j) acc[i][j] = 0.f;

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
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TCc; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
    if (kt + 1 < nkt) {
      store_s();
      __syncthreads();
    }
  }

  // ---- epilogue ----
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
  // latency shape when the throughput shape would leave most SMs idle (and no statistics are requested: the partial
  // slot layout is defined on 128 x 128 tiles)
  const int64_t big_ctas = d.batch * cdiv(d.rows, 128) * cdiv(d.M, 128);
  if (!d.stats_partials && big_ctas < 96) return launch_shape<32, 64>(d, s, x_vec, w_vec);
  return launch_shape<128, 128>(d, s, x_vec, w_vec);
}

}  // namespace ps
