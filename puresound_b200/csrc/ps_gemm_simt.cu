// Exact-fp32 CUDA-core GEMM with fused prologue/epilogue.
//
//   Y[b,r,m] = epi( sum_k pro(X[b,r,k]) * W[m,k] )
//
// This is the reference-accuracy back end (true fp32 FMA accumulation): it serves
// every shape the tcgen05 back end does not take (ragged K/M, framed encoder views
// with row stride < K, the iSTFT synthesis GEMM that must stay fp32 because of the
// 2.6e4x edge amplification of the window-sumsquare division, tiny streaming steps)
// and is the on-device cross-check for the tensor-core kernel.
//
// Tile 128 rows x 128 out-channels x 16 k, 256 threads, 8x8 outputs per thread,
// register prefetch of the next k-slab while the current one is consumed from
// shared memory.  The prologue (norm affine + activation, cLN row-norm, or mask
// product) is applied on the global->register leg so shared memory already holds
// the transformed operand; the epilogue adds bias / per-item bias / activation /
// residual and emits one Welford partial (count, mean, M2) per CTA.
#include "ps_common.cuh"

namespace ps {

constexpr int BR = 128, BC = 128, BK = 16, NT = 256, LDS = 132;

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

template <int PRO>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const ps_gemm_t d, const int x_vec, const int w_vec) {
  __shared__ __align__(16) float Xs[BK][LDS];
  __shared__ __align__(16) float Ws[BK][LDS];
  __shared__ Wf red[NT / 32];

  const int tid = threadIdx.x;
  const int64_t b = blockIdx.z;
  const int64_t row0 = (int64_t)blockIdx.y * BR;
  const int64_t m0 = (int64_t)blockIdx.x * BC;

  // loader mapping: 2 threads per tile row, 8 consecutive k each
  const int lrow = tid >> 1;
  const int lk0 = (tid & 1) * 8;
  const int64_t xr = row0 + lrow;
  const int64_t wm = m0 + lrow;
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

  float xreg[8], wreg[8];
  auto load_g = [&](int64_t k0) {
    const int64_t kb = k0 + lk0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t k = kb + h * 4;
      float xv[4] = {0.f, 0.f, 0.f, 0.f}, x2v[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
      if (xr_ok) {
        if (x_vec && k + 3 < d.K) {
          float4 t = __ldg(reinterpret_cast<const float4*>(xp + k));
          xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
          if (x2p) {
            float4 u = __ldg(reinterpret_cast<const float4*>(x2p + k));
            x2v[0] = u.x; x2v[1] = u.y; x2v[2] = u.z; x2v[3] = u.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (k + i < d.K) {
              xv[i] = __ldg(xp + k + i);
              if (x2p) x2v[i] = __ldg(x2p + k + i);
            }
        }
      }
      if (wm_ok) {
        if (w_vec && k + 3 < d.K) {
          float4 t = __ldg(reinterpret_cast<const float4*>(wp + k));
          wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (k + i < d.K) wv[i] = __ldg(wp + k + i);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = xr_ok && (k + i < d.K);
        xreg[h * 4 + i] = ok ? pro_apply<PRO>(d, xv[i], x2v[i], b, k + i, ctx) : 0.f;
        wreg[h * 4 + i] = wv[i];
      }
    }
  };
  auto store_s = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      Xs[lk0 + i][lrow] = xreg[i];
      Ws[lk0 + i][lrow] = wreg[i];
    }
  };

  // compute mapping: 16x16 threads; rows ty*8..+7; cols tx*4..+3 and 64+tx*4..+3
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int64_t nkt = (d.K + BK - 1) / BK;
  load_g(0);
  store_s();
  __syncthreads();
  for (int64_t kt = 0; kt < nkt; ++kt) {
    if (kt + 1 < nkt) load_g((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], w[8];
      float4 a0 = *reinterpret_cast<const float4*>(&Xs[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&Xs[k][ty * 8 + 4]);
      float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
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
  float bj[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t m = m0 + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4)));
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
  for (int i = 0; i < 8; ++i) {
    const int64_t r = row0 + ty * 8 + i;
    if (r >= d.rows) continue;
    float* yp = d.Y + b * d.y_batch_stride + r * d.y_row_stride;
    const float* rp = d.residual ? d.residual + b * d.res_batch_stride + r * d.res_row_stride : nullptr;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t mb = m0 + (h ? 64 + tx * 4 : tx * 4);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = apply_act(acc[i][h * 4 + j] + bj[h * 4 + j], d.epi_act, eslope);
      if (mb + 3 < d.M) {
        if (rp) {
          if (r_vec) {
            float4 t = __ldg(reinterpret_cast<const float4*>(rp + mb));
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
            float o = v[j] + (rp ? __ldg(rp + mb + j) : 0.f);
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

int gemm_simt_launch(const ps_gemm_t& d, cudaStream_t s) {
  const int64_t nrt = cdiv(d.rows, BR), nmt = cdiv(d.M, BC);
  if (nrt > 65535 || d.batch > 65535) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)nmt, (unsigned)nrt, (unsigned)d.batch);
  const int x_vec = ((d.x_row_stride & 3) == 0) && ((d.x_batch_stride & 3) == 0) &&
                    ((reinterpret_cast<uintptr_t>(d.X) & 15) == 0) &&
                    (!d.X2 || (reinterpret_cast<uintptr_t>(d.X2) & 15) == 0);
  const int w_vec = ((d.w_row_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.W) & 15) == 0);
  switch (d.pro_mode) {
    case PS_PRO_AFFINE: gemm_simt_kernel<PS_PRO_AFFINE><<<grid, NT, 0, s>>>(d, x_vec, w_vec); break;
    case PS_PRO_ROWNORM: gemm_simt_kernel<PS_PRO_ROWNORM><<<grid, NT, 0, s>>>(d, x_vec, w_vec); break;
    case PS_PRO_MASK: gemm_simt_kernel<PS_PRO_MASK><<<grid, NT, 0, s>>>(d, x_vec, w_vec); break;
    default: gemm_simt_kernel<PS_PRO_NONE><<<grid, NT, 0, s>>>(d, x_vec, w_vec); break;
  }
  PS_CHECK_LAUNCH("gemm_simt_kernel");
  return PS_OK;
}

}  // namespace ps
