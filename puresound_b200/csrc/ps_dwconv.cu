// Dilated depthwise Conv1d over frames-major activations with the preceding
// norm + PReLU applied on load and Welford partials of the output on store.
//
// HBM-bound: each CTA streams a tile of TT consecutive frames for a group of
// channels; the +-d dilation taps of neighbouring tiles re-hit L2 (tile rows are
// whole 2 KB frame rows at C=512), so compulsory traffic is one read and one
// write of the tensor.  Threads run along channels (float4 per thread), so every
// global access is a fully coalesced 512 B warp request.
#include "ps_common.cuh"

namespace ps {

constexpr int DW_TT = 16;       // frames per CTA
constexpr int DW_THREADS = 128;

struct DwGeom {
  int vec, threads_c, rows_par, chan_per_block;
};

static inline DwGeom dw_geom(int64_t C) {
  DwGeom g;
  g.vec = (C % 4 == 0) ? 4 : 1;
  int64_t groups = cdiv(C, g.vec);
  int tc = groups <= 32 ? 32 : (groups <= 64 ? 64 : DW_THREADS);  // power of two so it divides the block
  g.threads_c = tc;
  g.rows_par = DW_THREADS / tc;
  g.chan_per_block = tc * g.vec;
  return g;
}

template <int VEC>
struct VecT;
template <>
struct VecT<4> {
  using T = float4;
};
template <>
struct VecT<1> {
  using T = float;
};

template <int VEC>
__device__ __forceinline__ void ldv(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = __ldg(p);
  }
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    p[0] = v[0];
  }
}

template <int VEC, int PT>  // PT = compile-time taps (3) or 0 for a runtime count <= 8
__global__ void __launch_bounds__(DW_THREADS) dwconv_kernel(const ps_dwconv_t d, const int threads_c, const int rows_par) {
  __shared__ Wf red[DW_THREADS / 32];
  const int tc = threadIdx.x % threads_c;
  const int tr = threadIdx.x / threads_c;
  const int64_t b = blockIdx.z;
  const int64_t c0 = ((int64_t)blockIdx.y * threads_c + tc) * VEC;
  const int64_t t0 = (int64_t)blockIdx.x * DW_TT;
  const bool c_ok = c0 < d.C;
  const int P = PT ? PT : d.P;
  const float slope = d.pro_slope ? __ldg(d.pro_slope) : 0.f;

  float w[PT ? PT : 1][VEC], bias[VEC], pa[VEC], pb[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { bias[i] = 0.f; pa[i] = 1.f; pb[i] = 0.f; }
  if (c_ok) {
    if constexpr (PT > 0) {
#pragma unroll
      for (int p = 0; p < PT; ++p)
#pragma unroll
        for (int i = 0; i < VEC; ++i) w[p][i] = __ldg(d.w + (c0 + i) * PT + p);
    }
    if (d.bias) ldv<VEC>(d.bias + c0, bias);
    if (d.pro_mode == PS_PRO_AFFINE) {
      ldv<VEC>(d.pro_a + b * d.pro_batch_stride + c0, pa);
      ldv<VEC>(d.pro_b + b * d.pro_batch_stride + c0, pb);
    } else if (d.pro_mode == PS_PRO_ROWNORM) {
      ldv<VEC>(d.pro_a + c0, pa);
      ldv<VEC>(d.pro_b + c0, pb);
    }
  }

  WfAcc st;
  st.init();
  const float* xb = d.x + b * d.T * d.C;
  float* yb = d.y + b * d.T * d.C;
  if (c_ok) {
    for (int64_t t = t0 + tr; t < t0 + DW_TT && t < d.T; t += rows_par) {
      float acc[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = bias[i];
#pragma unroll
      for (int p = 0; p < (PT ? PT : 8); ++p) {
        if (p >= P) break;
        const int64_t tt = d.causal ? t - (int64_t)(P - 1 - p) * d.dilation : t + (int64_t)(p - (P - 1) / 2) * d.dilation;
        if (tt < 0 || tt >= d.T) continue;
        float v[VEC];
        ldv<VEC>(xb + tt * d.C + c0, v);
        float mean = 0.f, rstd = 1.f;
        if (d.pro_mode == PS_PRO_ROWNORM) {
          const float* rs = d.pro_rowstats + (b * d.T + tt) * 2;
          mean = __ldg(rs);
          rstd = __ldg(rs + 1);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float u = v[i];
          if (d.pro_mode == PS_PRO_AFFINE) u = apply_act(fmaf(u, pa[i], pb[i]), d.pro_act, slope);
          else if (d.pro_mode == PS_PRO_ROWNORM) u = apply_act(fmaf((u - mean) * rstd, pa[i], pb[i]), d.pro_act, slope);
          float wv;
          if constexpr (PT > 0) wv = w[p][i];
          else wv = __ldg(d.w + (c0 + i) * P + p);
          acc[i] = fmaf(wv, u, acc[i]);
        }
      }
      stv<VEC>(yb + t * d.C + c0, acc);
#pragma unroll
      for (int i = 0; i < VEC; ++i) st.add(acc[i]);
    }
  }
  if (d.stats_partials) {
    Wf tot = wf_block_reduce(st.finish(), red);
    if (threadIdx.x == 0) {
      const int64_t slot = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
      const int64_t slots = (int64_t)gridDim.x * gridDim.y;
      float* o = d.stats_partials + (b * slots + slot) * 3;
      o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
    }
  }
}

}  // namespace ps

extern "C" int64_t ps_dwconv_stats_slots(int64_t T, int64_t C) {
  if (T <= 0 || C <= 0) return 0;
  ps::DwGeom g = ps::dw_geom(C);
  return ps::cdiv(T, ps::DW_TT) * ps::cdiv(C, g.chan_per_block);
}

extern "C" int ps_dwconv(const ps_dwconv_t* dp, void* stream) {
  PS_REQUIRE(dp != nullptr);
  const ps_dwconv_t& d = *dp;
  PS_REQUIRE(d.x && d.y && d.w && d.batch > 0 && d.T > 0 && d.C > 0 && d.P >= 1 && d.P <= 8 && d.dilation >= 1);
  PS_REQUIRE(d.causal || (d.P % 2 == 1));
  PS_REQUIRE(d.pro_mode == PS_PRO_NONE || d.pro_mode == PS_PRO_AFFINE || d.pro_mode == PS_PRO_ROWNORM);
  if (d.pro_mode != PS_PRO_NONE) PS_REQUIRE(d.pro_a && d.pro_b);
  if (d.pro_mode == PS_PRO_ROWNORM) PS_REQUIRE(d.pro_rowstats);
  if (d.pro_act == PS_ACT_PRELU) PS_REQUIRE(d.pro_slope);
  if (d.batch > 65535) return PS_ERR_UNSUPPORTED;
  ps::DwGeom g = ps::dw_geom(d.C);
  if (g.vec == 4) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    PS_REQUIRE(al(d.x) && al(d.y) && (!d.bias || al(d.bias)) && (d.pro_mode == PS_PRO_NONE || (al(d.pro_a) && al(d.pro_b))));
    if (d.pro_mode == PS_PRO_AFFINE) PS_REQUIRE(d.pro_batch_stride % 4 == 0);
  }
  dim3 grid((unsigned)ps::cdiv(d.T, ps::DW_TT), (unsigned)ps::cdiv(d.C, g.chan_per_block), (unsigned)d.batch);
  cudaStream_t s = (cudaStream_t)stream;
  if (g.vec == 4) {
    if (d.P == 3) ps::dwconv_kernel<4, 3><<<grid, ps::DW_THREADS, 0, s>>>(d, g.threads_c, g.rows_par);
    else ps::dwconv_kernel<4, 0><<<grid, ps::DW_THREADS, 0, s>>>(d, g.threads_c, g.rows_par);
  } else {
    if (d.P == 3) ps::dwconv_kernel<1, 3><<<grid, ps::DW_THREADS, 0, s>>>(d, g.threads_c, g.rows_par);
    else ps::dwconv_kernel<1, 0><<<grid, ps::DW_THREADS, 0, s>>>(d, g.threads_c, g.rows_par);
  }
  PS_CHECK_LAUNCH("dwconv_kernel");
  return PS_OK;
}
