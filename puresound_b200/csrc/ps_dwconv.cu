// Dilated depthwise Conv1d over frames-major activations with the preceding
// norm + PReLU applied on load and Welford partials of the output on store.
//
// HBM-bound: each CTA streams a tile of TT consecutive frames for a group of
// channels; the +-d dilation taps of neighbouring tiles re-hit L2 (tile rows are
// whole 2 KB frame rows at C=512), so compulsory traffic is one read and one
// write of the tensor.  Threads run along channels (float4 per thread), so every
// global access is a fully coalesced 512 B warp request.
#include <stdlib.h>

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
      float* o = d.stats_partials + (b * d.stats_slots + slot) * 3;
      o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// Tiled variant (the one used for the TCN shapes).  The streaming kernel above reads every input element
// P times; with dilations up to 128 those re-reads are L2 hits but still cost L2->SM bandwidth (measured:
// 44 % of HBM peak, L2-bound).  Here a CTA owns TC frames x 32 channels of one item, stages the tile plus
// its (P-1)*d halo in shared memory ONCE — already transformed by the norm+PReLU prologue, so the
// transform is also done once instead of P times — and computes all taps from shared memory.
// Read amplification drops from P to (TC + (P-1)d)/TC (1.25x averaged over d = 1..128 at TC = 256).
// ---------------------------------------------------------------------------------------------------
constexpr int DT_THREADS = 256;
constexpr int DT_CG = 32;  // channels per CTA: one 128-byte row segment

static inline int dt_chunk(int P, int dilation) {
  static EnvInt env;  // PS_DW_TC: frames per CTA for A/B runs (the statistics slot count assumes >= 256)
  const int forced = env.get("PS_DW_TC", 0);
  if (forced >= 256) return forced;
  return ((P - 1) * dilation <= 128) ? 256 : 512;
}

// PT: compile-time taps (3) or 0 = runtime (<= 8).  PRO: 0 none, 1 folded affine + PReLU, 2 row-norm (cLN) + PReLU;
// "no activation" runs as PReLU with slope 1.  Everything per-element is compile-time selected: with runtime
// mode/activation dispatch inside the unrolled loops this kernel executed 65 instructions per element and was
// issue-bound at 27 % of HBM peak (ncu, round 1 run 5).
// MINB: CTAs per SM the register allocation must allow (4 caps the kernel at 64 registers: four 54 KB CTAs = 32 warps per SM
// instead of three at 80 registers)
template <int PT, int PRO, int MINB = 1>
__global__ void __launch_bounds__(DT_THREADS, MINB) dwconv_tile_kernel(const ps_dwconv_t d, const int TC) {
  extern __shared__ __align__(16) float tile[];  // [(TC + halo) rows][32 channels]
  __shared__ Wf red[DT_THREADS / 32];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int64_t b = blockIdx.z;
  // blockIdx.x = channel group (fastest): the 16 CTAs that share a time chunk run together, so every 2 KB frame row is
  // fetched from HBM in one go instead of as sixteen 128-byte pieces spread over the kernel's lifetime
  const int c0 = blockIdx.x * DT_CG + tx * 4;
  const int t0 = blockIdx.y * TC;
  const int T = (int)d.T, C = (int)d.C;
  const int P = PT ? PT : d.P;
  const int dil = d.dilation;
  const int halo = (P - 1) * dil;
  const int halo_l = d.causal ? halo : halo / 2;
  const int rows_total = TC + halo;
  const bool c_ok = c0 < C;  // C % 4 == 0, so the whole float4 is in range
  const float slope = (d.pro_act == PS_ACT_PRELU && d.pro_slope) ? __ldg(d.pro_slope) : 1.f;

  float pa[4] = {1.f, 1.f, 1.f, 1.f}, pb[4] = {0.f, 0.f, 0.f, 0.f}, bias[4] = {0.f, 0.f, 0.f, 0.f};
  float w[PT ? PT : 1][4];
  if (c_ok) {
    if constexpr (PRO == 1) {
      ldv<4>(d.pro_a + b * d.pro_batch_stride + c0, pa);
      ldv<4>(d.pro_b + b * d.pro_batch_stride + c0, pb);
    } else if constexpr (PRO == 2) {
      ldv<4>(d.pro_a + c0, pa);
      ldv<4>(d.pro_b + c0, pb);
    }
    if (d.bias) ldv<4>(d.bias + c0, bias);
    if constexpr (PT > 0) {
#pragma unroll
      for (int p = 0; p < PT; ++p)
#pragma unroll
        for (int i = 0; i < 4; ++i) w[p][i] = __ldg(d.w + (c0 + i) * PT + p);
    }
  }
  const float* xb = d.x + b * d.T * d.C + c0;

  // ---- stage: 8 independent 128-bit loads per thread in flight, transform once, store to shared memory ----
  for (int base = 0; base < rows_total; base += 32 * 8) {
    float4 v[8];
    float2 rs[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int rr = base + u * 32 + ty;
      const int t = t0 - halo_l + rr;
      const bool ok = c_ok && rr < rows_total && t >= 0 && t < T;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      rs[u] = make_float2(0.f, 1.f);
      if (ok) {
        v[u] = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)t * C));
        if constexpr (PRO == 2) rs[u] = __ldg(reinterpret_cast<const float2*>(d.pro_rowstats + (b * d.T + t) * 2));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int rr = base + u * 32 + ty;
      if (rr >= rows_total) continue;
      const int t = t0 - halo_l + rr;
      const bool ok = c_ok && t >= 0 && t < T;
      float o[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      if constexpr (PRO != 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float z = o[i];
          if constexpr (PRO == 2) z = (z - rs[u].x) * rs[u].y;
          z = fmaf(z, pa[i], pb[i]);
          z = z > 0.f ? z : z * slope;
          o[i] = ok ? z : 0.f;  // zero padding applies AFTER the prologue: the reference pads the activated tensor
        }
      }
      *reinterpret_cast<float4*>(tile + rr * DT_CG + tx * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();

  // ---- taps from shared memory; statistics as shifted sums about a per-thread pivot (count known analytically) ----
  float piv = 0.f, ssum = 0.f, ssq = 0.f;
  int nout = 0;
  float* yb = d.y + b * d.T * d.C + c0;
  if (c_ok) {
    const int jmax = (T - t0) < TC ? (T - t0) : TC;
    for (int j = ty; j < jmax; j += 32) {
      float acc[4] = {bias[0], bias[1], bias[2], bias[3]};
#pragma unroll
      for (int p = 0; p < (PT ? PT : 8); ++p) {
        if (p >= P) break;
        const float4 u4 = *reinterpret_cast<const float4*>(tile + (j + p * dil) * DT_CG + tx * 4);
        const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float wv;
          if constexpr (PT > 0) wv = w[p][i];
          else wv = __ldg(d.w + (c0 + i) * P + p);
          acc[i] = fmaf(wv, uu[i], acc[i]);
        }
      }
      *reinterpret_cast<float4*>(yb + (int64_t)(t0 + j) * C) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      if (nout == 0) piv = acc[0];
      ++nout;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dv = acc[i] - piv;
        ssum += dv;
        ssq = fmaf(dv, dv, ssq);
      }
    }
  }
  if (d.stats_partials) {
    Wf mine;
    mine.n = (float)(nout * 4);
    mine.mean = 0.f; mine.m2 = 0.f;
    if (nout > 0) {
      const float md = ssum / mine.n;
      mine.mean = piv + md;
      mine.m2 = fmaxf(ssq - ssum * md, 0.f);
    }
    Wf tot = wf_block_reduce(mine, red);
    __shared__ int fin_last;
    if (threadIdx.x == 0) {
      const int64_t slot = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
      float* o = d.stats_partials + (b * d.stats_slots + slot) * 3;
      o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
      if (d.fin_scale) {
        // fused gLN/gGN finalize: the CTA that writes an item's last partial merges them all (see ps_gemm_pair.cu)
        __threadfence();
        const unsigned int total = gridDim.x * gridDim.y;
        fin_last = (atomicAdd(d.fin_counter + b, 1u) + 1u == total) ? 1 : 0;
      }
    }
    if (d.fin_scale) {
      static_assert(DT_THREADS == 256, "the fused finalize runs on a 256-thread CTA");
      __shared__ double fin_s[3 * 256];
      __syncthreads();
      if (fin_last) {
        __threadfence();
        stats_finalize_item<0>(d.stats_partials + b * d.stats_slots * 3, d.stats_slots, d.fin_gamma, d.fin_beta, d.fin_eps, d.C,
                               d.fin_scale + b * d.C, d.fin_shift + b * d.C, nullptr, (int)threadIdx.x, fin_s, fin_s + 256, fin_s + 512);
        if (threadIdx.x == 0) d.fin_counter[b] = 0;
      }
    }
  }
}

template <int PT, int PRO, int MINB = 1>
static int launch_tile(const ps_dwconv_t& dd, int TC, size_t smem, dim3 grid, cudaStream_t s) {
  static SmemOnce<1> once;  // per instantiation and device
  int dev = 0;
  if (int rc = current_device(&dev)) return rc;
  if (int rc = once.ensure(dev, 0, dwconv_tile_kernel<PT, PRO, MINB>, 200 * 1024, "cudaFuncSetAttribute(dwconv_tile_kernel)")) return rc;
  dwconv_tile_kernel<PT, PRO, MINB><<<grid, DT_THREADS, smem, s>>>(dd, TC);
  PS_CHECK_LAUNCH("dwconv_tile_kernel");
  return PS_OK;
}

}  // namespace ps

namespace ps {
bool dwconv_tma_eligible(const ps_dwconv_t& d);
int dwconv_tma_launch(const ps_dwconv_t& d, cudaStream_t s);
int64_t dwconv_tma_slots(int64_t T, int64_t C);
}

extern "C" int64_t ps_dwconv_stats_slots(int64_t T, int64_t C) {
  if (T <= 0 || C <= 0) return 0;
  ps::DwGeom g = ps::dw_geom(C);
  const int64_t a = ps::cdiv(T, ps::DW_TT) * ps::cdiv(C, g.chan_per_block);  // streaming kernel
  int64_t t = ps::cdiv(T, 256) * ps::cdiv(C, ps::DT_CG);                     // tiled kernel (its largest grid)
  if (C % 32 == 0 && ps::dwconv_tma_slots(T, C) > t) t = ps::dwconv_tma_slots(T, C);  // TMA sliding-window kernel
  return a > t ? a : t;
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
  if (d.fin_scale) PS_REQUIRE(d.stats_partials && d.fin_shift && d.fin_counter);
  if (d.batch > 65535) return PS_ERR_UNSUPPORTED;
  ps::DwGeom g = ps::dw_geom(d.C);
  if (g.vec == 4) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    PS_REQUIRE(al(d.x) && al(d.y) && (!d.bias || al(d.bias)) && (d.pro_mode == PS_PRO_NONE || (al(d.pro_a) && al(d.pro_b))));
    if (d.pro_mode == PS_PRO_AFFINE) PS_REQUIRE(d.pro_batch_stride % 4 == 0);
  }
  cudaStream_t s = (cudaStream_t)stream;
  ps_dwconv_t dd = d;
  dd.stats_slots = ps_dwconv_stats_slots(d.T, d.C);
  if (ps::dwconv_tma_eligible(dd)) return ps::dwconv_tma_launch(dd, s);  // TMA-fed sliding window (ps_dwconv_tma.cu); zeroes its own unused slots
  if (d.stats_partials) {
    // both kernels fill a prefix of the slot array; the rest must read as empty (count 0) partials
    cudaError_t e = cudaMemsetAsync(d.stats_partials, 0, (size_t)d.batch * dd.stats_slots * 3 * sizeof(float), s);
    if (e != cudaSuccess) { ps::set_cuda_error(e, "cudaMemsetAsync(stats_partials)"); return PS_ERR_CUDA; }
  }
  const int halo = (d.P - 1) * d.dilation;
  const bool act_ok = d.pro_mode == PS_PRO_NONE || d.pro_act == PS_ACT_PRELU || d.pro_act == PS_ACT_NONE;
  if (g.vec == 4 && halo <= 1024 && act_ok && d.T < (1 << 30) && !getenv("PS_DWCONV_STREAMING")) {
    // the gLN-prologue variant runs its 64-register build (four CTAs = 32 warps per SM instead of three: cfg2 step 42.1 ->
    // 40.2 ms on the same box, run 98); PS_DW_LB4=0 restores the 80-register one for A/B runs
    static ps::EnvInt lb4_env;
    const int lb4 = lb4_env.get("PS_DW_LB4", 1);
    const int TC = ps::dt_chunk(d.P, d.dilation);
    const size_t smem = (size_t)(TC + halo) * ps::DT_CG * sizeof(float);
    int which = (d.P == 3 ? 3 : 0) + d.pro_mode;
    if (which == 4 && lb4) which = 6;
    dim3 tgrid((unsigned)ps::cdiv(d.C, ps::DT_CG), (unsigned)ps::cdiv(d.T, TC), (unsigned)d.batch);
    switch (which) {
      case 0: return ps::launch_tile<0, 0>(dd, TC, smem, tgrid, s);
      case 1: return ps::launch_tile<0, 1>(dd, TC, smem, tgrid, s);
      case 2: return ps::launch_tile<0, 2>(dd, TC, smem, tgrid, s);
      case 3: return ps::launch_tile<3, 0>(dd, TC, smem, tgrid, s);
      case 4: return ps::launch_tile<3, 1>(dd, TC, smem, tgrid, s);
      case 6: return ps::launch_tile<3, 1, 4>(dd, TC, smem, tgrid, s);
      default: return ps::launch_tile<3, 2>(dd, TC, smem, tgrid, s);
    }
  }
  dim3 grid((unsigned)ps::cdiv(d.T, ps::DW_TT), (unsigned)ps::cdiv(d.C, g.chan_per_block), (unsigned)d.batch);
  dd.fin_scale = nullptr;  // the streaming kernel does not fuse the finalize: a follow-up launch below does it
  if (g.vec == 4) {
    if (d.P == 3) ps::dwconv_kernel<4, 3><<<grid, ps::DW_THREADS, 0, s>>>(dd, g.threads_c, g.rows_par);
    else ps::dwconv_kernel<4, 0><<<grid, ps::DW_THREADS, 0, s>>>(dd, g.threads_c, g.rows_par);
  } else {
    if (d.P == 3) ps::dwconv_kernel<1, 3><<<grid, ps::DW_THREADS, 0, s>>>(dd, g.threads_c, g.rows_par);
    else ps::dwconv_kernel<1, 0><<<grid, ps::DW_THREADS, 0, s>>>(dd, g.threads_c, g.rows_par);
  }
  PS_CHECK_LAUNCH("dwconv_kernel");
  if (d.fin_scale)
    return ps_stats_finalize(d.stats_partials, d.batch, dd.stats_slots, d.fin_gamma, d.fin_beta, d.fin_eps, d.C, d.fin_scale,
                             d.fin_shift, nullptr, stream);
  return PS_OK;
}
