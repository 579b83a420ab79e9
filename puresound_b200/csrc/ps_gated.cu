// Gated product of the GatedTCN block (reference conv_tasnet.py:141-205):
//
//     y = PReLU(norm_l(L)) * sigmoid(PReLU(norm_r(R)))
//
// L, R are the RAW outputs of the two dilated convolutions; the norms arrive folded exactly as the GEMM prologues take
// them (PS_PRO_AFFINE: gLN / gGN / bN1d as per-item per-channel scale, shift; PS_PRO_ROWNORM: cLN from row statistics),
// so neither normalised tensor is ever written.  With b == NULL the kernel is the plain transform y = act(pro(a)) with a
// strided output - used for the FiLM-conditioned right branch (x_r = scale * x + bias, conv_tasnet.py:197-200) written
// into the zero-padded conv input.  One read of each operand, one write: HBM-bound.
#include "ps_common.cuh"

namespace ps {

struct GatedSide {
  const float* x; int64_t bs, ms, rs;  // batch / mid / row strides
  int mode, act; const float* pa; const float* pb; int64_t pbs; const float* rowstats; const float* slope_p;
};

template <int V>
__device__ __forceinline__ float4 gated_ld(const float* p) {
  if constexpr (V == 4) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), 0.f, 0.f, 0.f);
}

template <int V>
__device__ __forceinline__ float4 gated_load(const GatedSide& s, float slope, int64_t b, int64_t m, int64_t mid, int64_t r, int64_t rows,
                                             int c) {
  float4 v = gated_ld<V>(s.x + b * s.bs + m * s.ms + r * s.rs + c);
  if (s.mode != PS_PRO_NONE) {
    const int64_t po = (s.mode == PS_PRO_AFFINE ? b * s.pbs : 0) + c;
    const float4 a = gated_ld<V>(s.pa + po);
    const float4 sh = gated_ld<V>(s.pb + po);
    if (s.mode == PS_PRO_ROWNORM) {
      const float2 st = __ldg(reinterpret_cast<const float2*>(s.rowstats + ((b * mid + m) * rows + r) * 2));
      v.x = (v.x - st.x) * st.y; v.y = (v.y - st.x) * st.y; v.z = (v.z - st.x) * st.y; v.w = (v.w - st.x) * st.y;
    }
    v.x = fmaf(v.x, a.x, sh.x); v.y = fmaf(v.y, a.y, sh.y); v.z = fmaf(v.z, a.z, sh.z); v.w = fmaf(v.w, a.w, sh.w);
  }
  v.x = apply_act(v.x, s.act, slope); v.y = apply_act(v.y, s.act, slope);
  v.z = apply_act(v.z, s.act, slope); v.w = apply_act(v.w, s.act, slope);
  return v;
}

__device__ __forceinline__ float gated_sigmoid(float v) { return 1.f / (1.f + expf(-v)); }

// V = 4: float4 per thread (C % 4 == 0, 16-byte aligned operands and strides); V = 1: scalar fallback for odd shapes
template <int V>
__global__ void __launch_bounds__(256) gated_kernel(const GatedSide A, const GatedSide B, float* __restrict__ y, int64_t ybs,
                                                    int64_t yms, int64_t yrs, int64_t batch, int64_t mid, int64_t rows, int C4) {
  const float sa = A.slope_p ? __ldg(A.slope_p) : 0.f;
  const float sb = (B.x && B.slope_p) ? __ldg(B.slope_p) : 0.f;
  const int64_t n_rows = batch * mid * rows;
  if (C4 <= (int)blockDim.x && n_rows < (1LL << 31)) {
    // Fast path: a thread keeps its column and walks rows - the (column, row-in-CTA) split is computed once, a row index is
    // taken apart with two 32-bit divisions.  The generic loop below pays six 64-bit divisions / remainders per 16-byte
    // element: as the transform copy of the U-Net shells (tap buffers of tse_unet_tcn / DPCRN / DPARN) it was 20-28 % of
    // those recipes' steps, issue-bound (run 53).
    const uint32_t rpb = blockDim.x / (uint32_t)C4;  // rows per CTA pass
    const uint32_t rl = threadIdx.x / (uint32_t)C4;
    const int c = (int)(threadIdx.x - rl * (uint32_t)C4) * V;
    if (rl >= rpb) return;
    const uint32_t nr = (uint32_t)n_rows, rows32 = (uint32_t)rows, mid32 = (uint32_t)mid;
    for (uint32_t R = blockIdx.x * rpb + rl; R < nr; R += gridDim.x * rpb) {
      const uint32_t bm = R / rows32, r = R - bm * rows32;
      const uint32_t b = bm / mid32, m = bm - b * mid32;
      float4 v = gated_load<V>(A, sa, b, m, mid, r, rows, c);
      if (B.x) {
        const float4 g = gated_load<V>(B, sb, b, m, mid, r, rows, c);
        v.x *= gated_sigmoid(g.x); v.y *= gated_sigmoid(g.y); v.z *= gated_sigmoid(g.z); v.w *= gated_sigmoid(g.w);
      }
      float* yp = y + (int64_t)b * ybs + (int64_t)m * yms + (int64_t)r * yrs + c;
      if constexpr (V == 4) *reinterpret_cast<float4*>(yp) = v;
      else *yp = v.x;
    }
    return;
  }
  const int64_t total = n_rows * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * V;
    const int64_t br = i / C4, r = br % rows, bm = br / rows, m = bm % mid, b = bm / mid;
    float4 v = gated_load<V>(A, sa, b, m, mid, r, rows, c);
    if (B.x) {
      const float4 g = gated_load<V>(B, sb, b, m, mid, r, rows, c);
      v.x *= gated_sigmoid(g.x); v.y *= gated_sigmoid(g.y); v.z *= gated_sigmoid(g.z); v.w *= gated_sigmoid(g.w);
    }
    if constexpr (V == 4) *reinterpret_cast<float4*>(y + b * ybs + m * yms + r * yrs + c) = v;
    else y[b * ybs + m * yms + r * yrs + c] = v.x;
  }
}

static bool gated_side_ok(const GatedSide& s, int64_t C) {
  if (!s.x || s.rs < C) return false;
  if (s.mode != PS_PRO_NONE && s.mode != PS_PRO_AFFINE && s.mode != PS_PRO_ROWNORM) return false;
  if (s.mode != PS_PRO_NONE && (!s.pa || !s.pb)) return false;
  if (s.mode == PS_PRO_ROWNORM && (!s.rowstats || (reinterpret_cast<uintptr_t>(s.rowstats) & 7))) return false;
  if (s.act == PS_ACT_PRELU && !s.slope_p) return false;
  return true;
}
static bool gated_side_vec(const GatedSide& s) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(s.x) || (s.bs & 3) || (s.ms & 3) || (s.rs & 3)) return false;
  if (s.mode != PS_PRO_NONE && (!al(s.pa) || !al(s.pb) || (s.pbs & 3))) return false;
  return true;
}

}  // namespace ps

extern "C" int ps_gated(const ps_gated_t* dp, void* stream) {
  PS_REQUIRE(dp != nullptr);
  const ps_gated_t& d = *dp;
  PS_REQUIRE(d.batch > 0 && d.rows > 0 && d.C > 0 && d.y && d.y_row_stride >= d.C);
  const int64_t mid = d.mid > 0 ? d.mid : 1;
  const ps::GatedSide A{d.a, d.a_batch_stride, d.a_mid_stride, d.a_row_stride, d.a_mode, d.a_act, d.a_pa, d.a_pb, d.a_pro_batch_stride, d.a_rowstats, d.a_slope};
  const ps::GatedSide B{d.b, d.b_batch_stride, d.b_mid_stride, d.b_row_stride, d.b_mode, d.b_act, d.b_pa, d.b_pb, d.b_pro_batch_stride, d.b_rowstats, d.b_slope};
  PS_REQUIRE(ps::gated_side_ok(A, d.C));
  if (d.b) PS_REQUIRE(ps::gated_side_ok(B, d.C));
  const bool vec = d.C % 4 == 0 && (reinterpret_cast<uintptr_t>(d.y) & 15) == 0 && (d.y_batch_stride & 3) == 0 &&
                   (d.y_mid_stride & 3) == 0 && (d.y_row_stride & 3) == 0 && ps::gated_side_vec(A) && (!d.b || ps::gated_side_vec(B));
  const int64_t cols = vec ? d.C / 4 : d.C;
  const int64_t total = d.batch * mid * d.rows * cols;
  // fast path (cols <= 256): a CTA pass covers 256 / cols rows; else one element per thread and pass
  int64_t blocks = cols <= 256 ? ps::cdiv(d.batch * mid * d.rows, 256 / cols) : ps::cdiv(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride: 16 CTAs of 256 threads per SM
  if (vec)
    ps::gated_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A, B, d.y, d.y_batch_stride, d.y_mid_stride, d.y_row_stride, d.batch, mid, d.rows, (int)cols);
  else
    ps::gated_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A, B, d.y, d.y_batch_stride, d.y_mid_stride, d.y_row_stride, d.batch, mid, d.rows, (int)cols);
  PS_CHECK_LAUNCH("gated_kernel");
  return PS_OK;
}
