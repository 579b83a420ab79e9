// Memory-bound glue of the separator forward: overlap-add, mask apply, magnitude,
// attentive statistics pooling, L2 normalise, DPRNN segmentation, FiLM combine and
// the [N,C,T] <-> [N,T,C] boundary transpose.  All are single-pass, coalesced
// along the channel (fastest) axis.
#include "ps_common.cuh"

namespace ps {

__device__ __forceinline__ float constrain(float v, int mode) {
  if (mode == 1) return (v != v) ? v : fminf(fmaxf(v, -1.f), 1.f);  // torch.clamp keeps NaN
  if (mode == 2) return 1.f / (1.f + expf(-v));
  return v;
}

// gather form of overlap-add: deterministic, no atomics.  y[b,j] = sum_t frames[b,t,j-t*hop]
__global__ void __launch_bounds__(256) ola_kernel(const float* __restrict__ frames, int64_t T, int64_t win, int64_t hop,
                                                  const float* __restrict__ wsum, int constraint, float* __restrict__ y,
                                                  int64_t out_len) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = blockIdx.y;
  pdl_trigger();
  pdl_wait();  // the synthesis GEMM's frames are complete
  if (j >= out_len) return;
  int64_t t_hi = j / hop;
  if (t_hi > T - 1) t_hi = T - 1;
  int64_t t_lo = (j - win + hop) / hop;  // ceil((j - win + 1) / hop) for j-win+1 > 0
  if (j - win + 1 <= 0) t_lo = 0;
  const float* fb = frames + b * T * win;
  float acc = 0.f;
  for (int64_t t = t_lo; t <= t_hi; ++t) acc += fb[t * win + (j - t * hop)];
  if (wsum) {
    float w = wsum[j];
    if (w > 1e-10f) acc = acc / w;
  }
  y[b * out_len + j] = constrain(acc, constraint);
}

__global__ void __launch_bounds__(256) mask_apply_kernel(const float* __restrict__ f, const float* __restrict__ m,
                                                         float* __restrict__ y, int64_t n_rows, int64_t C, int act,
                                                         int is_complex) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (!is_complex) {
    if (i >= n_rows * C) return;
    y[i] = f[i] * apply_act(m[i], act, 0.f);
  } else {
    const int64_t F = C / 2;
    if (i >= n_rows * F) return;
    const int64_t r = i / F, k = i % F;
    const float a = f[r * C + k], bq = f[r * C + F + k];
    const float c = apply_act(m[r * C + k], act, 0.f), dq = apply_act(m[r * C + F + k], act, 0.f);
    y[r * C + k] = a * c - bq * dq;
    y[r * C + F + k] = a * dq + bq * c;
  }
}

__global__ void __launch_bounds__(256) magnitude_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        int64_t n_rows, int64_t F, int drop_first, int log1p_) {
  const int64_t Fo = F - drop_first;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * Fo) return;
  const int64_t r = i / Fo, k = i % Fo + drop_first;
  const float re = x[r * 2 * F + k], im = x[r * 2 * F + F + k];
  float mag = sqrtf(re * re + im * im + 1e-8f);
  if (log1p_) mag = log1pf(mag);
  y[i] = mag;
}

// SpecAugment (lobe/trivial.py:307-335; torchaudio mask_along_axis): one [start, end) band on the channel axis and one on
// the frame axis, the same for every item; the bounds are READ FROM DEVICE MEMORY so a captured CUDA graph replays with
// the bands the host drew for this call.
__global__ void __launch_bounds__(256) band_fill_kernel(float* __restrict__ x, int64_t n, int64_t T, int64_t C,
                                                        const int32_t* __restrict__ bounds, float value) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = i % C, t = (i / C) % T;
  const int32_t c0 = __ldg(bounds), c1 = __ldg(bounds + 1), t0 = __ldg(bounds + 2), t1 = __ldg(bounds + 3);
  if ((c >= c0 && c < c1) || (t >= t0 && t < t1)) x[i] = value;
}

// grid (ceil(C/32), batch); block (32 channels, 8 frame lanes).  Three sweeps over the
// [T, 32] column strip (L2 resident): max, exp-sum + weighted mean, weighted variance —
// the same two-stage definition the reference uses (pooling.py:120-126).
__global__ void __launch_bounds__(256) asp_pool_kernel(const float* __restrict__ x, const float* __restrict__ logits,
                                                       int64_t T, int64_t C, float* __restrict__ out) {
  __shared__ float sa[8][33], sb[8][33];
  const int cx = threadIdx.x, ty = threadIdx.y;
  const int64_t c = (int64_t)blockIdx.x * 32 + cx;
  const int64_t b = blockIdx.y;
  const bool ok = c < C;
  const float* xb = x + b * T * C;
  const float* lb = logits + b * T * C;

  float mx = -INFINITY;
  if (ok)
    for (int64_t t = ty; t < T; t += 8) mx = fmaxf(mx, lb[t * C + c]);
  sa[ty][cx] = mx;
  __syncthreads();
  mx = sa[0][cx];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, sa[i][cx]);
  __syncthreads();

  float se = 0.f, sx = 0.f;
  if (ok)
    for (int64_t t = ty; t < T; t += 8) {
      float e = expf(lb[t * C + c] - mx);
      se += e;
      sx = fmaf(e, xb[t * C + c], sx);
    }
  sa[ty][cx] = se;
  sb[ty][cx] = sx;
  __syncthreads();
  se = 0.f; sx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { se += sa[i][cx]; sx += sb[i][cx]; }
  __syncthreads();
  const float inv = 1.f / se;
  const float mean = sx * inv;

  float sv = 0.f;
  if (ok)
    for (int64_t t = ty; t < T; t += 8) {
      float wgt = expf(lb[t * C + c] - mx) * inv;
      float dlt = xb[t * C + c] - mean;
      sv = fmaf(wgt, dlt * dlt, sv);
    }
  sa[ty][cx] = sv;
  __syncthreads();
  if (ty == 0 && ok) {
    sv = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sv += sa[i][cx];
    out[b * 2 * C + c] = mean;
    out[b * 2 * C + C + c] = sqrtf(fmaxf(sv, 1e-12f));
  }
}

__global__ void __launch_bounds__(256) l2normalize_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          int64_t rows, int64_t E) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int64_t e = lane; e < E; e += 32) s = fmaf(x[r * E + e], x[r * E + e], s);
  const float nrm = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  for (int64_t e = lane; e < E; e += 32) y[r * E + e] = x[r * E + e] / nrm;
}

// seg[b,q,k,:] from x[b,t,:].  VEC = 4: a thread moves 16 bytes and a 256-thread CTA SEG_ROWS rows (the first version ran
// one CTA of C threads per 512-byte row: 646,400 CTAs at cfg3, 1.1 TB/s).
constexpr int SEG_ROWS = 32;

template <int VEC>
__global__ void __launch_bounds__(256) segment_kernel(const float* __restrict__ x, float* __restrict__ seg, int64_t T,
                                                      int64_t C, int64_t K, int64_t S, int overlap) {
  const int64_t b = blockIdx.y;
  const int cv = (int)(C / VEC);                      // vectors per row
  const int n = SEG_ROWS * cv;                        // vectors this CTA moves
  const int64_t row0 = (int64_t)blockIdx.x * SEG_ROWS;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t row = row0 + i / cv;  // q*K + k
    if (row >= S * K) break;
    const int c = (int)(i % cv) * VEC;
    const int64_t q = row / K, k = row % K;
    const int64_t t = overlap ? q * (K / 2) + k - K / 2 : row;
    float* o = seg + (b * S * K + row) * C + c;
    const bool in = t >= 0 && t < T;
    if constexpr (VEC == 4) {
      *reinterpret_cast<float4*>(o) = in ? __ldg(reinterpret_cast<const float4*>(x + (b * T + t) * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      *o = in ? __ldg(x + (b * T + t) * C + c) : 0.f;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) merge_kernel(const float* __restrict__ seg, float* __restrict__ y, int64_t T,
                                                    int64_t C, int64_t K, int64_t S, int overlap) {
  const int64_t b = blockIdx.y;
  const int cv = (int)(C / VEC);
  const int n = SEG_ROWS * cv;
  const int64_t t0 = (int64_t)blockIdx.x * SEG_ROWS;
  const float* sb = seg + b * S * K * C;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t t = t0 + i / cv;
    if (t >= T) break;
    const int c = (int)(i % cv) * VEC;
    float* o = y + (b * T + t) * C + c;
    if (overlap) {
      const int64_t h = K / 2;
      const int64_t j1 = (t + h) / K, k1 = (t + h) % K;  // even stream, left h dropped
      const int64_t j2 = t / K, k2 = t % K;              // odd stream
      const float* p1 = sb + ((2 * j1) * K + k1) * C + c;
      const float* p2 = sb + ((2 * j2 + 1) * K + k2) * C + c;
      if constexpr (VEC == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p1)), d = __ldg(reinterpret_cast<const float4*>(p2));
        *reinterpret_cast<float4*>(o) = make_float4((a.x + d.x) / 2.f, (a.y + d.y) / 2.f, (a.z + d.z) / 2.f, (a.w + d.w) / 2.f);
      } else {
        *o = (__ldg(p1) + __ldg(p2)) / 2.f;
      }
    } else {
      if constexpr (VEC == 4) *reinterpret_cast<float4*>(o) = __ldg(reinterpret_cast<const float4*>(sb + t * C + c));
      else *o = __ldg(sb + t * C + c);
    }
  }
}

__global__ void __launch_bounds__(256) film_combine_kernel(const float* __restrict__ sbuf, const float* __restrict__ xn,
                                                           float* __restrict__ y, int64_t rows, int64_t C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int64_t r = i / C, c = i % C;
  y[i] = fmaf(sbuf[r * 2 * C + c], xn[i], sbuf[r * 2 * C + C + c]);
}

// [batch, R, C] -> [batch, C, R] through a padded 32x32 shared tile
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t R,
                                                        int64_t C) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const float* xb = x + b * R * C;
  float* yb = y + b * R * C;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = xb[r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) yb[c * R + r] = tile[threadIdx.x][i];
  }
}

}  // namespace ps

using ps::cdiv;

extern "C" int ps_ola(const float* frames, int64_t batch, int64_t T, int64_t win, int64_t hop, const float* wsum,
                      int32_t constraint, float* y, void* stream) {
  PS_REQUIRE(frames && y && batch > 0 && T > 0 && win > 0 && hop > 0 && constraint >= 0 && constraint <= 2);
  if (batch > 65535) return PS_ERR_UNSUPPORTED;
  const int64_t out_len = (T - 1) * hop + win;
  dim3 grid((unsigned)cdiv(out_len, 256), (unsigned)batch);
  cudaError_t le = ps::launch_pdl(ps::ola_kernel, grid, dim3(256), 0, (cudaStream_t)stream, frames, T, win, hop, wsum, (int)constraint, y, out_len);
  if (le != cudaSuccess) { ps::set_cuda_error(le, "ola_kernel"); return PS_ERR_CUDA; }
  PS_CHECK_LAUNCH("ola_kernel");
  return PS_OK;
}

extern "C" int ps_mask_apply(const float* feats, const float* mask, float* y, int64_t n_rows, int64_t C, int32_t act,
                             int32_t is_complex, void* stream) {
  PS_REQUIRE(feats && mask && y && n_rows > 0 && C > 0);
  PS_REQUIRE(act == PS_ACT_NONE || act == PS_ACT_RELU || act == PS_ACT_SIGMOID);
  if (is_complex) PS_REQUIRE(C % 2 == 0);
  const int64_t n = is_complex ? n_rows * (C / 2) : n_rows * C;
  ps::mask_apply_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(feats, mask, y, n_rows, C, act,
                                                                                 is_complex);
  PS_CHECK_LAUNCH("mask_apply_kernel");
  return PS_OK;
}

extern "C" int ps_magnitude(const float* x, float* y, int64_t n_rows, int64_t F, int32_t drop_first, int32_t log1p_,
                            void* stream) {
  PS_REQUIRE(x && y && n_rows > 0 && F > (drop_first ? 1 : 0));
  const int64_t n = n_rows * (F - (drop_first ? 1 : 0));
  ps::magnitude_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n_rows, F, drop_first ? 1 : 0,
                                                                                log1p_);
  PS_CHECK_LAUNCH("magnitude_kernel");
  return PS_OK;
}

extern "C" int ps_band_fill(float* x, int64_t batch, int64_t T, int64_t C, const int32_t* bounds, float value, void* stream) {
  PS_REQUIRE(x && bounds && batch > 0 && T > 0 && C > 0);
  const int64_t n = batch * T * C;
  ps::band_fill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, T, C, bounds, value);
  PS_CHECK_LAUNCH("band_fill_kernel");
  return PS_OK;
}

extern "C" int ps_asp_pool(const float* x, const float* logits, int64_t batch, int64_t T, int64_t C, float* out,
                           void* stream) {
  PS_REQUIRE(x && logits && out && batch > 0 && T > 0 && C > 0);
  if (batch > 65535) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)cdiv(C, 32), (unsigned)batch), block(32, 8);
  ps::asp_pool_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, logits, T, C, out);
  PS_CHECK_LAUNCH("asp_pool_kernel");
  return PS_OK;
}

extern "C" int ps_l2normalize(const float* x, float* y, int64_t rows, int64_t E, void* stream) {
  PS_REQUIRE(x && y && rows > 0 && E > 0);
  ps::l2normalize_kernel<<<(unsigned)cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, y, rows, E);
  PS_CHECK_LAUNCH("l2normalize_kernel");
  return PS_OK;
}

extern "C" int ps_segment(const float* x, float* seg, int64_t batch, int64_t T, int64_t C, int64_t K, int64_t S,
                          int32_t overlap, void* stream) {
  PS_REQUIRE(x && seg && batch > 0 && T > 0 && C > 0 && K > 0 && S > 0);
  if (overlap) PS_REQUIRE(K % 2 == 0);
  if (batch > 65535) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)cdiv(S * K, ps::SEG_ROWS), (unsigned)batch);
  const bool vec = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(seg)) & 15) == 0;
  if (vec) ps::segment_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(x, seg, T, C, K, S, overlap);
  else ps::segment_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, seg, T, C, K, S, overlap);
  PS_CHECK_LAUNCH("segment_kernel");
  return PS_OK;
}

extern "C" int ps_merge(const float* seg, float* y, int64_t batch, int64_t T, int64_t C, int64_t K, int64_t S,
                        int32_t overlap, void* stream) {
  PS_REQUIRE(seg && y && batch > 0 && T > 0 && C > 0 && K > 0 && S > 0);
  if (overlap) PS_REQUIRE(K % 2 == 0 && 2 * ((T - 1 + K / 2) / K) < S && 2 * ((T - 1) / K) + 1 < S);
  else PS_REQUIRE(T <= S * K);
  if (batch > 65535) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)cdiv(T, ps::SEG_ROWS), (unsigned)batch);
  const bool vec = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(seg)) & 15) == 0;
  if (vec) ps::merge_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(seg, y, T, C, K, S, overlap);
  else ps::merge_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(seg, y, T, C, K, S, overlap);
  PS_CHECK_LAUNCH("merge_kernel");
  return PS_OK;
}

extern "C" int ps_film_combine(const float* sb, const float* xn, float* y, int64_t rows, int64_t C, void* stream) {
  PS_REQUIRE(sb && xn && y && rows > 0 && C > 0);
  ps::film_combine_kernel<<<(unsigned)cdiv(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(sb, xn, y, rows, C);
  PS_CHECK_LAUNCH("film_combine_kernel");
  return PS_OK;
}

extern "C" int ps_transpose(const float* x, float* y, int64_t batch, int64_t R, int64_t C, void* stream) {
  PS_REQUIRE(x && y && batch > 0 && R > 0 && C > 0);
  if (batch > 65535 || cdiv(R, 32) > 65535) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(R, 32), (unsigned)batch), block(32, 8);
  ps::transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, R, C);
  PS_CHECK_LAUNCH("transpose_kernel");
  return PS_OK;
}
