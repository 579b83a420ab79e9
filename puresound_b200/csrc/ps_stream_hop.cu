// One hop of every concurrent stream of the causal Conv-TasNet as ONE persistent cooperative kernel (sm_100a).
//
// Round 1 ran a hop as a CUDA graph of 125 kernels (per block: GEMM, ring + depthwise step, GEMM, cLN row kernel, GEMM):
// 1.99 ms at 256 streams - each [256 x 512 x 512] GEMM occupied 2 CTA pairs of the 74, each node cost a launch gap - and
// 0.68 ms for one stream.  With per-frame norms (cLN / eval-BatchNorm) nothing in a hop needs more than the frame itself and
// each stream's dilation-history ring, so the whole stack is a fixed sequence of ~100 small phases.  Here one grid of one
// CTA per SM walks all of them, separated by a grid barrier (sense-reversing, two words in global memory):
//
//   push        frame[s] = [hist[s] | chunk[s]], hist <- its tail                              (a CTA per stream)
//   encoder     feats = frame W_enc^T (ReLU)                                                    (GEMM phase)
//   per block   A  u1 = x W_in^T (+ per-stream speaker bias)                                    (GEMM phase)
//               B  PReLU(norm1(u1)) -> ring push -> causal dilated taps -> PReLU(norm2(.))      (a CTA per stream)
//               C  u3 = u2 W_pw^T + b                                                           (GEMM phase)
//               E  x += PReLU(norm3(u3)) W_out^T + b   (norm3 + PReLU on the operand rows as they are staged)
//   decoder     frames = (feats * act(x)) W_dec + overlap-add emit, step counter += 1
//
// GEMM phase: the [S x M] output is cut into tiles of 16 streams x tcc channels (tcc in 4..64, chosen so that there is about
// one tile per CTA: 128 tiles of 16 x 64 at 256 streams, 128 tiles of 1 x 4 for a single stream - every SM gets a slice of
// the weights whatever the stream count); a CTA stages its 16 operand rows in shared memory (per-row norm statistics over K
// computed right there), each warp takes 4 output channels at a time: lanes stride over k, 64 FMAs per 4 coalesced weight
// loads + 16 conflict-free shared loads, then a 62-shuffle transposing reduction.  Exact fp32 (FFMA), weights read in the
// reference's own [M, K] layout - 77 MB per hop in total, each byte by exactly one CTA.
//
// The reference has no streaming Conv-TasNet; the API pattern is StreamingSkiM's (skim_inference.py:142-218),
// the oracle is the offline causal forward (conv_tasnet.py:11-90,218-377; lobe/norm.py:37-50).
#include <cooperative_groups.h>

#include <cuda_bf16.h>

#include "ps_common.cuh"

namespace ps {

constexpr int HP_THREADS = 256;
constexpr int HP_TR = 16;  // streams (rows) per GEMM tile

enum { HP_PRO_NONE = 0, HP_PRO_CLN = 1, HP_PRO_AFF = 2, HP_PRO_MASK = 3 };

struct HopGemm {
  const float* X; int64_t ldx;   // [S, K]
  const float* X2;               // MASK: second operand [S, K], same ld
  const float* W; int64_t ldw;   // [M, K] row-major
  const uint16_t* Wp;            // or NULL: bf16 split of W, [M][K] hi then [M][K] lo (tensor-core path, many streams)
  float* Y; int64_t ldy;         // [S, M]
  int M, K, pro, mask_act, relu;
  const float* pa; const float* pb; const float* slope; float eps;
  const float* bias;             // [M] or null
  const float* rowbias;          // [S, M] or null (per-stream speaker bias)
  const float* res; int64_t ldres;  // [S, M] or null (may alias Y: an element is read and written by the same thread)
};

// sense-reversing grid barrier: bar[0] = arrivals, bar[1] = generation (monotonic across launches, never reset).  `gen` is
// this CTA's copy of the generation: read ONCE at kernel entry (nobody can pass the first barrier before every CTA has read
// it, because every CTA reads before it arrives), then counted locally - one L2 round trip less per barrier.
__device__ __forceinline__ unsigned int hop_grid_gen(unsigned int* bar) {
  unsigned int gen;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
  return gen;
}
__device__ __forceinline__ void hop_grid_sync(unsigned int* bar, unsigned int& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // this CTA's writes (ordered before by the bar.sync above) are visible device-wide before it arrives
    if (atomicAdd(bar, 1u) == gridDim.x - 1) {
      bar[0] = 0;  // (ordered before the release below)
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1) : "memory");
    } else {
      unsigned int now;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(bar + 1) : "memory");
      } while (now == gen);
    }
  }
  ++gen;
  __syncthreads();
}

__device__ __forceinline__ float hop_block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect red from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < HP_THREADS / 32; ++i) t += red[i];
  return t;
}

// Shared-memory plan of a CTA: operand rows xs [16][Kmax], weight slab ws [64][Kmax] (filled by per-row cp.async.bulk copies,
// completion on the mbarrier `wbar`), K-split partial sums.
struct HopCtx {
  float* xs;
  float* ws;
  float* part;       // [8 warps][64]
  uint8_t* xb;       // bf16 operand rows: hi [16] then lo [16], row stride K*2 + 16 bytes (tensor-core path)
  float* pas;        // [Kmax] norm scale / gamma of the phase, [Kmax] shift / beta right behind it
  uint32_t wbar;     // shared-memory address of the weight-slab mbarrier
  uint32_t wphase;   // parity of the next completion to wait for
};

__device__ __forceinline__ void hop_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

// Request the weight slab of `tile` of GEMM phase g: rows [cgi*tcc, +tcc) of W, K floats each, one bulk copy per row.  The
// weights depend on nothing that happens in the hop, so a CTA asks for its NEXT phase's slab before it goes into the grid
// barrier: the L2 round trip of the copy hides behind the barrier and the operand staging.  (Caller: every thread, after a
// __syncthreads that ended the last use of ws.)
__device__ __forceinline__ void hop_w_issue(const HopGemm& g, int S, int tcc, int tile, const HopCtx& cx) {
  const int n_rg = (S + HP_TR - 1) / HP_TR, n_cg = (g.M + tcc - 1) / tcc;
  if (tile >= n_rg * n_cg) return;
  if (threadIdx.x < 32) {
    const int cgi = tile % n_cg;
    const int c0 = cgi * tcc;
    const int rows = (g.M - c0) < tcc ? (g.M - c0) : tcc;
    const uint32_t ws_u = (uint32_t)__cvta_generic_to_shared(cx.ws);
    if (g.Wp) {
      // bf16 hi rows then lo rows, each K*2 bytes, padded to a row stride of K*2 + 16 (conflict-free fragment loads)
      const uint32_t row_bytes = (uint32_t)g.K * 2u, stride = row_bytes + 16u;
      if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cx.wbar), "r"(row_bytes * 2u * (uint32_t)rows) : "memory");
      __syncwarp();
      for (int r = threadIdx.x; r < 2 * rows; r += 32) {
        const int half = r >= rows, rr = half ? r - rows : r;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ws_u + (uint32_t)(half * 64 + rr) * stride),
                     "l"(g.Wp + ((int64_t)half * g.M + c0 + rr) * g.K), "r"(row_bytes), "r"(cx.wbar)
                     : "memory");
      }
      return;
    }
    const uint32_t row_bytes = (uint32_t)g.K * 4u;
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cx.wbar), "r"(row_bytes * (uint32_t)rows) : "memory");
    __syncwarp();
    for (int r = threadIdx.x; r < rows; r += 32)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ws_u + (uint32_t)r * row_bytes),
                   "l"(g.W + (int64_t)(c0 + r) * g.ldw), "r"(row_bytes), "r"(cx.wbar)
                   : "memory");
  }
}

// One GEMM phase.  `pre`: the slab of this CTA's first tile (tile == blockIdx.x) was already requested by hop_w_issue.
__device__ __forceinline__ void hop_gemm(const HopGemm& g, const int S, HopCtx& cx, const int tcc, bool pre) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = g.K, M = g.M;
  const int n_rg = (S + HP_TR - 1) / HP_TR, n_cg = (M + tcc - 1) / tcc;
  const float slope = g.slope ? __ldg(g.slope) : 1.f;
  float* xs = cx.xs;
  const float* ws = cx.ws;
  for (int tile = blockIdx.x; tile < n_rg * n_cg; tile += gridDim.x) {
    const int rg = tile / n_cg, cgi = tile - rg * n_cg;  // consecutive CTAs share a row group: its operand rows hit L2
    const int row0 = rg * HP_TR;
    if (!pre) hop_w_issue(g, S, tcc, tile, cx);
    pre = false;
    // ---- stage the operand rows (K % 4 == 0, 16-byte aligned rows: checked by the launcher)
    const int K4 = K >> 2;
    for (int idx = tid; idx < HP_TR * K4; idx += HP_THREADS) {
      const int r = idx / K4, k4 = idx - r * K4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < S) {
        v = __ldcg(reinterpret_cast<const float4*>(g.X + (int64_t)(row0 + r) * g.ldx) + k4);  // written by other CTAs this launch: L2
        if (g.pro == HP_PRO_MASK) {
          const float4 m = __ldcg(reinterpret_cast<const float4*>(g.X2 + (int64_t)(row0 + r) * g.ldx) + k4);
          v.x *= apply_act(m.x, g.mask_act, 0.f); v.y *= apply_act(m.y, g.mask_act, 0.f);
          v.z *= apply_act(m.z, g.mask_act, 0.f); v.w *= apply_act(m.w, g.mask_act, 0.f);
        }
      }
      *reinterpret_cast<float4*>(xs + r * K + k4 * 4) = v;
    }
    if (g.pro == HP_PRO_CLN || g.pro == HP_PRO_AFF) {  // the norm's per-channel parameters ride the same round trip
      for (int k = tid; k < K; k += HP_THREADS) {
        cx.pas[k] = __ldg(g.pa + k);
        cx.pas[K + k] = __ldg(g.pb + k);
      }
    }
    __syncthreads();
    if (g.pro == HP_PRO_CLN || g.pro == HP_PRO_AFF) {
      // per-row norm + PReLU on the staged rows: warp w takes rows 2w, 2w+1 (cLN: mean / biased variance over the K channels,
      // lobe/norm.py:40-50; AFF: eval-BatchNorm folded to a per-channel affine)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (row0 + 2 * warp + h >= S) continue;
        float* xr = xs + (2 * warp + h) * K;
        float mean = 0.f, rstd = 1.f;
        if (g.pro == HP_PRO_CLN) {
          // one pass: sums of (x - pivot) and (x - pivot)^2 about the row's first element (no cancellation for a row with a
          // large common offset), mean / biased variance from them
          const float piv = xr[0];
          float s1 = 0.f, s2 = 0.f;
          for (int k = lane; k < K; k += 32) { const float dl = xr[k] - piv; s1 += dl; s2 = fmaf(dl, dl, s2); }
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          const float md = s1 / (float)K;
          mean = piv + md;
          rstd = 1.f / sqrtf(fmaxf(s2 / (float)K - md * md, 0.f) + g.eps);
        }
        for (int k = lane; k < K; k += 32) {
          float z = fmaf((xr[k] - mean) * rstd, cx.pas[k], cx.pas[K + k]);
          xr[k] = z > 0.f ? z : z * slope;
        }
      }
      __syncthreads();
    }
    // ---- the weight slab has landed
    hop_mbar_wait(cx.wbar, cx.wphase);
    cx.wphase ^= 1u;
    const int c_base = cgi * tcc;
    const int rows_w = (M - c_base) < tcc ? (M - c_base) : tcc;
    if (g.Wp) {
      // ---- tensor-core path (many streams): 3xBF16 split on mma.sync m16n8k16 - the 16 streams of the tile are the MMA's M,
      // a warp takes 8 output channels at a time, fp32 accumulate (hi*lo + lo*hi + hi*hi, ~2^-17 per product like the
      // offline GEMMs).  Operand rows: fp32 xs -> bf16 hi / lo rows in xb (stride K*2 + 16 bytes).
      const uint32_t stride = (uint32_t)K * 2u + 16u;
      for (int idx = tid; idx < HP_TR * (K >> 1); idx += HP_THREADS) {
        const int r = idx / (K >> 1), k2 = idx - r * (K >> 1);
        const float2 v = *reinterpret_cast<const float2*>(xs + r * K + 2 * k2);
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v.x, v.y);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v.x - __uint_as_float(hb << 16), v.y - __uint_as_float(hb & 0xFFFF0000u));
        *reinterpret_cast<uint32_t*>(cx.xb + r * stride + k2 * 4) = hb;
        *reinterpret_cast<uint32_t*>(cx.xb + (HP_TR + r) * stride + k2 * 4) = *reinterpret_cast<const uint32_t*>(&l2);
      }
      __syncthreads();
      const int gq = lane >> 2, tq = lane & 3;
      const uint8_t* wsb = reinterpret_cast<const uint8_t*>(ws);
      for (int o8 = warp; o8 * 8 < rows_w; o8 += HP_THREADS / 32) {
        float dacc[4] = {0.f, 0.f, 0.f, 0.f};
        // epilogue operands of this lane's four outputs, requested BEFORE the inner product (one round trip, overlapped)
        float eb[4] = {0.f, 0.f, 0.f, 0.f}, er[4] = {0.f, 0.f, 0.f, 0.f};
        {
          const int cc0 = o8 * 8 + tq * 2;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cq1 = cc0 + (e & 1), row = row0 + gq + (e >> 1) * 8;
            if (cq1 < rows_w && row < S) {
              const int ch = c_base + cq1;
              if (g.bias) eb[e] = __ldg(g.bias + ch);
              if (g.rowbias) eb[e] += __ldg(g.rowbias + (int64_t)row * M + ch);
              if (g.res) er[e] = __ldcg(g.res + (int64_t)row * g.ldres + ch);
            }
          }
        }
        const uint8_t* a_hi = cx.xb + gq * stride + tq * 4;
        const uint8_t* a_lo = a_hi + HP_TR * stride;
        const int wrow = o8 * 8 + gq;  // this lane's weight row (output channel within the slab); rows past the slab read junk, never stored
        const uint8_t* b_hi = wsb + (wrow < 64 ? wrow : 63) * stride + tq * 4;
        const uint8_t* b_lo = b_hi + 64 * stride;
#pragma unroll 4
        for (int k0 = 0; k0 < K; k0 += 16) {
          uint32_t ah[4], al[4], bh[2], bl[2];
          ah[0] = *reinterpret_cast<const uint32_t*>(a_hi + k0 * 2);
          ah[1] = *reinterpret_cast<const uint32_t*>(a_hi + 8 * stride + k0 * 2);
          ah[2] = *reinterpret_cast<const uint32_t*>(a_hi + k0 * 2 + 16);
          ah[3] = *reinterpret_cast<const uint32_t*>(a_hi + 8 * stride + k0 * 2 + 16);
          al[0] = *reinterpret_cast<const uint32_t*>(a_lo + k0 * 2);
          al[1] = *reinterpret_cast<const uint32_t*>(a_lo + 8 * stride + k0 * 2);
          al[2] = *reinterpret_cast<const uint32_t*>(a_lo + k0 * 2 + 16);
          al[3] = *reinterpret_cast<const uint32_t*>(a_lo + 8 * stride + k0 * 2 + 16);
          bh[0] = *reinterpret_cast<const uint32_t*>(b_hi + k0 * 2);
          bh[1] = *reinterpret_cast<const uint32_t*>(b_hi + k0 * 2 + 16);
          bl[0] = *reinterpret_cast<const uint32_t*>(b_lo + k0 * 2);
          bl[1] = *reinterpret_cast<const uint32_t*>(b_lo + k0 * 2 + 16);
#define HP_MMA(A, B)                                                                                                           \
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"     \
               : "+f"(dacc[0]), "+f"(dacc[1]), "+f"(dacc[2]), "+f"(dacc[3])                                                     \
               : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B[0]), "r"(B[1]))
          HP_MMA(al, bh);
          HP_MMA(ah, bl);
          HP_MMA(ah, bh);
#undef HP_MMA
        }
        // D fragment: dacc[0..1] = (row gq, channels 2 tq, 2 tq + 1), dacc[2..3] = (row gq + 8, same channels)
        const int cc = o8 * 8 + tq * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cq1 = cc + (e & 1), row = row0 + gq + (e >> 1) * 8;
          if (cq1 < rows_w && row < S) {
            const int ch = c_base + cq1;
            float y = dacc[e] + eb[e];
            if (g.relu) y = (y != y) ? y : fmaxf(y, 0.f);
            y += er[e];
            g.Y[(int64_t)row * g.ldy + ch] = y;
          }
        }
      }
      __syncthreads();  // the operand rows and the weight slab are free for the next tile
      continue;
    }
    // ---- warp per (channel quad, K slice): with fewer than 8 quads in the slab the K range is split over the idle warps
    const int nq = (rows_w + 3) >> 2;
    int ks = 1;
    while (ks * 2 * nq <= HP_THREADS / 32 && (K / (ks * 2)) % 32 == 0) ks *= 2;  // 1, 2, 4 or 8 K slices
    const int kslice = K / ks;
    auto quad_sums = [&](int q, int sl, float (&acc)[64]) {  // lane L ends with the sums of two (channel, row) pairs in acc[0..1]
      const int cq = q * 4;  // first channel of the quad within the slab
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      const float* w0 = ws + cq * K;
      const bool ok1 = cq + 1 < rows_w, ok2 = cq + 2 < rows_w, ok3 = cq + 3 < rows_w;
      for (int k = sl * kslice + lane; k < (sl + 1) * kslice; k += 32) {
        float wv[4];
        wv[0] = w0[k];
        wv[1] = ok1 ? w0[K + k] : 0.f;
        wv[2] = ok2 ? w0[2 * K + k] : 0.f;
        wv[3] = ok3 ? w0[3 * K + k] : 0.f;
#pragma unroll
        for (int r = 0; r < HP_TR; ++r) {
          const float xv = xs[r * K + k];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i * 16 + r] = fmaf(wv[i], xv, acc[i * 16 + r]);
        }
      }
      // transposing reduction over the 32 lanes: 62 shuffles leave lane L with the full sums of two (channel, row) pairs
#pragma unroll
      for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < n) {
            const float send = up ? acc[i] : acc[i + n];
            const float keep = up ? acc[i + n] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
      }
    };
    // epilogue operands of this lane's two outputs of quad q (requested before the inner product: one overlapped round trip)
    auto quad_fetch = [&](int q, float (&eb)[2], float (&er)[2]) {
      const int cq = q * 4;
      const int ci = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
      const int rb = ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
      const int ch = c_base + cq + ci;
      eb[0] = eb[1] = er[0] = er[1] = 0.f;
      if (cq + ci < rows_w) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int row = row0 + rb + j;
          if (row < S) {
            if (g.bias) eb[j] = __ldg(g.bias + ch);
            if (g.rowbias) eb[j] += __ldg(g.rowbias + (int64_t)row * M + ch);
            if (g.res) er[j] = __ldcg(g.res + (int64_t)row * g.ldres + ch);
          }
        }
      }
    };
    auto quad_store = [&](int q, float a0, float a1, const float (&eb)[2], const float (&er)[2]) {
      const int cq = q * 4;
      const int ci = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
      const int rb = ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
      const int ch = c_base + cq + ci;
      if (cq + ci < rows_w) {
        const float av[2] = {a0, a1};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int row = row0 + rb + j;
          if (row < S) {
            float y = av[j] + eb[j];
            if (g.relu) y = (y != y) ? y : fmaxf(y, 0.f);
            y += er[j];
            g.Y[(int64_t)row * g.ldy + ch] = y;
          }
        }
      }
    };
    if (ks == 1) {
      for (int q = warp; q < nq; q += HP_THREADS / 32) {
        float acc[64], eb[2], er[2];
        quad_fetch(q, eb, er);
        quad_sums(q, 0, acc);
        quad_store(q, acc[0], acc[1], eb, er);
      }
    } else {
      // one pass: warp u < nq * ks takes quad u % nq, K slice u / nq; the slices meet in shared memory
      const bool active = warp < nq * ks;
      const int q = warp % nq, sl = warp / nq;
      float a0 = 0.f, a1 = 0.f, eb[2] = {0.f, 0.f}, er[2] = {0.f, 0.f};
      if (active && sl == 0) quad_fetch(q, eb, er);
      if (active) {
        float acc[64];
        quad_sums(q, sl, acc);
        a0 = acc[0]; a1 = acc[1];
        cx.part[warp * 64 + lane * 2] = a0;
        cx.part[warp * 64 + lane * 2 + 1] = a1;
      }
      __syncthreads();
      if (active && sl == 0) {
        for (int o = 1; o < ks; ++o) {
          a0 += cx.part[(q + o * nq) * 64 + lane * 2];
          a1 += cx.part[(q + o * nq) * 64 + lane * 2 + 1];
        }
        quad_store(q, a0, a1, eb, er);
      }
    }
    __syncthreads();  // the operand rows and the weight slab are free for the next tile
  }
}

// tcc: the smallest channel-slab width (4..64) that gives at most one tile per CTA
__device__ __forceinline__ int hop_tcc(int S, int M) {
  const int n_rg = (S + HP_TR - 1) / HP_TR;
  int tcc = 4;
  while (tcc < 64 && n_rg * ((M + tcc - 1) / tcc) > (int)gridDim.x) tcc <<= 1;
  return tcc;
}

// row-wise norm + PReLU on channels held as v[i] = row[tid + i * HP_THREADS]; kind 0: cLN, 1: per-channel affine
template <int NV>
__device__ __forceinline__ void hop_row_norm(float (&v)[NV], int nv, int C, int kind, float eps, const float* a, const float* b, float slope, float* red) {
  float mean = 0.f, rstd = 1.f;
  if (kind == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) if (i < nv) s += v[i];
    mean = hop_block_sum(s, red) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) if (i < nv) { const float dl = v[i] - mean; q = fmaf(dl, dl, q); }
    rstd = 1.f / sqrtf(hop_block_sum(q, red) / (float)C + eps);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      const int c = threadIdx.x + i * HP_THREADS;
      float x = (kind == 0) ? (v[i] - mean) * rstd : v[i];
      x = fmaf(x, __ldg(a + c), __ldg(b + c));
      v[i] = x > 0.f ? x : x * slope;
    }
  }
}

constexpr int HP_NV = 8;  // channels per thread of the row phases: H <= 2048

#ifdef PS_EXPERIMENTS
// phase timeline of CTA 0 (globaltimer ns at every phase boundary), read back by ps_debug_hop_times - experiments build only
__device__ unsigned long long hop_times_dev[1024];
__device__ __forceinline__ void hop_stamp(int& n) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && n < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    hop_times_dev[n] = t;
  }
  ++n;
}
#define HP_STAMP() hop_stamp(n_stamp)
#else
#define HP_STAMP()
#endif

__global__ void __launch_bounds__(HP_THREADS, 1) stream_hop_kernel(const ps_stream_hop_t d, const int kmax) {
  extern __shared__ __align__(16) float hop_smem[];  // xs [16][kmax] | ws [64][kmax] | part [8][64]
  __shared__ float red[HP_THREADS / 32];
  __shared__ __align__(8) uint64_t wbar_s;
  const int tid = threadIdx.x;
  const int S = (int)d.streams, C = d.C, H = d.H, win = d.win, hop = d.hop;
  const int64_t step = *d.step;
  unsigned int gen = 0;
  if (tid == 0) gen = hop_grid_gen(d.barrier);  // (only thread 0 uses it)
#ifdef PS_EXPERIMENTS
  int n_stamp = 0;
#endif
  HP_STAMP();
  HopCtx cx;
  cx.xs = hop_smem;
  cx.ws = hop_smem + HP_TR * kmax;
  // ws holds either the fp32 slab [64][kmax] or the bf16 hi / lo slabs [128][kmax*2 + 16 bytes]: sized for the larger
  const int ws_bytes = (64 * kmax * 4) > (128 * (kmax * 2 + 16)) ? (64 * kmax * 4) : (128 * (kmax * 2 + 16));
  cx.part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(cx.ws) + ws_bytes);
  cx.xb = reinterpret_cast<uint8_t*>(cx.part + 8 * 64);
  cx.pas = reinterpret_cast<float*>(cx.xb + 2 * HP_TR * (kmax * 2 + 16));
  cx.wbar = (uint32_t)__cvta_generic_to_shared(&wbar_s);
  cx.wphase = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cx.wbar), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // GEMM phases of the hop, in order; each one's weight slab is requested as soon as the previous GEMM phase is done with ws
  HopGemm genc = {};
  genc.X = d.frame; genc.ldx = win; genc.W = d.w_enc; genc.ldw = win; genc.Y = d.feats; genc.ldy = C; genc.M = C; genc.K = win; genc.relu = d.enc_relu; genc.Wp = d.w_enc_p;
  const int tcc_c = hop_tcc(S, C), tcc_h = hop_tcc(S, H), tcc_w = hop_tcc(S, win);
  hop_w_issue(genc, S, tcc_c, blockIdx.x, cx);

  // ---- push: frame = [hist | chunk], hist <- last (win - hop) samples
  const int keep = win - hop;
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    float* f = d.frame + (int64_t)s * win;
    float* h = d.hist + (int64_t)s * (keep > 0 ? keep : 1);
    const float* c = d.chunk + (int64_t)s * hop;
    for (int i = tid; i < win; i += HP_THREADS) f[i] = (i < keep) ? h[i] : c[i - keep];
    __syncthreads();
    for (int i = tid; i < keep; i += HP_THREADS) h[i] = f[i + hop];
  }
  HP_STAMP(); hop_grid_sync(d.barrier, gen); HP_STAMP();

  // ---- encoder
  hop_gemm(genc, S, cx, tcc_c, true);

  auto gemm_a = [&](int j, const float* xin) {
    const ps_stream_hop_block_t& bk = d.blocks[j];
    HopGemm g = {};
    g.X = xin; g.ldx = C; g.W = bk.w_in; g.ldw = bk.w_in_ld; g.Y = d.u1; g.ldy = H; g.M = H; g.K = C; g.rowbias = bk.ebias; g.Wp = bk.w_in_p;
    return g;
  };
  auto gemm_c = [&](int j) {
    const ps_stream_hop_block_t& bk = d.blocks[j];
    HopGemm g = {};
    g.X = d.u2; g.ldx = H; g.W = bk.w_pw; g.ldw = H; g.Y = d.u3; g.ldy = H; g.M = H; g.K = H; g.bias = bk.b_pw; g.Wp = bk.w_pw_p;
    return g;
  };
  auto gemm_e = [&](int j, const float* xin) {
    const ps_stream_hop_block_t& bk = d.blocks[j];
    HopGemm g = {};
    g.X = d.u3; g.ldx = H; g.W = bk.w_out; g.ldw = H; g.Y = d.x; g.ldy = C; g.M = C; g.K = H; g.bias = bk.b_out;
    g.pro = d.norm_kind == 0 ? HP_PRO_CLN : HP_PRO_AFF; g.pa = bk.n3_a; g.pb = bk.n3_b; g.slope = bk.slope3; g.eps = d.eps;
    g.res = xin; g.ldres = C; g.Wp = bk.w_out_p;
    return g;
  };
  auto gemm_dec = [&](const float* mask) {
    HopGemm g = {};
    g.X = d.feats; g.X2 = mask; g.ldx = C; g.W = d.w_dec_t; g.ldw = C; g.Y = d.frame_out; g.ldy = win; g.M = win; g.K = C;
    g.pro = HP_PRO_MASK; g.mask_act = d.mask_act; g.Wp = d.w_dec_p;
    return g;
  };

  const float* xin = d.feats;  // block input: the encoder output for block 0, then the residual stream x
  if (d.n_blocks > 0) hop_w_issue(gemm_a(0, xin), S, tcc_h, blockIdx.x, cx);
  else hop_w_issue(gemm_dec(xin), S, tcc_w, blockIdx.x, cx);
  HP_STAMP(); hop_grid_sync(d.barrier, gen); HP_STAMP();

  for (int j = 0; j < d.n_blocks; ++j) {
    const ps_stream_hop_block_t& bk = d.blocks[j];
    // ---- A: u1 = x W_in^T (+ per-stream speaker bias)
    hop_gemm(gemm_a(j, xin), S, cx, tcc_h, true);
    hop_w_issue(gemm_c(j), S, tcc_h, blockIdx.x, cx);  // lands while the grid goes through B
    HP_STAMP(); hop_grid_sync(d.barrier, gen); HP_STAMP();
    // ---- B: per stream: PReLU(norm1(u1)) -> ring push -> causal dilated taps -> PReLU(norm2(.)) -> u2
    {
      const int P = bk.P, dil = bk.dilation;
      const int64_t RL = (int64_t)(P - 1) * dil + 1;
      const float s1 = __ldg(bk.slope1), s2 = __ldg(bk.slope2);
      const int nv = (H + HP_THREADS - 1 - tid) / HP_THREADS;  // channels tid, tid + 256, ... < H
      for (int s = blockIdx.x; s < S; s += gridDim.x) {
        float v[HP_NV], a[HP_NV], hist[HP_NV][2];
        float* ring = bk.ring + (int64_t)s * RL * H;
        const int64_t slot = step % RL;
        // the row and (P <= 3) its two history taps are requested together: the taps come from HBM (the rings of 256 streams
        // are 800 MB) and depend on nothing computed in this hop
#pragma unroll
        for (int i = 0; i < HP_NV; ++i) {
          v[i] = (i < nv) ? __ldcg(d.u1 + (int64_t)s * H + tid + i * HP_THREADS) : 0.f;
          hist[i][0] = hist[i][1] = 0.f;
          if (i < nv && P <= 3) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              const int64_t back = (int64_t)(P - 1 - p) * dil;
              if (p < P - 1 && back <= step) hist[i][p] = ring[((step - back) % RL) * H + tid + i * HP_THREADS];
            }
          }
        }
        hop_row_norm<HP_NV>(v, nv, H, d.norm_kind, d.eps, bk.n1_a, bk.n1_b, s1, red);
#pragma unroll
        for (int i = 0; i < HP_NV; ++i) {
          if (i < nv) {
            const int c = tid + i * HP_THREADS;
            ring[slot * H + c] = v[i];
            float acc = bk.dw_b ? __ldg(bk.dw_b + c) : 0.f;
            for (int p = 0; p < P; ++p) {
              const int64_t back = (int64_t)(P - 1 - p) * dil;
              float x;
              if (back == 0) x = v[i];
              else if (back > step) x = 0.f;  // causal zero padding before the stream started
              else if (P <= 3) x = hist[i][p < 2 ? p : 1];
              else x = ring[((step - back) % RL) * H + c];
              acc = fmaf(__ldg(bk.dw_w + c * P + p), x, acc);
            }
            a[i] = acc;
          } else {
            a[i] = 0.f;
          }
        }
        hop_row_norm<HP_NV>(a, nv, H, d.norm_kind, d.eps, bk.n2_a, bk.n2_b, s2, red);
#pragma unroll
        for (int i = 0; i < HP_NV; ++i) if (i < nv) d.u2[(int64_t)s * H + tid + i * HP_THREADS] = a[i];
      }
    }
    HP_STAMP(); hop_grid_sync(d.barrier, gen); HP_STAMP();
    // ---- C: u3 = u2 W_pw^T + b
    hop_gemm(gemm_c(j), S, cx, tcc_h, true);
    hop_w_issue(gemm_e(j, xin), S, tcc_c, blockIdx.x, cx);
    HP_STAMP(); hop_grid_sync(d.barrier, gen); HP_STAMP();
    // ---- E: x = PReLU(norm3(u3)) W_out^T + b + x_in
    hop_gemm(gemm_e(j, xin), S, cx, tcc_c, true);
    xin = d.x;
    if (j + 1 < d.n_blocks) hop_w_issue(gemm_a(j + 1, xin), S, tcc_h, blockIdx.x, cx);
    else hop_w_issue(gemm_dec(xin), S, tcc_w, blockIdx.x, cx);
    HP_STAMP(); hop_grid_sync(d.barrier, gen); HP_STAMP();
  }

  // ---- decoder: frames = (feats * act(mask)) W_dec ([win, C] = the transposed ConvTranspose1d weight)
  hop_gemm(gemm_dec(xin), S, cx, tcc_w, true);
  HP_STAMP(); hop_grid_sync(d.barrier, gen); HP_STAMP();
  // ---- overlap-add emit: acc += frame; out = constrain(acc[:hop]); acc <- shift left by hop
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    float* a = d.acc + (int64_t)s * win;
    float* tmp = hop_smem;
    for (int i = tid; i < win; i += HP_THREADS) tmp[i] = a[i] + __ldcg(d.frame_out + (int64_t)s * win + i);
    __syncthreads();
    for (int i = tid; i < win; i += HP_THREADS) {
      if (i < hop) {
        float v = tmp[i];
        if (d.constraint == 1) v = (v != v) ? v : fminf(fmaxf(v, -1.f), 1.f);
        else if (d.constraint == 2) v = 1.f / (1.f + expf(-v));
        d.out[(int64_t)s * hop + i] = v;
      }
      a[i] = (i + hop < win) ? tmp[i + hop] : 0.f;
    }
    __syncthreads();
  }
  // every phase that reads the step counter is behind a barrier: safe to advance it now
  if (blockIdx.x == 0 && tid == 0) *d.step = step + 1;
}

}  // namespace ps

#ifdef PS_EXPERIMENTS
extern "C" __attribute__((visibility("default"))) int ps_debug_hop_times(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, ps::hop_times_dev, sizeof(unsigned long long) * (size_t)n);
}
#endif

extern "C" int ps_stream_hop(const ps_stream_hop_t* dp, void* stream) {
  PS_REQUIRE(dp != nullptr);
  const ps_stream_hop_t& d = *dp;
  PS_REQUIRE(d.streams > 0 && d.C > 0 && d.H > 0 && d.win >= d.hop && d.hop > 0 && d.n_blocks >= 0);
  PS_REQUIRE(d.w_enc && d.w_dec_t && d.chunk && d.frame && d.frame_out && d.acc && d.out && d.step && d.feats && d.x && d.u1 && d.u2 && d.u3 && d.barrier);
  PS_REQUIRE(d.win == d.hop || d.hist);
  PS_REQUIRE(d.n_blocks == 0 || d.blocks);
  PS_REQUIRE(d.norm_kind == 0 || d.norm_kind == 1);
  if ((d.C & 3) || (d.H & 3) || (d.win & 3) || d.H > ps::HP_THREADS * ps::HP_NV) return PS_ERR_UNSUPPORTED;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!(al(d.frame) && al(d.feats) && al(d.x) && al(d.u1) && al(d.u2) && al(d.u3))) return PS_ERR_UNSUPPORTED;
  int kmax = d.C > d.H ? d.C : d.H;
  kmax = kmax > d.win ? kmax : d.win;
  const size_t ws_bytes = (size_t)64 * kmax * 4 > (size_t)128 * (kmax * 2 + 16) ? (size_t)64 * kmax * 4 : (size_t)128 * (kmax * 2 + 16);
  const size_t smem = (size_t)ps::HP_TR * kmax * 4 + ws_bytes + 8 * 64 * 4 + (size_t)2 * ps::HP_TR * (kmax * 2 + 16) + (size_t)2 * kmax * 4;
  if (smem > 220 * 1024) return PS_ERR_UNSUPPORTED;
  if ((d.C & 15) || (d.H & 15) || (d.win & 15)) { if (d.w_enc_p || d.w_dec_p) return PS_ERR_UNSUPPORTED; }  // the MMA path steps k by 16
  int dev = 0, sms = 0;
  if (int rc = ps::current_device(&dev)) return rc;
  if (int rc = ps::sm_count_of(dev, &sms)) return rc;
  static ps::SmemOnce<1> once;
  if (int rc = once.ensure(dev, 0, ps::stream_hop_kernel, 220 * 1024, "cudaFuncSetAttribute(stream_hop_kernel)")) return rc;
  // cooperative launch: the grid barrier needs every CTA resident (one per SM)
  void* args[] = {const_cast<ps_stream_hop_t*>(dp), &kmax};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(ps::stream_hop_kernel), dim3((unsigned)sms), dim3(ps::HP_THREADS), args, smem,
                                              (cudaStream_t)stream);
  if (e != cudaSuccess) { ps::set_cuda_error(e, "cudaLaunchCooperativeKernel(stream_hop_kernel)"); return PS_ERR_CUDA; }
  return PS_OK;
}
