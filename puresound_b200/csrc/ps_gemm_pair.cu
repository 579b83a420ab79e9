// CTA-pair tcgen05 / TMEM GEMM for the 1x1 convolutions of the TCN stack (sm_100a, cta_group::2).
//
//   Y[b,f,c] = epi( sum_k pro(X[b,f,k]) * W[c,k] ),  fp32 in HBM, fp32-grade result (3xBF16 split, see ps_gemm_tc.cu).
//
// Why a second kernel.  ncu on the single-CTA kernel (profiles/r01_tc_gemm_ncu_notes.md) showed the tensor pipe 48 %
// busy and the MMA thread waiting for the activation producers: with a 128-frame x 256-channel tile every activation
// tile is loaded and transformed (norm affine + PReLU + bf16 hi/lo split) once per 256-channel half, and the LSU data
// pipe (71 %) carries that twice plus an epilogue transpose through shared memory.  Here the operand roles are
// swapped and two SMs share one tile:
//
//   * A (M side, TMEM lanes)   = weights: 256 channels per MMA, 128 from each CTA of the pair;
//   * B (N side, TMEM columns) = activations: 128 frames per MMA, 64 transformed by each CTA - the tensor core reads
//     the other half straight from the peer's shared memory, so every activation element is transformed ONCE for all
//     (up to 512) output channels of the tile;
//   * D: each CTA holds [128 channels x 128 frames] per 256-channel block; two blocks (512 channels) use 256 TMEM
//     columns, so two tiles ping-pong in the 512 columns and the epilogue of one overlaps the MMAs of the next;
//   * the epilogue needs no transpose: a TMEM lane is a channel, a column is a frame, so for one frame a warp's 32
//     lanes are 32 consecutive channels = one coalesced 128-byte store (and residual load) into the [frames, channels]
//     tensor.
//
// Warp roles per CTA (576 threads): warp 0 weight loader (cp.async.bulk of the pre-packed, pre-swizzled image),
// warp 1 MMA issuer (leader CTA) / stage relay (peer CTA), warps 2-9 epilogue (two per TMEM lane quarter), warps
// 10-17 activation producers.
// Pipelines: 4 shared-memory stages of 32 k (full / peer-full / empty), 2 TMEM accumulator sets (tfull / tempty).
// The leader's MMA thread issues for both SMs; tcgen05.commit multicasts the stage-free and accumulator-ready
// arrivals to both CTAs; the peer reports its filled stages and drained accumulators with remote mbarrier arrives.
#include <stdlib.h>
#include <string.h>

#include "ps_tc_ptx.cuh"

// Bottleneck experiments (PS_PAIR_DBG: remove one stream of work at a time; results are garbage) exist only in builds with
// -DPS_EXPERIMENTS; the release library has no such switch and no mutable state.
#ifdef PS_EXPERIMENTS
#define PR_DBG(bit) ((dbg & (bit)) != 0)
#else
#define PR_DBG(bit) false
#endif

namespace ps {

constexpr int PR_FRAMES = 128;      // frames per pair tile (MMA N)
constexpr int PR_FR_CTA = 64;       // frames staged by each CTA
constexpr int PR_BK = 32;           // k per shared-memory stage (64-byte swizzle rows)
#ifndef PR_STAGES
#define PR_STAGES 4                 // shared-memory ring depth (4 x 40 KB at 512 channels; 5 measured no faster)
#endif
constexpr int PR_XPART = PR_FR_CTA * PR_BK * 2;   // 4 KB: activation hi (or lo) of one stage
constexpr int PR_WBLK = 128 * PR_BK * 2;          // 8 KB: one 128-channel weight block, hi (or lo), of one stage
#ifndef PR_EPI_WARPS
#define PR_EPI_WARPS 8              // 4 or 8: warps draining TMEM (one or two per lane quarter)
#endif
constexpr int PR_PRODUCERS = 256, PR_EPI = PR_EPI_WARPS * 32;
constexpr int PR_THREADS = 64 + PR_EPI + PR_PRODUCERS;
constexpr int PR_MAXK = 1024;
constexpr int PR_AFF_BYTES = 2 * PR_MAXK * 4;
constexpr int PR_CW = 16;           // frames (TMEM columns) per epilogue chunk

// SUB = 32-k sub-tiles per shared-memory stage.  A stage hand-over (producers -> full -> relay -> MMA issue -> commit -> empty ->
// producers) has a round trip of ~1.9 us whatever it carries (profiles/r02_gemm_notes.md): with one 256-channel block (NB = 1:
// M = 32 ... 256, the DPRNN projections, the U-Net shell, the decoder) a 32-k stage is only 8 KB of fp32 operand per CTA, and
// that hand-over rate - not HBM, not the tensor pipe - capped those launches at ~2.5 TB/s of operand reads.  NB = 1 has the
// shared memory for 64-k stages (SUB = 2: 48 KB x 4), which halves the hand-overs per byte.
template <int NB, int SUB = 1>
struct PairCfg {
  static constexpr int kWBytes = 2 * NB * PR_WBLK;                          // weight bytes per 32-k sub-tile per CTA
  static constexpr int kXBytes = 2 * PR_XPART;                              // operand bytes (hi + lo) per sub-tile per CTA
  static constexpr int kStageBytes = SUB * (kXBytes + kWBytes);             // NB=2: 40 KB, NB=1: 24 KB (SUB=2: 48 KB)
  static constexpr int kWOff = SUB * kXBytes;                               // stage layout: [X sub 0 .. X sub SUB-1 | W sub 0 .. ]
  static constexpr int kAccCols = NB * PR_FRAMES;                           // TMEM columns of one accumulator set
  static constexpr int kFinBytes = 3 * 256 * 8;                             // fp64 scratch of the fused statistics finalize
  static constexpr int kSmem = PR_STAGES * kStageBytes + PR_AFF_BYTES + 1024 /*align*/ + 512 /*barriers, scratch*/ + kFinBytes;
};

// byte offset of 16-byte chunk c of row r in a [rows x 32 k] bf16 tile, 64-byte swizzle (Swizzle<2,4,3>)
__host__ __device__ constexpr uint32_t pr_swz(uint32_t r, uint32_t c) { return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4); }

// K-major, SWIZZLE_64B, 8-row atoms 512 B apart
__device__ __forceinline__ uint64_t pr_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// D=f32, A=B=bf16, K-major both, M=256 (pair), N=128
constexpr uint32_t PR_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PR_FRAMES >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

struct PairTile {
  uint32_t b, rt, nh;
};
__device__ __forceinline__ PairTile pr_tile(int64_t t, int64_t n_rt, int64_t n_nh) {
  const uint32_t tt = (uint32_t)t, nrt = (uint32_t)n_rt, nnh = (uint32_t)n_nh;
  const uint32_t q = tt / nnh;
  PairTile c;
  c.nh = tt - q * nnh;
  c.b = q / nrt;
  c.rt = q - c.b * nrt;
  return c;
}

// Epilogue of one PR_CW-frame chunk of one channel (thread): out = act(acc + bias) (+ residual), strided scalar stores that
// are coalesced across the warp's 32 channels, and shifted sums about a pivot for the gLN statistics.  ACT = -1 takes
// the activation at run time (uniform branch per element); kFull = all frames of the chunk valid, no predication.
template <int ACT, bool kRes, bool kFull>
__device__ __forceinline__ void epi_chunk(const float (&v)[PR_CW], float bsum, float eslope, float* yp, int64_t ystride, const float* rp,
                                          int64_t rstride, int nj, bool first, float& piv, float& ssum, float& ssq, int act = ACT) {
  float r[PR_CW];
  if constexpr (kRes) {  // all residual loads of the chunk in flight together
    const float* p = rp;
#pragma unroll
    for (int j = 0; j < PR_CW; ++j) {
      r[j] = (kFull || j < nj) ? __ldg(p) : 0.f;
      p += rstride;
    }
  }
#pragma unroll
  for (int j = 0; j < PR_CW; ++j) {
    float o = v[j] + bsum;
    if (ACT == PS_ACT_NONE) {
    } else if (act == PS_ACT_PRELU) o = o > 0.f ? o : o * eslope;
    else if (act == PS_ACT_RELU) o = (o != o) ? o : fmaxf(o, 0.f);
    if constexpr (kRes) o += r[j];
    if (j == 0 && first) piv = o;  // statistics pivot: this thread's first output of the block
    if (kFull || j < nj) {
      *yp = o;
      const float dv = o - piv;
      ssum += dv;
      ssq = fmaf(dv, dv, ssq);
    }
    yp += ystride;
  }
}

// Epilogue chunk with a fused LayerNorm over the 128 channels of a frame (nn.LayerNorm + residual of the DPRNN blocks,
// dprnn.py:161-163,173-175):  out = residual + (x - mean_f) * rstd_f * gamma[c] + beta[c],  x = acc + bias.
// A frame's 128 channels are the 128 TMEM lanes of the leader CTA = the four warps (lane quarters) of one epilogue warp
// group.  Per 16-frame chunk: a reduce-scatter butterfly leaves lane l with the 32-lane sum / sum of squares of frame
// (l >> 1) & 15 (16 + 16 shuffles instead of 160), the four quarters meet in shared memory (two 128-thread named
// barriers), 16 threads turn them into (mean, rstd), everyone reads them back as broadcast loads.
__device__ __forceinline__ float ln_butterfly16(const float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) a8[i] = (b4 ? v[i + 8] : v[i]) + __shfl_xor_sync(0xffffffffu, b4 ? v[i] : v[i + 8], 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) a4[i] = (b3 ? a8[i + 4] : a8[i]) + __shfl_xor_sync(0xffffffffu, b3 ? a8[i] : a8[i + 4], 8);
#pragma unroll
  for (int i = 0; i < 2; ++i) a2[i] = (b2 ? a4[i + 2] : a4[i]) + __shfl_xor_sync(0xffffffffu, b2 ? a4[i] : a4[i + 2], 4);
  float a1 = (b1 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? a2[0] : a2[1], 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}

template <bool kRes>
__device__ __forceinline__ void epi_chunk_ln(float (&v)[PR_CW], float bsum, float gam, float bet, float eps, float* yp, int64_t ystride,
                                             const float* rp, int64_t rstride, int nj, int lane, int q, int g, float2* part_s, float2* stat_s) {
  static_assert(PR_CW == 16, "the butterfly is written for 16-frame chunks");
  float r[PR_CW];
  if constexpr (kRes) {
    const float* p = rp;
#pragma unroll
    for (int j = 0; j < PR_CW; ++j) {
      r[j] = j < nj ? __ldg(p) : 0.f;
      p += rstride;
    }
  }
  float sq[PR_CW];
#pragma unroll
  for (int j = 0; j < PR_CW; ++j) {
    v[j] += bsum;
    sq[j] = v[j] * v[j];
  }
  const float s1 = ln_butterfly16(v, lane), s2 = ln_butterfly16(sq, lane);
  const int f = (lane >> 1) & 15;
  if ((lane & 1) == 0) part_s[(g * 4 + q) * 16 + f] = make_float2(s1, s2);
  asm volatile("bar.sync %0, 128;" ::"r"(3 + g) : "memory");
  if (q == 0 && lane < 16) {
    float S = 0.f, SS = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float2 t = part_s[(g * 4 + w) * 16 + lane];
      S += t.x;
      SS += t.y;
    }
    const float mean = S * (1.f / 128.f);
    const float var = fmaxf(SS * (1.f / 128.f) - mean * mean, 0.f);
    stat_s[g * 16 + lane] = make_float2(mean, rsqrtf(var + eps));
  }
  asm volatile("bar.sync %0, 128;" ::"r"(3 + g) : "memory");
#pragma unroll
  for (int j = 0; j < PR_CW; ++j) {
    const float2 st = stat_s[g * 16 + j];
    float o = fmaf((v[j] - st.x) * st.y, gam, bet);
    if constexpr (kRes) o += r[j];
    if (j < nj) *yp = o;
    yp += ystride;
  }
}

// PRO: PS_PRO_NONE, PS_PRO_AFFINE (norm affine + PReLU), PS_PRO_MASK (x * act(x2): mask apply in front of the decoder) or
// PR_PRO_AFFINE_TANH (kernel-internal: the affine followed by tanh - eval BatchNorm + nn.Tanh in front of the second conv of
// AttentiveStatisticsPooling, lobe/pooling.py:71-86,104-105; its own instantiation so the TCN path's producers are untouched)
constexpr int PR_PRO_AFFINE_TANH = 4;
template <int PRO, int NB, bool kLN = false, int SUB = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PR_THREADS, 1)
    gemm_pair_kernel(const ps_gemm_t d, const int64_t n_rt, const int64_t n_nh, const int64_t n_tiles, const int dbg, const int w_from2) {
  using Cfg = PairCfg<NB, SUB>;
  constexpr int STAGE = Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* aff_s = reinterpret_cast<float*>(sm + PR_STAGES * STAGE);
  const uint32_t bars = base + PR_STAGES * STAGE + PR_AFF_BYTES;
  // barrier map (8 B each, room for 8 stages): full @0, empty @64, pfull @128, tfull[0..1] @192, tempty[0..1] @208
  static_assert(PR_STAGES <= 8, "barrier map holds 8 stages");
  const uint32_t bar_full = bars, bar_empty = bars + 64, bar_pfull = bars + 128, bar_tfull = bars + 192, bar_tempty = bars + 208;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + PR_STAGES * STAGE + PR_AFF_BYTES + 224);
  Wf* wf_s = reinterpret_cast<Wf*>(sm + PR_STAGES * STAGE + PR_AFF_BYTES + 240);  // [NB][epilogue warps]
  volatile int* fin_flag_s = reinterpret_cast<volatile int*>(sm + PR_STAGES * STAGE + PR_AFF_BYTES + 232);
  double* fin_s = reinterpret_cast<double*>(sm + PR_STAGES * STAGE + PR_AFF_BYTES + 512);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int64_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int KB = (int)(d.K / (PR_BK * SUB));  // stages per tile

  if (tid == 0) {
    for (int s = 0; s < PR_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, PR_PRODUCERS / (SUB == 2 ? 32 : 64) + 1);  // the producer warps of this stage (4, or all 8 at SUB = 2) + the weight copy
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_pfull + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 2 * PR_EPI_WARPS);  // one arrival per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  pdl_trigger();  // (PS_PDL=1) the next kernel of the stream may start its prologue on SMs this grid leaves idle / frees
  if (warp == 1) tmem_alloc_pair(smem_u32((const void*)tmem_ptr_s), 512);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();  // every earlier kernel of the stream is complete: operands, folded norms, residual are in place

  if (warp == 0) {
    // ===================== weight loader =====================
    // (whole warp walks the loop so every value stays warp-uniform; one elected lane issues)
    int s = 0;
    uint32_t ph = 0;
    const uint8_t* wp = reinterpret_cast<const uint8_t*>(d.W_packed);
    for (int64_t t = pair; t < n_tiles; t += n_pairs) {
      const PairTile tc = pr_tile(t, n_rt, n_nh);
      constexpr uint32_t WST = SUB * Cfg::kWBytes;  // the SUB sub-tiles of a stage are adjacent in the packed image
      const uint8_t* src = wp + (size_t)(tc.nh * 2 + rank) * KB * WST;
      // w_from2 (NB = 1 only): the image is the TWO-block one of a 512-channel group - per 32-k stage
      // [hi block 0 | hi block 1 | lo block 0 | lo block 1] - and this tile is block nh & 1 of group nh >> 1: the hi and lo
      // parts of every sub-tile are fetched as two 8 KB copies
      const uint8_t* src2 = wp + (size_t)((tc.nh >> 1) * 2 + rank) * (KB * SUB) * (4 * PR_WBLK) + (size_t)(tc.nh & 1) * PR_WBLK;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_full + 8 * s, WST);
          if (PR_DBG(1)) {  // experiment: no weight traffic (results are garbage)
            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(bar_full + 8 * s), "r"(WST) : "memory");
          } else if (NB == 1 && w_from2) {
#pragma unroll
            for (int sub = 0; sub < SUB; ++sub) {
              const uint8_t* g = src2 + (size_t)(kb * SUB + sub) * (4 * PR_WBLK);
              const uint32_t dst = base + s * STAGE + Cfg::kWOff + sub * Cfg::kWBytes;
              bulk_g2s(dst, g, PR_WBLK, bar_full + 8 * s);
              bulk_g2s(dst + PR_WBLK, g + 2 * PR_WBLK, PR_WBLK, bar_full + 8 * s);
            }
          } else
            bulk_g2s(base + s * STAGE + Cfg::kWOff, src + (size_t)kb * WST, WST, bar_full + 8 * s);
        }
        __syncwarp();
        if (++s == PR_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    int s = 0;
    uint32_t ph = 0;
    if (rank == 0) {
      // ===================== MMA issuer (leader) =====================
      int64_t it = 0;
      for (int64_t t = pair; t < n_tiles; t += n_pairs, ++it) {
        const int a = (int)(it & 1);
        const uint32_t aph = (uint32_t)((it >> 1) & 1);
        mbar_wait(bar_tempty + 8 * a, aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(a * Cfg::kAccCols);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);   // this CTA's stage: producers + weight copy
          mbar_wait(bar_pfull + 8 * s, ph);  // the peer's stage (relayed)
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int sub = 0; sub < SUB; ++sub) {
              const uint32_t sa = base + s * STAGE;
              const uint32_t xa = sa + sub * Cfg::kXBytes, wa = sa + Cfg::kWOff + sub * Cfg::kWBytes;
              const uint64_t x_hi = pr_desc(xa), x_lo = pr_desc(xa + PR_XPART);
#pragma unroll
              for (int mb = 0; mb < NB; ++mb) {
                const uint64_t w_hi = pr_desc(wa + mb * PR_WBLK);
                const uint64_t w_lo = pr_desc(wa + (NB + mb) * PR_WBLK);
#pragma unroll
                for (int k = 0; k < PR_BK / 16; ++k) {
                  const uint64_t ko = (uint64_t)((k * 32) >> 4);
                  const uint32_t dd = tmem_d + (uint32_t)(mb * PR_FRAMES);
                  // small cross terms first, the dominant hi*hi last
                  umma_bf16_pair(dd, w_lo + ko, x_hi + ko, PR_IDESC, (kb | sub | k) != 0);
                  umma_bf16_pair(dd, w_hi + ko, x_lo + ko, PR_IDESC, 1);
                  umma_bf16_pair(dd, w_hi + ko, x_hi + ko, PR_IDESC, 1);
                }
              }
            }
            umma_commit_pair(bar_empty + 8 * s);                  // frees this stage in both CTAs when the MMAs retire
            if (kb == KB - 1) umma_commit_pair(bar_tfull + 8 * a);  // accumulators complete -> both epilogues
          }
          __syncwarp();
          if (++s == PR_STAGES) { s = 0; ph ^= 1; }
        }
      }
    } else {
      // ===================== stage relay (peer) =====================
      // tells the leader that this CTA's half of a stage (its 64 activation frames and its weight blocks) landed
      for (int64_t t = pair; t < n_tiles; t += n_pairs) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          if (elect_one()) mbar_arrive_remote(bar_pfull + 8 * s, 0);
          __syncwarp();
          if (++s == PR_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 2 + PR_EPI_WARPS) {
    // ===================== epilogue (warps 2..2+EW) =====================
    // Work unit = (256-channel block mb, 32-frame chunk c); warp group g = (warp-2)/4 takes the units u = mb*4+c with
    // u % G == g, so with 8 epilogue warps two warps drain each TMEM lane quarter concurrently.
    constexpr int G = PR_EPI_WARPS / 4;
    constexpr int NCH = PR_FRAMES / PR_CW;  // chunks per 256-channel block
    const int q = warp & 3;            // TMEM lane quarter this warp may read (hardware rule: warp id % 4)
    const int g = (warp - 2) >> 2;
    const int cl = q * 32 + lane;      // channel within this CTA's 128-channel block = TMEM lane
    const int et = tid - 64;
    const float eslope = d.epi_slope ? __ldg(d.epi_slope) : 0.f;
    const int epi_act = d.epi_act;
    const int64_t ystride = d.y_row_stride, rstride = d.res_row_stride;
    const float* const bias = d.bias;
    const float* const bias_batch = d.bias_batch;
    const bool want_stats = d.stats_partials != nullptr;
    int64_t it = 0;
    for (int64_t t = pair; t < n_tiles; t += n_pairs, ++it) {
      const PairTile tc = pr_tile(t, n_rt, n_nh);
      const int a = (int)(it & 1);
      const uint32_t aph = (uint32_t)((it >> 1) & 1);
      const int64_t frame0 = (int64_t)tc.rt * PR_FRAMES;
      const int nvalid = (int)((d.rows - frame0) < PR_FRAMES ? (d.rows - frame0) : PR_FRAMES);
      const int64_t ch0 = (int64_t)tc.nh * (256 * NB) + rank * 128;  // + mb*256 + cl
      float* yb = d.Y + tc.b * d.y_batch_stride + frame0 * ystride + ch0 + cl;
      const float* rb = d.residual ? d.residual + tc.b * d.res_batch_stride + frame0 * rstride + ch0 + cl : nullptr;
      if (rb) {
        // this tile's MMAs are still in flight: pull the residual lines of this warp's chunks (32 channels = 128 B per
        // frame) into L2, one frame per lane
#pragma unroll
        for (int u = 0; u < NB * 4; ++u) {
          if (u % G != g) continue;
          const int mb = u >> 2, f = (u & 3) * 32 + lane;
          if (f < nvalid && ch0 + mb * 256 < d.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(rb - lane + f * rstride + mb * 256));
        }
        // (32-frame granularity: the two PR_CW-frame chunks a warp group drains back to back)
      }
      mbar_wait_relaxed(bar_tfull + 8 * a, aph);
      tc_fence_after();
      float piv[NB], ssum[NB], ssq[NB], cnt[NB];
#pragma unroll
      for (int mb = 0; mb < NB; ++mb) { piv[mb] = 0.f; ssum[mb] = 0.f; ssq[mb] = 0.f; cnt[mb] = 0.f; }
#pragma unroll
      for (int mb = 0; mb < NB; ++mb) {
        const int64_t ch = ch0 + mb * 256 + cl;
        if (ch0 + mb * 256 + q * 32 >= d.M) continue;  // zero-padded rows (M = 128: the peer's block; M = 32: lane quarters 1-3)
        float bsum = bias ? __ldg(bias + ch) : 0.f;
        if (bias_batch) bsum += __ldg(bias_batch + tc.b * d.M + ch);
        // real loops (not unrolled): the chunk body exists once per variant, which keeps the kernel inside the
        // instruction cache (the fully unrolled version stalled 4 warps per issue on instruction fetch)
#pragma unroll 1
        for (int c = g * (32 / PR_CW); c < NCH; c += (c % (32 / PR_CW) == 32 / PR_CW - 1) ? (G - 1) * (32 / PR_CW) + 1 : 1) {
          const int nj = nvalid - c * PR_CW;  // valid frames of this chunk (warp-uniform)
          if (nj <= 0) break;
          float v[PR_CW];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * Cfg::kAccCols + mb * PR_FRAMES + c * PR_CW), v);
          float* yp = yb + (int64_t)(c * PR_CW) * ystride + mb * 256;
          const float* rp = rb ? rb + (int64_t)(c * PR_CW) * rstride + mb * 256 : nullptr;
          const bool first = cnt[mb] == 0.f;
          cnt[mb] += (float)(nj < PR_CW ? nj : PR_CW);
          if (PR_DBG(4)) continue;
          if constexpr (kLN) {
            float2* part_s = reinterpret_cast<float2*>(fin_s);  // [2 groups][4 quarters][16 frames]; then [2][16] (mean, rstd)
            float2* stat_s = part_s + 2 * 4 * 16;
            const float gam = d.ln_gamma ? __ldg(d.ln_gamma + ch) : 1.f, bet = d.ln_beta ? __ldg(d.ln_beta + ch) : 0.f;
            if (rp) epi_chunk_ln<true>(v, bsum, gam, bet, d.ln_eps, yp, ystride, rp, rstride, nj, lane, q, g, part_s, stat_s);
            else epi_chunk_ln<false>(v, bsum, gam, bet, d.ln_eps, yp, ystride, rp, rstride, nj, lane, q, g, part_s, stat_s);
          } else if (nj >= PR_CW) {
            if (rp) {
              if (epi_act == PS_ACT_NONE) epi_chunk<PS_ACT_NONE, true, true>(v, bsum, eslope, yp, ystride, rp, rstride, PR_CW, first, piv[mb], ssum[mb], ssq[mb]);
              else epi_chunk<-1, true, true>(v, bsum, eslope, yp, ystride, rp, rstride, PR_CW, first, piv[mb], ssum[mb], ssq[mb], epi_act);
            } else {
              if (epi_act == PS_ACT_NONE) epi_chunk<PS_ACT_NONE, false, true>(v, bsum, eslope, yp, ystride, rp, rstride, PR_CW, first, piv[mb], ssum[mb], ssq[mb]);
              else epi_chunk<-1, false, true>(v, bsum, eslope, yp, ystride, rp, rstride, PR_CW, first, piv[mb], ssum[mb], ssq[mb], epi_act);
            }
          } else {
            if (rp) epi_chunk<-1, true, false>(v, bsum, eslope, yp, ystride, rp, rstride, nj, first, piv[mb], ssum[mb], ssq[mb], epi_act);
            else epi_chunk<-1, false, false>(v, bsum, eslope, yp, ystride, rp, rstride, nj, first, piv[mb], ssum[mb], ssq[mb], epi_act);
          }
        }
      }
      // all TMEM reads of this accumulator set are complete (tcgen05.wait::ld): hand it back to the leader's MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(bar_tempty + 8 * a, 0);
      if (want_stats) {
#pragma unroll
        for (int mb = 0; mb < NB; ++mb) {
          Wf mine;
          mine.n = cnt[mb]; mine.mean = 0.f; mine.m2 = 0.f;
          if (mine.n > 0.f) {
            const float md = ssum[mb] / mine.n;
            mine.mean = piv[mb] + md;
            mine.m2 = fmaxf(ssq[mb] - ssum[mb] * md, 0.f);
          }
          Wf w = wf_warp_reduce(mine);
          if (lane == 0) wf_s[mb * PR_EPI_WARPS + (warp - 2)] = w;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(PR_EPI) : "memory");
        const int64_t slots_m = (d.M + 127) / 128;
        const int64_t slots = n_rt * slots_m;
        if (et < NB && ch0 + et * 256 < d.M) {
          // fixed merge order over the epilogue warps -> deterministic
          const int mb = et;
          Wf tot = wf_s[mb * PR_EPI_WARPS];
#pragma unroll
          for (int w = 1; w < PR_EPI_WARPS; ++w) tot = wf_merge(tot, wf_s[mb * PR_EPI_WARPS + w]);
          float* o = d.stats_partials + (tc.b * slots + tc.rt * slots_m + tc.nh * (2 * NB) + mb * 2 + rank) * 3;
          o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
          if (d.fin_scale) __threadfence();  // visible device-wide before this CTA's count below
        }
        asm volatile("bar.sync 1, %0;" ::"n"(PR_EPI) : "memory");
        if (d.fin_scale) {
          // Fused gLN/gGN finalize: count the slots written for this item; whoever writes the last one merges all of
          // the item's partials (fixed slot order: the result does not depend on which CTA does it) into the folded
          // affine the consumer kernel reads - the separate 64-CTA finalize launch after every producer is gone.
          static_assert(PR_EPI == 256, "the fused finalize runs on the 256 epilogue threads");
          if (et == 0) {
            int nblk = 0;
#pragma unroll
            for (int mb = 0; mb < NB; ++mb) nblk += (ch0 + mb * 256 < d.M) ? 1 : 0;
            int last = 0;
            if (nblk > 0) {
              __threadfence();
              const unsigned int old = atomicAdd(d.fin_counter + tc.b, (unsigned int)nblk);
              last = (old + (unsigned int)nblk == (unsigned int)slots) ? 1 : 0;
            }
            *fin_flag_s = last;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(PR_EPI) : "memory");
          if (*fin_flag_s) {
            __threadfence();
            stats_finalize_item<1>(d.stats_partials + tc.b * slots * 3, slots, d.fin_gamma, d.fin_beta, d.fin_eps, d.M,
                                   d.fin_scale + tc.b * d.M, d.fin_shift + tc.b * d.M, nullptr, et, fin_s, fin_s + 256, fin_s + 512);
            if (et == 0) d.fin_counter[tc.b] = 0;  // ready for the next launch
          }
          asm volatile("bar.sync 1, %0;" ::"n"(PR_EPI) : "memory");
        }
      }
    }
  } else {
    // ===================== activation producers (last 8 warps) =====================
    // One iteration covers 64 k of this CTA's 64 frames = two stages: threads 0..127 fill stage s (k 0..31), threads
    // 128..255 stage s+1 (k 32..63).  Within a half, 4 threads cover the 32 k of a frame (coalesced 128 B) and a
    // warp covers 8 consecutive frames, so each 8-lane phase of a 128-bit shared store lands on two consecutive swizzle
    // rows = 32 distinct banks.  Every thread handles its chunk for frames r and r+32, so the affine is read once
    // per 16 elements.
    const int pt = tid - (64 + PR_EPI);
    const int half = pt >> 7;
    const int u = pt & 127;
    const uint32_t chunk = (uint32_t)(u & 3);
    const int r0 = u >> 2;  // 0..31
    // K == 32 (the learned encoder's window, lobe/encoder.py:50-56): a tile is ONE stage, so the two halves take alternate
    // tiles of this pair instead of alternate stages of one tile - the stage walk (s, s+2, ..) is the same.
    const bool k32 = d.K == 32;
    const int kofs = (k32 ? 0 : half * 32) + (int)chunk * 8;
    const int KB64 = k32 ? 1 : (int)(d.K / 64);
    const int K = (int)d.K;
    const int64_t t_step = k32 ? 2 * n_pairs : n_pairs;
    const float pslope = (d.pro_act == PS_ACT_PRELU && d.pro_slope) ? __ldg(d.pro_slope) : 1.f;  // AFFINE w/o act = slope 1
    // this thread's stage: half 0 walks stages 0,2,4,.., half 1 walks 1,3,5,.. (mod the ring depth); at SUB = 2 a stage is
    // 64 k, so both halves fill the SAME stage (sub-tile = half) and walk 0,1,2,..
    int s = SUB == 2 ? 0 : half;
    uint32_t ph = 0;

    struct XBuf {
      float4 v[2][2];
      float4 m[PRO == PS_PRO_MASK ? 2 : 1][2];  // mask operand (MASK prologue only)
    };
    struct Cur {
      int64_t t;
      int kb;
      const float* x0;   // &X[b][0][kofs]
      int64_t row_base;  // rt*128 + rank*64 + r0
      int64_t b;
    };
    const int64_t last_row = d.rows - 1;
    auto decode = [&](Cur& c) {
      if (c.t < n_tiles) {
        const PairTile tc = pr_tile(c.t, n_rt, n_nh);
        c.b = tc.b;
        c.row_base = (int64_t)(PR_DBG(64) ? 0 : tc.rt) * PR_FRAMES + rank * PR_FR_CTA + r0;
        c.x0 = d.X + (PR_DBG(64) ? 0 : tc.b) * d.x_batch_stride + kofs;  // (64: experiment, every tile reads tile 0)
      }
    };
    auto advance = [&](Cur& c) {
      if (++c.kb == KB64) {
        c.kb = 0;
        c.t += t_step;
        decode(c);
      }
    };
    // frames past the end of the item re-read its last frame: their accumulator columns are never stored
    const int64_t x2_delta = (PRO == PS_PRO_MASK) ? (d.X2 - d.X) : 0;  // the mask has the strides of X
    const int mask_act = d.pro_act;
    auto issue = [&](XBuf& x, const Cur& c) {
      if (PR_DBG(2)) return;  // experiment: no activation loads
      const float* xb = c.x0 + c.kb * 64;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        int64_t row = c.row_base + p * 32;
        row = row < last_row ? row : last_row;
        const float4* src = reinterpret_cast<const float4*>(xb + row * d.x_row_stride);
        x.v[p][0] = __ldg(src);
        x.v[p][1] = __ldg(src + 1);
        if constexpr (PRO == PS_PRO_MASK) {
          const float4* msrc = reinterpret_cast<const float4*>(xb + row * d.x_row_stride + x2_delta);
          x.m[p][0] = __ldg(msrc);
          x.m[p][1] = __ldg(msrc + 1);
        }
      }
    };
    int64_t staged_b = -1;
    auto stage_affine = [&](int64_t b) {
      if constexpr (PRO == PS_PRO_AFFINE || PRO == PR_PRO_AFFINE_TANH) {
        if (b != staged_b) {
          asm volatile("bar.sync 2, %0;" ::"n"(PR_PRODUCERS) : "memory");  // every producer is done with the old rows
          const float* pa = d.pro_a + b * d.pro_batch_stride;
          const float* pb = d.pro_b + b * d.pro_batch_stride;
          for (int k = pt * 4; k < K; k += PR_PRODUCERS * 4) {
            *reinterpret_cast<float4*>(aff_s + k) = __ldg(reinterpret_cast<const float4*>(pa + k));
            *reinterpret_cast<float4*>(aff_s + K + k) = __ldg(reinterpret_cast<const float4*>(pb + k));
          }
          asm volatile("bar.sync 2, %0;" ::"n"(PR_PRODUCERS) : "memory");
          staged_b = b;
        }
      }
    };
    auto process = [&](const XBuf& x, int kb) {
      float sc[8], sh[8];
      if constexpr (PRO == PS_PRO_AFFINE || PRO == PR_PRO_AFFINE_TANH) {
        const int k0 = kb * 64 + kofs;
        const float4 a0 = *reinterpret_cast<const float4*>(aff_s + k0), a1 = *reinterpret_cast<const float4*>(aff_s + k0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(aff_s + K + k0), b1 = *reinterpret_cast<const float4*>(aff_s + K + k0 + 4);
        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
      }
      const int my_s = s;
      mbar_wait(bar_empty + 8 * my_s, ph ^ 1);
      uint8_t* x_hi = sm + my_s * STAGE + (SUB == 2 ? half * Cfg::kXBytes : 0);
      uint8_t* x_lo = x_hi + PR_XPART;
#pragma unroll
      for (int p = 0; p < 2 && !PR_DBG(8); ++p) {
        const int r = p * 32 + r0;
        float v[8] = {x.v[p][0].x, x.v[p][0].y, x.v[p][0].z, x.v[p][0].w, x.v[p][1].x, x.v[p][1].y, x.v[p][1].z, x.v[p][1].w};
        if constexpr (PRO == PS_PRO_MASK) {
          const float mk[8] = {x.m[p][0].x, x.m[p][0].y, x.m[p][0].z, x.m[p][0].w, x.m[p][1].x, x.m[p][1].y, x.m[p][1].z, x.m[p][1].w};
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= apply_act(mk[i], mask_act, 0.f);
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float u0 = v[i], u1 = v[i + 1];
          if constexpr (PRO == PS_PRO_AFFINE) {
            u0 = fmaf(u0, sc[i], sh[i]);
            u1 = fmaf(u1, sc[i + 1], sh[i + 1]);
            u0 = u0 > 0.f ? u0 : u0 * pslope;
            u1 = u1 > 0.f ? u1 : u1 * pslope;
          } else if constexpr (PRO == PR_PRO_AFFINE_TANH) {
            u0 = tanhf(fmaf(u0, sc[i], sh[i]));
            u1 = tanhf(fmaf(u1, sc[i + 1], sh[i + 1]));
          }
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(u0, u1);
          const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
          const float r0f = u0 - __uint_as_float(hb << 16);
          const float r1f = u1 - __uint_as_float(hb & 0xFFFF0000u);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0f, r1f);
          hi[i >> 1] = hb;
          lo[i >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        const uint32_t off = pr_swz((uint32_t)r, chunk);
        *reinterpret_cast<uint4*>(x_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(x_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor cores of both SMs (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * my_s);  // one arrival per warp (a warp lies entirely in one half)
      s += SUB == 2 ? 1 : 2;
      if (s >= PR_STAGES) { s -= PR_STAGES; ph ^= 1; }
    };

    // Register double buffer: the loads of block n+1 are in flight while block n is transformed.  (Measured: a third
    // buffer and an L2 software prefetch 4 blocks ahead were both SLOWER - the stream is not latency-bound.)
    XBuf x0, x1;
    Cur pr, ld;
    pr.t = pair + (k32 ? half * n_pairs : 0); pr.kb = 0;
    decode(pr);
    ld = pr;
    auto ld_next = [&](XBuf& x) {
      if (ld.t < n_tiles) {
        issue(x, ld);
        advance(ld);
      }
    };
    auto pr_next = [&](const XBuf& x) -> bool {
      if (pr.t >= n_tiles) return false;
      stage_affine(pr.b);
      process(x, pr.kb);
      advance(pr);
      return true;
    };
    ld_next(x0);
    while (true) {
      ld_next(x1);
      if (!pr_next(x0)) break;
      ld_next(x0);
      if (!pr_next(x1)) break;
    }
  }

  // nobody leaves while the pair's MMAs may still read this CTA's shared memory or write its TMEM
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- weight packing
// W [M, K] fp32 -> per (512-channel group nh, pair rank, 32-k stage): [hi block 0 .. hi block NB-1 | lo block 0 ..],
// each block = 128 channels x 32 k bf16 in the K-major 64-byte-swizzled shared-memory image.  Channel of (nh, mb,
// rank, r) = nh*256*NB + mb*256 + rank*128 + r.
__global__ void pack_weights_pair_kernel(const float* __restrict__ W, int64_t ldw, int64_t M, int64_t K, int NB, uint8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t Mp = (M + 255) / 256 * 256;  // channels are padded with zero rows to whole 256-channel blocks
  if (i >= Mp * K) return;
  const int64_t n = i / K, k = i % K;
  const int64_t grp = 256 * NB;
  const int64_t nh = n / grp, w = n % grp, mb = w / 256, rank = (w % 256) / 128, r = w % 128;
  const int64_t kb = k / PR_BK, kk = k % PR_BK, KB = K / PR_BK;
  const float x = n < M ? W[n * ldw + k] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  const size_t stage = ((size_t)(nh * 2 + rank) * KB + kb) * (size_t)(2 * NB * PR_WBLK);
  const size_t off = (size_t)pr_swz((uint32_t)r, (uint32_t)(kk >> 3)) + (size_t)(kk & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(out + stage + mb * PR_WBLK + off) = h;
  *reinterpret_cast<__nv_bfloat16*>(out + stage + (NB + mb) * PR_WBLK + off) = l;
}

static inline int pair_nb(int64_t M) { return (((M + 255) / 256 * 256) % 512 == 0) ? 2 : 1; }

int gemm_pair_pack(const float* W, int64_t ldw, int64_t M, int64_t K, void* packed, cudaStream_t s) {
  const unsigned blocks = (unsigned)cdiv((M + 255) / 256 * 256 * K, 256);
  pack_weights_pair_kernel<<<blocks, 256, 0, s>>>(W, ldw, M, K, pair_nb(M), reinterpret_cast<uint8_t*>(packed));
  PS_CHECK_LAUNCH("pack_weights_pair_kernel");
  return PS_OK;
}

template <int PRO, int NB, bool kLN = false, int SUB = 1>
static int launch_pair(const ps_gemm_t& d, cudaStream_t s, int dev, int64_t grid, int64_t n_rt, int64_t n_nh, int64_t n_tiles, int w_from2 = 0) {
  static SmemOnce<1> once;  // per instantiation and device
  if (int rc = once.ensure(dev, 0, gemm_pair_kernel<PRO, NB, kLN, SUB>, PairCfg<NB, SUB>::kSmem, "cudaFuncSetAttribute(gemm_pair_kernel)")) return rc;
#ifdef PS_EXPERIMENTS
  static EnvInt dbg_e;
  const int dbg = dbg_e.get("PS_PAIR_DBG", 0);
#else
  const int dbg = 0;
#endif
  cudaError_t le = launch_pdl(gemm_pair_kernel<PRO, NB, kLN, SUB>, dim3((unsigned)grid), dim3(PR_THREADS), PairCfg<NB, SUB>::kSmem, s, d, n_rt, n_nh, n_tiles, dbg, w_from2);
  if (le != cudaSuccess) { set_cuda_error(le, "gemm_pair_kernel"); return PS_ERR_CUDA; }
  return PS_OK;
}

bool gemm_pair_ln_eligible(const ps_gemm_t& d) {
  return d.M == 128 && d.pro_mode == PS_PRO_NONE && d.epi_act == PS_ACT_NONE && !d.stats_partials && !d.bias_batch;
}

bool gemm_wide_eligible(const ps_gemm_t& d, int sms);
int gemm_wide_launch(const ps_gemm_t& d, cudaStream_t s, int dev, int sms);
bool gemm_wide_tma_eligible(const ps_gemm_t& d, int sms);
int gemm_wide_tma_launch(const ps_gemm_t& d, cudaStream_t s, int dev, int sms);

// see gemm_pair_launch: a 512-channel layer with so few tiles that one-block tiles with 64-k stages win
bool gemm_pair_few_tiles(const ps_gemm_t& d, int sms) {
  static EnvInt from2_env;
  // measured (run 29, isolated launches, 74 CTA pairs): 16 ... 64 tiles of 256 frames (batch 1 ... 4 x 4 s) 0.016-0.043 ms
  // against 0.021-0.040; at 128 tiles (1.7 per pair: batch 8, or the 64 x 497-frame TSE batch) the 256-frame kernel wins
  // without a residual (0.054 against 0.057-0.064 ms) and loses with one (0.074 / 0.060 against 0.070 / 0.051 ms); from 3
  // tiles per pair on it wins everywhere.  In the step: cfg1 2.03 -> 1.52 ms, cfg4 5.27 -> 5.00 ms.
  const int thr = from2_env.get("PS_PAIR_FROM2", 10) * (d.residual ? 2 : 1);
  if (thr <= 0 || pair_nb(d.M) != 2 || d.K % 64 != 0 || d.ln_eps > 0.f) return false;
  const int64_t wide_tiles = d.batch * cdiv(d.rows, 256) * cdiv(d.M, 512);
  return wide_tiles * 10 < (int64_t)thr * (sms / 2);
}

int gemm_pair_launch(const ps_gemm_t& d, cudaStream_t s) {
  int dev = 0, sms = 0;
  if (int rc = current_device(&dev)) return rc;
  if (int rc = sm_count_of(dev, &sms)) return rc;
  if (gemm_wide_tma_eligible(d, sms)) return gemm_wide_tma_launch(d, s, dev, sms);  // PS_TC_WIDE=2: TMA-fed variant (A/B)
  int pro = d.pro_mode;  // NONE (0), AFFINE (1) or MASK (3): checked by gemm_tc_eligible
  if (pro == PS_PRO_AFFINE && d.pro_act == PS_ACT_TANH) pro = PR_PRO_AFFINE_TANH;
  // FEW tiles of a 512-channel layer (batch 1, the 497-frame items of the TSE model, a batch shard of 8 utterances): what
  // a launch costs there is the chain of stage hand-overs of ONE tile (profiles/r02_gemm_notes.md section 5), so the tile is
  // cut to one 256-channel block with 64-k stages - half the hand-overs per tile and twice the CTA pairs at work - reading
  // the hi / lo parts out of the two-block image.  PS_PAIR_FROM2 = number of 256-frame tiles per CTA pair, in tenths,
  // below which this path is taken (default 10, doubled for launches with a residual; 0 = never).
  if (gemm_pair_few_tiles(d, sms)) {
    {
      const int64_t n_rt = cdiv(d.rows, PR_FRAMES), n_nh = cdiv(d.M, 256);
      const int64_t n_tiles = d.batch * n_rt * n_nh;
      if (n_tiles >= (1LL << 31)) return PS_ERR_UNSUPPORTED;
      const int64_t max_pairs = sms / 2;
      const int64_t grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
      if (pro == PS_PRO_AFFINE) return launch_pair<PS_PRO_AFFINE, 1, false, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles, 1);
      if (pro == PR_PRO_AFFINE_TANH) return launch_pair<PR_PRO_AFFINE_TANH, 1, false, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles, 1);
      if (pro == PS_PRO_MASK) return launch_pair<PS_PRO_MASK, 1, false, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles, 1);
      return launch_pair<PS_PRO_NONE, 1, false, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles, 1);
    }
  }
  if (gemm_wide_eligible(d, sms)) return gemm_wide_launch(d, s, dev, sms);  // 256-frame tiles (ps_gemm_wide.cu)
  const int nb = pair_nb(d.M);
  const bool ln = d.ln_eps > 0.f;  // fused LayerNorm epilogue: M == 128, no prologue (checked by gemm_pair_ln_eligible)
  const int64_t n_rt = cdiv(d.rows, PR_FRAMES), n_nh = cdiv(d.M, 256 * nb);
  const int64_t n_tiles = d.batch * n_rt * n_nh;
  if (n_tiles >= (1LL << 31)) return PS_ERR_UNSUPPORTED;
  const int64_t max_pairs = sms / 2;
  const int64_t grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
  // one 256-channel block and K a multiple of 64: 64-k stages (SUB = 2), half the stage hand-overs per operand byte.
  // PS_PAIR_SUB=1 keeps 32-k stages (A/B).
  static EnvInt sub_env;
  const bool sub2 = nb == 1 && d.K % 64 == 0 && sub_env.get("PS_PAIR_SUB", 2) == 2;
  // (the LayerNorm-epilogue variant keeps 32-k stages: measured slower with 64-k ones, cfg3 projection 0.495 -> 0.538 ms)
  if (ln) return launch_pair<PS_PRO_NONE, 1, true>(d, s, dev, grid, n_rt, n_nh, n_tiles);
  if (sub2) {
    if (pro == PS_PRO_AFFINE) return launch_pair<PS_PRO_AFFINE, 1, false, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles);
    if (pro == PR_PRO_AFFINE_TANH) return launch_pair<PR_PRO_AFFINE_TANH, 1, false, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles);
    if (pro == PS_PRO_MASK) return launch_pair<PS_PRO_MASK, 1, false, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles);
    return launch_pair<PS_PRO_NONE, 1, false, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles);
  }
  if (nb == 2) {
    if (pro == PS_PRO_AFFINE) return launch_pair<PS_PRO_AFFINE, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles);
    if (pro == PR_PRO_AFFINE_TANH) return launch_pair<PR_PRO_AFFINE_TANH, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles);
    if (pro == PS_PRO_MASK) return launch_pair<PS_PRO_MASK, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles);
    return launch_pair<PS_PRO_NONE, 2>(d, s, dev, grid, n_rt, n_nh, n_tiles);
  }
  if (pro == PS_PRO_AFFINE) return launch_pair<PS_PRO_AFFINE, 1>(d, s, dev, grid, n_rt, n_nh, n_tiles);
  if (pro == PR_PRO_AFFINE_TANH) return launch_pair<PR_PRO_AFFINE_TANH, 1>(d, s, dev, grid, n_rt, n_nh, n_tiles);
  if (pro == PS_PRO_MASK) return launch_pair<PS_PRO_MASK, 1>(d, s, dev, grid, n_rt, n_nh, n_tiles);
  return launch_pair<PS_PRO_NONE, 1>(d, s, dev, grid, n_rt, n_nh, n_tiles);
}

}  // namespace ps
