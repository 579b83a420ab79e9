// tcgen05 / TMEM GEMM for the 1x1 convolutions of the TCN stack (sm_100a).
//
//   Y[b,r,n] = epi( sum_k pro(X[b,r,k]) * W[n,k] ),  fp32 in HBM, fp32-grade result.
//
// Precision: the reference tolerance (1e-3 on the waveform) rules out single-pass
// BF16 (4e-3..4e-2) and TF32 is marginal.  Both operands are
// therefore split x = hi + lo with hi = bf16(x), lo = bf16(x - hi) and three
// tensor-core passes accumulate hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator
// ("3xBF16": ~2^-17 relative per product, 30-60x better than TF32).
//
// Structure (one persistent CTA per SM, 448 threads, warp-specialised):
//   warp 0      weight loader: the weights are pre-split and pre-swizzled into the exact
//               shared-memory tile image by ps_gemm_pack_weights, so a tile is two 32 KB
//               cp.async.bulk copies (TMA bulk engine, mbarrier complete_tx) — no tensor map.
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=256, K=16, kind::f16) from
//               shared-memory descriptors (K-major, 128B swizzle) into a double-buffered
//               128x256 fp32 accumulator in TMEM (2 x 256 columns = all 512).
//   warps 2-5   epilogue: tcgen05.ld 32 columns at a time, + bias / per-item bias, activation,
//               + residual, store, and one Welford partial (count, mean, M2) per tile for the
//               gLN/gGN that follows.
//   warps 6-13  activation producers: read the fp32 activation tile from HBM (coalesced float4),
//               apply the preceding norm's folded affine + PReLU (the *prologue transform*),
//               split to bf16 hi/lo and write both into the swizzled UMMA layout in shared
//               memory.  The normalised tensor never exists in HBM.
// Pipelines: smem stages full/empty (producers+TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
#include <stdlib.h>

#include "ps_tc_ptx.cuh"

namespace ps {

constexpr int TC_BM = 128, TC_BN = 256;
constexpr int TC_PRODUCERS = 256, TC_EPI = 128;
constexpr int TC_THREADS = 64 + TC_EPI + TC_PRODUCERS;
constexpr int TC_EPI_PITCH = 36;                                  // floats per staged row: 16 B aligned, conflict-free for 128-bit access
constexpr int TC_EPI_STAGE = 4 * 32 * TC_EPI_PITCH * 4;            // one 32x32 fp32 transpose buffer per epilogue warp
constexpr int TC_MAXK = 1024;                                      // affine prologue: K <= 1024 (scale|shift staged in smem)
constexpr int TC_AFF_BYTES = 2 * TC_MAXK * 4;
constexpr int TC_PF_DIST = 4;                                      // L2 prefetch distance of the activation stream, in 64-k blocks

// Pipeline geometry.  BK = k extent of one shared-memory stage: 64 (128-byte swizzle rows, 2 stages of 96 KB) or
// 32 (64-byte swizzle rows, 4 stages of 48 KB).  Same bytes in flight, but the deeper ring lets the weight copies
// and the activation producers run three MMA-stages ahead instead of one.
template <int BK>
struct TcCfg {
  static constexpr int kBK = BK;
  static constexpr int kRowBytes = BK * 2;                   // bf16 row of one operand tile = swizzle span
  static constexpr int kStages = BK == 64 ? 2 : 4;
  static constexpr int kAPart = TC_BM * kRowBytes;
  static constexpr int kBPart = TC_BN * kRowBytes;
  static constexpr int kStageBytes = 2 * kAPart + 2 * kBPart;
  static constexpr int kSub = 64 / BK;                       // stages filled by one producer iteration (64 k)
  static constexpr int kChunks = kRowBytes / 16;             // 16-byte chunks per row: 8 or 4
  static constexpr uint32_t kSBO = 8 * kRowBytes;            // byte distance between 8-row swizzle atoms
  static constexpr uint64_t kLayout = BK == 64 ? 2 : 4;      // UMMA LayoutType: SWIZZLE_128B / SWIZZLE_64B
  static constexpr int kSmem = kStages * kStageBytes + TC_EPI_STAGE + TC_AFF_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  // byte offset of 16-byte chunk c of row r inside a tile (Swizzle<3,4,3> resp. Swizzle<2,4,3>)
  __host__ __device__ static constexpr uint32_t swz(uint32_t r, uint32_t c) {
    return r * kRowBytes + ((BK == 64 ? (c ^ (r & 7u)) : (c ^ ((r >> 1) & 3u))) << 4);
  }
};

// shared-memory matrix descriptor: K-major, swizzled rows, 8-row atoms kSBO bytes apart (cute::UMMA::SmemDescriptor)
template <class Cfg>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);      // start address      [0,14)
  d |= (uint64_t)1 << 16;                      // leading byte off.  [16,30) (unused for swizzled K-major)
  d |= (uint64_t)(Cfg::kSBO >> 4) << 32;       // stride byte offset [32,46): next 8-row atom
  d |= (uint64_t)1 << 46;                      // descriptor version [46,48) = 1 on sm_100
  d |= (uint64_t)Cfg::kLayout << 61;           // layout type        [61,64)
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, M=128, N=256
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

struct TileCoord {
  int64_t b, rt, nh;
};
// 32-bit decode (tile counts are < 2^31, checked at launch): 64-bit div/mod costs ~100 instructions each and this
// used to run four times per k-block in the producer warps
__device__ __forceinline__ TileCoord tile_coord(int64_t t, int64_t n_rt, int64_t n_nh) {
  const uint32_t tt = (uint32_t)t, nrt = (uint32_t)n_rt, nnh = (uint32_t)n_nh;
  const uint32_t q = tt / nnh;
  TileCoord c;
  c.nh = tt - q * nnh;
  const uint32_t bb = q / nrt;
  c.rt = q - bb * nrt;
  c.b = bb;
  return c;
}

// epilogue of one 32-column chunk for one thread's 4-column slice of 8 rows; ACT is compile-time so that no
// per-element switch (and none of the tanh/sigmoid code) lands in the hot loop
template <int ACT>
__device__ __forceinline__ float epi_act(float x, float slope) {
  if constexpr (ACT == PS_ACT_RELU) return (x != x) ? x : fmaxf(x, 0.f);
  if constexpr (ACT == PS_ACT_PRELU) return x > 0.f ? x : x * slope;
  return x;
}

template <bool kAffine, int BK>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const ps_gemm_t d, const int64_t n_rt, const int64_t n_nh,
                                                                const int64_t n_tiles) {
  using Cfg = TcCfg<BK>;
  constexpr int TC_STAGES = Cfg::kStages, TC_STAGE_BYTES = Cfg::kStageBytes, TC_A_PART = Cfg::kAPart, TC_B_PART = Cfg::kBPart;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;       // 128B swizzle atoms need 1024 B alignment
  uint8_t* sm = smem_raw + (base - raw);
  float* epi_stage = reinterpret_cast<float*>(sm + TC_STAGES * TC_STAGE_BYTES);
  float* aff_s = reinterpret_cast<float*>(sm + TC_STAGES * TC_STAGE_BYTES + TC_EPI_STAGE);  // [scale K | shift K] of the current item
  const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES + TC_EPI_STAGE + TC_AFF_BYTES;
  // barrier map (8 B each): full[0..3] empty[0..3] tfull[0..1] tempty[0..1], then tmem ptr, then stats scratch
  const uint32_t bar_full = bars, bar_empty = bars + 32, bar_tfull = bars + 64, bar_tempty = bars + 80;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + TC_STAGES * TC_STAGE_BYTES + TC_EPI_STAGE + TC_AFF_BYTES + 128);
  Wf* wf_s = reinterpret_cast<Wf*>(sm + TC_STAGES * TC_STAGE_BYTES + TC_EPI_STAGE + TC_AFF_BYTES + 144);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = (int)(d.K / BK);  // shared-memory stages per tile

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, TC_PRODUCERS / Cfg::kSub + 1);  // the producer threads that own this stage + the weight copy
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, TC_EPI);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_ptr_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== weight loader (bulk async copies) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint8_t* wp = reinterpret_cast<const uint8_t*>(d.W_packed);
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(t, n_rt, n_nh);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * s, 2 * TC_B_PART);
          const uint8_t* src = wp + ((size_t)(tc.nh * KB + kb) * 2) * TC_B_PART;
          const uint32_t dst = base + s * TC_STAGE_BYTES + 2 * TC_A_PART;
          bulk_g2s(dst, src, TC_B_PART, bar_full + 8 * s);
          bulk_g2s(dst + TC_B_PART, src + TC_B_PART, TC_B_PART, bar_full + 8 * s);
          if (++s == TC_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int64_t it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int a = (int)(it & 1);
        const uint32_t aph = (uint32_t)((it >> 1) & 1);
        mbar_wait(bar_tempty + 8 * a, aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(a * TC_BN);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = base + s * TC_STAGE_BYTES;
          const uint64_t a_hi = make_smem_desc<Cfg>(sa), a_lo = make_smem_desc<Cfg>(sa + TC_A_PART);
          const uint64_t b_hi = make_smem_desc<Cfg>(sa + 2 * TC_A_PART), b_lo = make_smem_desc<Cfg>(sa + 2 * TC_A_PART + TC_B_PART);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);  // +32 B per K=16 step inside the swizzle row
            // small cross terms first, the dominant hi*hi last
            umma_bf16(tmem_d, a_lo + ko, b_hi + ko, TC_IDESC, (kb | k) != 0);
            umma_bf16(tmem_d, a_hi + ko, b_lo + ko, TC_IDESC, 1);
            umma_bf16(tmem_d, a_hi + ko, b_hi + ko, TC_IDESC, 1);
          }
          umma_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs retire
          if (++s == TC_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(bar_tfull + 8 * a);  // accumulator complete -> epilogue
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM -> registers (one accumulator row per thread) -> per-warp shared-memory transpose -> coalesced
    // 128-bit global accesses: a warp instruction touches 4 rows x 128 B instead of 32 rows x 16 B.
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int et = tid - 64;
    const float eslope = d.epi_slope ? __ldg(d.epi_slope) : 0.f;
    float* stg = epi_stage + (warp - 2) * 32 * TC_EPI_PITCH;
    const int c4 = (lane & 7) * 4;  // this thread's 4 columns inside a 32-column chunk
    const int rsub = lane >> 3;     // and its row within each group of 4 rows
    int64_t it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const TileCoord tc = tile_coord(t, n_rt, n_nh);
      const int a = (int)(it & 1);
      const uint32_t aph = (uint32_t)((it >> 1) & 1);
      const int64_t row0 = tc.rt * TC_BM + q * 32;
      const int64_t n0 = tc.nh * TC_BN;
      float* yb = d.Y + tc.b * d.y_batch_stride + n0 + c4;
      const float* rb = d.residual ? d.residual + tc.b * d.res_batch_stride + n0 + c4 : nullptr;
      if (rb && row0 + lane < d.rows) {
        // this tile's MMAs are still in flight: pull the residual rows (1 KB each) into L2 now, so the epilogue's
        // loads below are L2 hits instead of 64 serial HBM round trips per tile
        const char* pr = reinterpret_cast<const char*>(d.residual + tc.b * d.res_batch_stride + (row0 + lane) * d.res_row_stride + n0);
#pragma unroll
        for (int l = 0; l < TC_BN * 4 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + l * 128));
      }
      mbar_wait_relaxed(bar_tfull + 8 * a, aph);
      tc_fence_after();
      const float* bp = d.bias ? d.bias + n0 + c4 : nullptr;
      const float* bbp = d.bias_batch ? d.bias_batch + tc.b * d.M + n0 + c4 : nullptr;
      // statistics: shifted sums about a per-thread pivot (its first output of the tile); the count is known
      // analytically, so the hot loop pays 3 flops per element instead of a full Welford update
      const int nvalid = (int)((d.rows - row0) < 32 ? (d.rows - row0) : 32);  // valid rows of this warp's quarter (may be <= 0)
      float piv = 0.f, ssum = 0.f, ssq = 0.f;
      bool have_piv = false;
#pragma unroll 1
      for (int c = 0; c < TC_BN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TC_BN + c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stg + lane * TC_EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bp) bs = __ldg(reinterpret_cast<const float4*>(bp + c * 32));
        if (bbp) {
          const float4 b2 = __ldg(reinterpret_cast<const float4*>(bbp + c * 32));
          bs.x += b2.x; bs.y += b2.y; bs.z += b2.z; bs.w += b2.w;
        }
        float4 res4[8];
        if (rb) {  // all eight residual loads of this chunk in flight together
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + rsub;
            res4[i] = rr < nvalid ? __ldg(reinterpret_cast<const float4*>(rb + (row0 + rr) * d.res_row_stride + c * 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + rsub;
          const int64_t row = row0 + rr;
          if (rr < nvalid) {
            const float4 x4 = *reinterpret_cast<const float4*>(stg + rr * TC_EPI_PITCH + c4);
            float o[4] = {x4.x + bs.x, x4.y + bs.y, x4.z + bs.z, x4.w + bs.w};
            if (d.epi_act == PS_ACT_RELU) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = epi_act<PS_ACT_RELU>(o[e], eslope);
            } else if (d.epi_act == PS_ACT_PRELU) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = epi_act<PS_ACT_PRELU>(o[e], eslope);
            }
            if (rb) { o[0] += res4[i].x; o[1] += res4[i].y; o[2] += res4[i].z; o[3] += res4[i].w; }
            *reinterpret_cast<float4*>(yb + row * d.y_row_stride + c * 32) = make_float4(o[0], o[1], o[2], o[3]);
            if (!have_piv) { piv = o[0]; have_piv = true; }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float dv = o[e] - piv;
              ssum += dv;
              ssq = fmaf(dv, dv, ssq);
            }
          }
        }
        __syncwarp();
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * a);
      if (d.stats_partials) {
        // rows rsub, rsub+4, ... below nvalid, 4 columns x 8 chunks each
        const int nrows_t = nvalid > rsub ? (nvalid - rsub + 3) / 4 : 0;
        Wf mine;
        mine.n = (float)(nrows_t * 4 * (TC_BN / 32));
        mine.mean = 0.f; mine.m2 = 0.f;
        if (mine.n > 0.f) {
          const float md = ssum / mine.n;
          mine.mean = piv + md;
          mine.m2 = fmaxf(ssq - ssum * md, 0.f);
        }
        Wf w = wf_warp_reduce(mine);
        if (lane == 0) wf_s[q] = w;
        asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI) : "memory");
        if (et == 0) {
          // fixed merge order over the four row quarters -> deterministic
          Wf tot = wf_merge(wf_merge(wf_s[0], wf_s[1]), wf_merge(wf_s[2], wf_s[3]));
          const int64_t slots_m = (d.M + 127) / 128;
          const int64_t slots = n_rt * slots_m;
          float* o = d.stats_partials + (tc.b * slots + tc.rt * slots_m + tc.nh * 2) * 3;
          o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
          o[3] = 0.f; o[4] = 0.f; o[5] = 0.f;  // second 128-column slot of this 256-wide tile stays empty
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI) : "memory");
      }
    }
  } else {
    // ===================== activation producers (warps 6..13) =====================
    // A producer iteration covers 64 k = one stage (BK=64) or two consecutive stages (BK=32).  Thread mapping: 8
    // threads cover the 64 k of a row (coalesced 256 B), 32 rows per pass, 4 passes.
    const int pt = tid - 192;
    const int kc = pt & 7;                          // k = kc*8 .. kc*8+7 of the 64-k block
    const int r0 = pt >> 3;                         // 0..31
    const int sub = kc / Cfg::kChunks;              // which of the iteration's stages this thread fills
    const uint32_t cch = (uint32_t)(kc % Cfg::kChunks);  // 16-byte chunk inside that stage's row
    const int KB64 = (int)(d.K / 64);
    const int K = (int)d.K;
    // AFFINE with no activation is PReLU with slope 1
    const float pslope = (d.pro_act == PS_ACT_PRELU && d.pro_slope) ? __ldg(d.pro_slope) : 1.f;
    int s = 0;
    uint32_t ph = 0;

    // A cursor walks this CTA's (tile, 64-k block) sequence; the tile is decoded once per tile, not per block.
    struct Cur {
      int64_t t;
      int kb;
      const float* x0;   // &X[b][rt*128 + r0][kc*8]  (row clamped per pass below)
      int64_t row_base;  // rt*128 + r0
      int64_t b;
    };
    const int64_t last_row = d.rows - 1;
    auto decode = [&](Cur& c) {
      if (c.t < n_tiles) {
        const TileCoord tc = tile_coord(c.t, n_rt, n_nh);
        c.b = tc.b;
        c.row_base = tc.rt * TC_BM + r0;
        c.x0 = d.X + tc.b * d.x_batch_stride + kc * 8;
      }
    };
    auto advance = [&](Cur& c) {
      if (++c.kb == KB64) {
        c.kb = 0;
        c.t += gridDim.x;
        decode(c);
      }
    };
    // global loads of one (tile, k-block): 8 x LDG.128 per thread, issued one k-block AHEAD of their use so the
    // HBM latency hides behind the transform of the previous block and the wait for a free stage.  Rows past the end
    // of the item re-read its last row: their accumulator rows are computed but never stored (and never enter the
    // statistics), which is cheaper than predicating the whole transform.
    auto issue = [&](float4(&x)[4][2], const Cur& c) {
      const float* xb = c.x0 + c.kb * 64;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        int64_t row = c.row_base + p * 32;
        row = row < last_row ? row : last_row;
        const float4* src = reinterpret_cast<const float4*>(xb + row * d.x_row_stride);
        x[p][0] = __ldg(src);
        x[p][1] = __ldg(src + 1);
      }
    };
    // software prefetch into L2, TC_PF_DIST blocks ahead of the register loads: one 128-byte line per 4 threads
    auto prefetch = [&](const Cur& c) {
      if ((kc & 3) == 0) {
        const float* xb = c.x0 + c.kb * 64;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          int64_t row = c.row_base + p * 32;
          row = row < last_row ? row : last_row;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + row * d.x_row_stride));
        }
      }
    };
    int64_t staged_b = -1;
    // when a block belongs to another batch item than the staged one, (re)load that item's scale|shift rows
    auto stage_affine = [&](int64_t b) {
      if constexpr (kAffine) {
        if (b != staged_b) {
          asm volatile("bar.sync 2, %0;" ::"n"(TC_PRODUCERS) : "memory");  // every producer is done reading the old rows
          const float* pa = d.pro_a + b * d.pro_batch_stride;
          const float* pb = d.pro_b + b * d.pro_batch_stride;
          for (int k = pt * 4; k < K; k += TC_PRODUCERS * 4) {
            *reinterpret_cast<float4*>(aff_s + k) = __ldg(reinterpret_cast<const float4*>(pa + k));
            *reinterpret_cast<float4*>(aff_s + K + k) = __ldg(reinterpret_cast<const float4*>(pb + k));
          }
          asm volatile("bar.sync 2, %0;" ::"n"(TC_PRODUCERS) : "memory");
          staged_b = b;
        }
      }
    };
    // transform + bf16 hi/lo split + swizzled store of one k-block into stage s (+ s+1 for the 32-k geometry)
    auto process = [&](const float4(&x)[4][2], int kb) {
      float sc[8], sh[8];
      if constexpr (kAffine) {
        // the folded norm affine of this item was staged in shared memory at the item boundary: two 29-cycle LDS.128
        // pairs instead of L2-latency global loads in the dependency chain of every block
        const int k0 = kb * 64 + kc * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(aff_s + k0), a1 = *reinterpret_cast<const float4*>(aff_s + k0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(aff_s + K + k0), b1 = *reinterpret_cast<const float4*>(aff_s + K + k0 + 4);
        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
      }
      const int my_s = s + sub;
      mbar_wait(bar_empty + 8 * my_s, ph ^ 1);
      uint8_t* a_hi = sm + my_s * TC_STAGE_BYTES;
      uint8_t* a_lo = a_hi + TC_A_PART;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int r = p * 32 + r0;
        const float v[8] = {x[p][0].x, x[p][0].y, x[p][0].z, x[p][0].w, x[p][1].x, x[p][1].y, x[p][1].z, x[p][1].w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float u0 = v[i], u1 = v[i + 1];
          if constexpr (kAffine) {
            u0 = fmaf(u0, sc[i], sh[i]);
            u1 = fmaf(u1, sc[i + 1], sh[i + 1]);
            u0 = u0 > 0.f ? u0 : u0 * pslope;
            u1 = u1 > 0.f ? u1 : u1 * pslope;
          }
          // packed conversions (one F2FP per pair) keep the slow XU pipe out of the loop
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(u0, u1);
          const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
          const float r0f = u0 - __uint_as_float(hb << 16);
          const float r1f = u1 - __uint_as_float(hb & 0xFFFF0000u);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0f, r1f);
          hi[i >> 1] = hb;
          lo[i >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        const uint32_t off = Cfg::swz((uint32_t)r, cch);
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(bar_full + 8 * my_s);
      s += Cfg::kSub;
      if (s == TC_STAGES) { s = 0; ph ^= 1; }
    };

    float4 xa[4][2], xb2[4][2];
    Cur cur, nxt, pf;            // block being transformed, block being loaded, block being prefetched
    cur.t = blockIdx.x; cur.kb = 0;
    decode(cur);
    pf = cur;
    for (int i = 0; i < TC_PF_DIST && pf.t < n_tiles; ++i) {
      if (i > 0) prefetch(pf);
      advance(pf);
    }
    if (cur.t < n_tiles) issue(xa, cur);
    while (cur.t < n_tiles) {
      nxt = cur;
      advance(nxt);
      if (nxt.t < n_tiles) issue(xb2, nxt);
      if (pf.t < n_tiles) { prefetch(pf); advance(pf); }
      stage_affine(cur.b);
      process(xa, cur.kb);
      if (nxt.t >= n_tiles) break;
      cur = nxt;
      advance(cur);
      if (cur.t < n_tiles) issue(xa, cur);
      if (pf.t < n_tiles) { prefetch(pf); advance(pf); }
      stage_affine(nxt.b);
      process(xb2, nxt.kb);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- weight packing
// W [M, K] fp32 -> per (n-half, stage k-block): [hi 256xBK bf16 | lo 256xBK bf16], each already in the K-major
// swizzled shared-memory image of the chosen stage geometry, so the GEMM loads it with a plain bulk copy.
template <int BK>
__global__ void pack_weights_kernel(const float* __restrict__ W, int64_t ldw, int64_t M, int64_t K, uint8_t* __restrict__ out) {
  using Cfg = TcCfg<BK>;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * K) return;
  const int64_t n = i / K, k = i % K;
  const int64_t nh = n / TC_BN, r = n % TC_BN, kb = k / BK, kk = k % BK;
  const int64_t KB = K / BK;
  const float w = W[n * ldw + k];
  const __nv_bfloat16 h = __float2bfloat16_rn(w);
  const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
  const size_t tile = ((size_t)(nh * KB + kb) * 2) * Cfg::kBPart;
  const size_t off = (size_t)Cfg::swz((uint32_t)r, (uint32_t)(kk >> 3)) + (size_t)(kk & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(out + tile + off) = h;
  *reinterpret_cast<__nv_bfloat16*>(out + tile + Cfg::kBPart + off) = l;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// stage geometry, fixed per process (the packed weight image depends on it): PS_TC_BK=64|32
int tc_bk() {
  static EnvInt env;
  return env.get("PS_TC_BK", 32) == 64 ? 64 : 32;
}

// which tcgen05 kernel serves eligible shapes, fixed per process (the packed weight image depends on it):
// PS_TC_KERNEL=pair (default; CTA-pair kernel of ps_gemm_pair.cu) | single (the one-CTA kernel of this file)
int gemm_pair_launch(const ps_gemm_t& d, cudaStream_t s);
int gemm_pair_pack(const float* W, int64_t ldw, int64_t M, int64_t K, void* packed, cudaStream_t s);
// few output channels (M <= 128) with the frames on the MMA's M side and the weights resident (ps_gemm_rows.cu)
bool gemm_rows_eligible(const ps_gemm_t& d);
int gemm_rows_launch(const ps_gemm_t& d, const void* wimg, cudaStream_t s);
int64_t gemm_rows_image_bytes(int64_t M, int64_t K);
int gemm_rows_pack(const float* W, int64_t ldw, int64_t M, int64_t K, void* packed, cudaStream_t s);
bool tc_pair() {
  static std::atomic<int> mode{-1};
  int m = mode.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = getenv("PS_TC_KERNEL");
    m = (e && e[0] == 's') ? 0 : 1;
    mode.store(m, std::memory_order_relaxed);
  }
  return m == 1;
}

bool gemm_tc_eligible(const ps_gemm_t& d) {
  if (!d.W_packed) return false;
  const bool pair = tc_pair();
  // the pair kernel zero-pads the channels to whole 256-blocks (any multiple of 32 goes), takes the mask-apply prologue
  // and overlapping rows (framed filterbank / STFT analysis views); the single-CTA kernel does none of these
  // K == 32 (one shared-memory stage per tile): pair kernel without a prologue only
  const bool k32 = pair && d.K == 32 && d.pro_mode == PS_PRO_NONE;
  if (d.M % (pair ? 32 : TC_BN) != 0 || (d.K % 64 != 0 && !k32)) return false;
  if (!(d.pro_mode == PS_PRO_NONE || d.pro_mode == PS_PRO_AFFINE || (pair && d.pro_mode == PS_PRO_MASK))) return false;
  if (d.pro_mode == PS_PRO_MASK && (!al16(d.X2) || !(d.pro_act == PS_ACT_NONE || d.pro_act == PS_ACT_RELU || d.pro_act == PS_ACT_SIGMOID))) return false;
  if (d.pro_mode == PS_PRO_AFFINE && !(d.pro_act == PS_ACT_NONE || d.pro_act == PS_ACT_PRELU || (pair && d.pro_act == PS_ACT_TANH))) return false;
  if (!(d.epi_act == PS_ACT_NONE || d.epi_act == PS_ACT_RELU || d.epi_act == PS_ACT_PRELU)) return false;
  if ((d.bias && !al16(d.bias)) || (d.bias_batch && !al16(d.bias_batch))) return false;
  if ((d.x_row_stride & 3) || (d.x_batch_stride & 3) || !al16(d.X) || (!pair && d.x_row_stride < d.K)) return false;
  if ((d.y_row_stride & 3) || (d.y_batch_stride & 3) || !al16(d.Y)) return false;
  if (d.pro_mode == PS_PRO_AFFINE && ((d.pro_batch_stride & 3) || !al16(d.pro_a) || !al16(d.pro_b) || d.K > TC_MAXK)) return false;
  if (d.residual && ((d.res_row_stride & 3) || (d.res_batch_stride & 3) || !al16(d.residual))) return false;
  if (!al16(d.W_packed)) return false;
  return true;
}

template <bool kAffine, int BK>
static int launch_variant(const ps_gemm_t& d, cudaStream_t s, int dev, int64_t grid, int64_t n_rt, int64_t n_nh, int64_t n_tiles) {
  static SmemOnce<1> once;  // per instantiation and device
  if (int rc = once.ensure(dev, 0, gemm_tc_kernel<kAffine, BK>, TcCfg<BK>::kSmem, "cudaFuncSetAttribute(gemm_tc_kernel)")) return rc;
  gemm_tc_kernel<kAffine, BK><<<(unsigned)grid, TC_THREADS, TcCfg<BK>::kSmem, s>>>(d, n_rt, n_nh, n_tiles);
  PS_CHECK_LAUNCH("gemm_tc_kernel");
  return PS_OK;
}

// bytes of the CTA-pair image (channels padded to whole 256-channel blocks); the rows image (ps_gemm_rows.cu), when the
// shape has one, follows it in the same buffer
static inline int64_t pair_image_bytes(int64_t M, int64_t K) { return (M + 255) / 256 * 256 * K * 4; }

int gemm_tc_launch(const ps_gemm_t& d, cudaStream_t s) {
  if (tc_pair()) {
    if (gemm_rows_eligible(d)) return gemm_rows_launch(d, reinterpret_cast<const uint8_t*>(d.W_packed) + pair_image_bytes(d.M, d.K), s);
    return gemm_pair_launch(d, s);
  }
  int dev = 0, sms = 0;
  if (int rc = current_device(&dev)) return rc;
  if (int rc = sm_count_of(dev, &sms)) return rc;
  const bool affine = d.pro_mode == PS_PRO_AFFINE;
  const int64_t n_rt = cdiv(d.rows, TC_BM), n_nh = d.M / TC_BN;
  const int64_t n_tiles = d.batch * n_rt * n_nh;
  if (n_tiles >= (1LL << 31)) return PS_ERR_UNSUPPORTED;
  const int64_t grid = n_tiles < sms ? n_tiles : sms;
  if (tc_bk() == 64)
    return affine ? launch_variant<true, 64>(d, s, dev, grid, n_rt, n_nh, n_tiles) : launch_variant<false, 64>(d, s, dev, grid, n_rt, n_nh, n_tiles);
  return affine ? launch_variant<true, 32>(d, s, dev, grid, n_rt, n_nh, n_tiles) : launch_variant<false, 32>(d, s, dev, grid, n_rt, n_nh, n_tiles);
}

}  // namespace ps

extern "C" int64_t ps_gemm_packed_bytes(int64_t M, int64_t K) {
  if (M <= 0 || K <= 0 || (K % 64 != 0 && !(K == 32 && ps::tc_pair()))) return 0;
  if (ps::tc_pair()) {  // padded to whole 256-channel blocks (+ the resident image of the few-channel kernel)
    if (M % 32 != 0) return 0;
    return (M + 255) / 256 * 256 * K * 4 + (K % 64 == 0 ? ps::gemm_rows_image_bytes(M, K) : 0);
  }
  if (M % ps::TC_BN != 0) return 0;
  return M * K * 4;  // bf16 hi + bf16 lo
}

extern "C" int ps_gemm_pack_weights(const float* W, int64_t w_row_stride, int64_t M, int64_t K, void* packed, void* stream) {
  PS_REQUIRE(W && packed && w_row_stride >= K);
  if (ps_gemm_packed_bytes(M, K) == 0) return PS_ERR_UNSUPPORTED;
  if (ps::tc_pair()) {
    if (int rc = ps::gemm_pair_pack(W, w_row_stride, M, K, packed, (cudaStream_t)stream)) return rc;
    if (K % 64 == 0 && ps::gemm_rows_image_bytes(M, K) > 0)
      return ps::gemm_rows_pack(W, w_row_stride, M, K, reinterpret_cast<uint8_t*>(packed) + (M + 255) / 256 * 256 * K * 4, (cudaStream_t)stream);
    return PS_OK;
  }
  const unsigned blocks = (unsigned)ps::cdiv(M * K, 256);
  if (ps::tc_bk() == 64)
    ps::pack_weights_kernel<64><<<blocks, 256, 0, (cudaStream_t)stream>>>(W, w_row_stride, M, K, reinterpret_cast<uint8_t*>(packed));
  else
    ps::pack_weights_kernel<32><<<blocks, 256, 0, (cudaStream_t)stream>>>(W, w_row_stride, M, K, reinterpret_cast<uint8_t*>(packed));
  PS_CHECK_LAUNCH("pack_weights_kernel");
  return PS_OK;
}
