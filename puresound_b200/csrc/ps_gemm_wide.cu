// Wide CTA-pair tcgen05 / TMEM GEMM for the 512-channel 1x1 convolutions of the TCN stack (sm_100a, cta_group::2).
//
//   Y[b,f,c] = epi( sum_k pro(X[b,f,k]) * W[c,k] ),  fp32 in HBM, fp32-grade result (3xBF16 split, see ps_gemm_tc.cu).
//
// Same operand roles as gemm_pair_kernel (ps_gemm_pair.cu) - A = weights (256 channels per MMA across the pair), B =
// activations (frames on N), D = [128 channels x frames] per CTA and 256-channel block, transpose-free epilogue - rebuilt
// around what the round-2 measurements showed (profiles/r02_gemm_notes.md):
//
//   * the tensor pipe is NOT the limiter: with every tcgen05.mma removed the 128-frame kernel still took 0.31 of its
//     0.36 ms, and a bare MMA loop runs at 100 % with or without the weight stream and the producer stores
//     (profiles/microbench/mma_rate.cu).  What costs time is LATENCY in the feeding chains: producer threads that wait for
//     their own global loads one stage ahead, a stage that must be refilled with weights AND activations before it can be
//     reused, an epilogue that waits for each chunk's residual loads.
//
// So here every chain gets its own ring and nobody waits for a global load in registers:
//   R ring (3 x 16 KB)  raw fp32 activations [128 frames x 32 k], filled by TENSOR-MAP TMA (cp.async.bulk.tensor, 128-byte
//                       swizzle) from a dedicated warp that runs as far ahead as the ring allows - no registers, no
//                       address math, HBM latency hidden by ring depth;
//   X ring (3 x 16 KB)  bf16 hi / lo operand tiles in the UMMA layout, written by the 8 producer warps
//                       (LDS -> affine + PReLU -> split -> STS), who no longer touch global memory for the operand;
//   W ring (4 x 32 KB)  pre-packed weight stages (cp.async.bulk), 4 deep because their refill is an L2 round trip.
// A tile is 256 frames x 512 channels (MMA N = 256): every weight stage pulled from L2 serves 256 frames, halving the
// L2 -> shared-memory weight re-stream of the 128-frame kernel (2.1 GB per launch at cfg2 for 0.5 MB of weights).
// Two 256-channel blocks x 256 frames are all 512 TMEM columns, so the two BLOCKS hand over separately (tfull / tempty
// per block) and the MMA order at the tile boundaries is skewed - the last H stages of a tile are issued block 0 first
// (its drain starts while block 1 still computes), the first H stages of the next tile likewise (block 0 only needs block
// 0 drained).  The epilogue prefetches the residual of chunk i+1 into registers while chunk i is stored.
//
// Serves: M % 512 == 0, K % 32 == 0 (K >= 64), prologue none / folded affine + PReLU, epilogue bias / per-item bias /
// activation / residual / Welford partials.  Everything else (M = 128/256, LayerNorm epilogue, mask prologue, K = 32,
// small row counts) stays on gemm_pair_kernel.  Weight image: the NB = 2 image of gemm_pair_pack, unchanged.
#include "ps_tc_ptx.cuh"

// Bottleneck experiments (PS_WIDE_DBG: remove one stream of work at a time; results are garbage) exist only in builds with
// -DPS_EXPERIMENTS (python -m puresound_b200.build --experiments -> libpuresound_b200_exp.so, never loaded by default).
#ifdef PS_EXPERIMENTS
#define WD_DBG(bit) ((dbg & (bit)) != 0)
#else
#define WD_DBG(bit) false
#endif

namespace ps {

constexpr int WD_FRAMES = 256;      // frames per pair tile (MMA N)
constexpr int WD_FR_CTA = 128;      // frames staged by each CTA
constexpr int WD_BK = 32;           // k per shared-memory stage (64-byte swizzle rows)
constexpr int WD_STAGES = 4;        // ring depth (power of two)
constexpr int WD_HOLD = 3;          // stages held across the block skew at the tile boundaries
constexpr int WD_XPART = WD_FR_CTA * WD_BK * 2;   // 8 KB: activation hi (or lo) of one stage
constexpr int WD_WBLK = 128 * WD_BK * 2;          // 8 KB: one 128-channel weight block, hi (or lo), of one stage
constexpr int WD_STAGE = 2 * WD_XPART + 4 * WD_WBLK;  // 48 KB
constexpr int WD_WBYTES = 4 * WD_WBLK;                // weight bytes per stage per CTA
constexpr int WD_EPI_WARPS = 8;
constexpr int WD_PRODUCERS = 256, WD_EPI = WD_EPI_WARPS * 32;
constexpr int WD_THREADS = 64 + WD_EPI + WD_PRODUCERS;
constexpr int WD_MAXK = 1024;
constexpr int WD_AFF_BYTES = 2 * WD_MAXK * 4;
constexpr int WD_CW = 16;           // frames (TMEM columns) per epilogue chunk
constexpr int WD_SMEM = WD_STAGES * WD_STAGE + WD_AFF_BYTES + 1024 /*align*/ + 1024 /*barriers, scratch*/;

__host__ __device__ constexpr uint32_t wd_swz(uint32_t r, uint32_t c) { return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4); }

// K-major, SWIZZLE_64B, 8-row atoms 512 B apart
__device__ __forceinline__ uint64_t wd_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// D=f32, A=B=bf16, K-major both, M=256 (pair), N=256
constexpr uint32_t WD_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(WD_FRAMES >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

struct WideTile {
  uint32_t b, rt, nh;
};
__device__ __forceinline__ WideTile wd_tile(int64_t t, int64_t n_rt, int64_t n_nh) {
  const uint32_t tt = (uint32_t)t, nrt = (uint32_t)n_rt, nnh = (uint32_t)n_nh;
  const uint32_t q = tt / nnh;
  WideTile c;
  c.nh = tt - q * nnh;
  c.b = q / nrt;
  c.rt = q - c.b * nrt;
  return c;
}

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// (a, b) * (s0, s1) + (t0, t1) in one packed fp32 FMA (FFMA2); each half rounds like fmaf
__device__ __forceinline__ void ffma2(float& a, float& b, float s0, float s1, float t0, float t1) {
  uint64_t d = pack2(t0, t1);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(pack2(a, b)), "l"(pack2(s0, s1)));
  unpack2(d, a, b);
}

// Epilogue of one WD_CW-frame chunk of one channel (thread): out = act(acc + bias) (+ residual), strided scalar stores that
// are coalesced across the warp's 32 channels, and (kStats) shifted sums about a pivot for the gLN statistics.
template <int ACT, bool kRes, bool kFull, bool kStats>
__device__ __forceinline__ void wd_epi_chunk(const float (&v)[WD_CW], float bsum, float eslope, float* yp, int64_t ystride, const float* rp,
                                             int64_t rstride, int nj, bool first, float& piv, float& ssum, float& ssq, int act = ACT) {
  float r[WD_CW];
  if constexpr (kRes) {  // all residual loads of the chunk in flight together
    const float* p = rp;
#pragma unroll
    for (int j = 0; j < WD_CW; ++j) {
      r[j] = (kFull || j < nj) ? __ldg(p) : 0.f;
      p += rstride;
    }
  }
#pragma unroll
  for (int j = 0; j < WD_CW; ++j) {
    float o = v[j] + bsum;
    if (ACT == PS_ACT_NONE) {
    } else if (act == PS_ACT_PRELU) o = o > 0.f ? o : o * eslope;
    else if (act == PS_ACT_RELU) o = (o != o) ? o : fmaxf(o, 0.f);
    if constexpr (kRes) o += r[j];
    if constexpr (kStats) {
      if (j == 0 && first) piv = o;  // statistics pivot: this thread's first output of the block
    }
    if (kFull || j < nj) {
      *yp = o;
      if constexpr (kStats) {
        const float dv = o - piv;
        ssum += dv;
        ssq = fmaf(dv, dv, ssq);
      }
    }
    yp += ystride;
  }
}

// PRO: PS_PRO_NONE or PS_PRO_AFFINE (norm affine + PReLU; "no activation" = slope 1)
template <int PRO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WD_THREADS, 1)
    gemm_wide_kernel(const ps_gemm_t d, const int64_t n_rt, const int64_t n_nh, const int64_t n_tiles, const int dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* aff_s = reinterpret_cast<float*>(sm + WD_STAGES * WD_STAGE);
  const uint32_t bars = base + WD_STAGES * WD_STAGE + WD_AFF_BYTES;
  // barrier map (8 B each): full @0, empty @64, pfull @128, tfull[0..1] @192, tempty[0..1] @208
  const uint32_t bar_full = bars, bar_empty = bars + 64, bar_pfull = bars + 128, bar_tfull = bars + 192, bar_tempty = bars + 208;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + WD_STAGES * WD_STAGE + WD_AFF_BYTES + 224);
  Wf* wf_s = reinterpret_cast<Wf*>(sm + WD_STAGES * WD_STAGE + WD_AFF_BYTES + 256);  // [2 blocks][epilogue warps]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int64_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int KB = (int)(d.K / WD_BK);

  if (tid == 0) {
    for (int s = 0; s < WD_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, WD_PRODUCERS / 32 + 1);  // the 8 producer warps + the weight copy
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_pfull + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 2 * WD_EPI_WARPS);  // one arrival per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  pdl_trigger();  // the next kernel of the stream may start its prologue on SMs this grid leaves idle / frees
  if (warp == 1) tmem_alloc_pair(smem_u32((const void*)tmem_ptr_s), 512);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  pdl_wait();  // every earlier kernel of the stream is complete: operands, folded norms, residual are in place

  if (warp == 0) {
    // ===================== weight loader =====================
    uint32_t gs = 0;  // stages issued so far (ring position = gs & 3, phase = (gs >> 2) & 1)
    const uint8_t* wp = reinterpret_cast<const uint8_t*>(d.W_packed);
    for (int64_t t = pair; t < n_tiles; t += n_pairs) {
      const WideTile tc = wd_tile((dbg & 0x200) ? n_tiles - 1 - t : t, n_rt, n_nh);
      const uint8_t* src = wp + (size_t)(tc.nh * 2 + rank) * KB * WD_WBYTES;
      for (int kb = 0; kb < KB; ++kb, ++gs) {
        const uint32_t s = gs & (WD_STAGES - 1), ph = (gs >> 2) & 1u;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_full + 8 * s, WD_WBYTES);
          if (WD_DBG(1))  // experiment: no weight traffic
            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(bar_full + 8 * s), "r"((uint32_t)WD_WBYTES) : "memory");
          else
            bulk_g2s(base + s * WD_STAGE + 2 * WD_XPART, src + (size_t)kb * WD_WBYTES, WD_WBYTES, bar_full + 8 * s);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader) =====================
      uint32_t gs = 0;
      int64_t it = 0;
      const int h1 = KB < WD_HOLD ? KB : WD_HOLD;         // head: stages [0, h1)
      const int t0 = (KB - WD_HOLD) > h1 ? (KB - WD_HOLD) : h1;  // tail: stages [t0, KB); mid: [h1, t0)
      auto wait_stage = [&](int kb) {
        const uint32_t g = gs + (uint32_t)kb, s = g & (WD_STAGES - 1), ph = (g >> 2) & 1u;
        mbar_wait(bar_full + 8 * s, ph);   // this CTA's stage: producers + weight copy
        mbar_wait(bar_pfull + 8 * s, ph);  // the peer's stage (relayed)
        tc_fence_after();
      };
      // MMAs of block mb for stage kb; `release` frees the stage, `done` hands the block's accumulators to the epilogues
      auto issue = [&](int kb, int mb, bool release, bool done) {
        const uint32_t g = gs + (uint32_t)kb, s = g & (WD_STAGES - 1);
        if (elect_one()) {
          const uint32_t sa = base + s * WD_STAGE;
          const uint64_t x_hi = wd_desc(sa), x_lo = wd_desc(sa + WD_XPART);
          const uint64_t w_hi = wd_desc(sa + 2 * WD_XPART + mb * WD_WBLK);
          const uint64_t w_lo = wd_desc(sa + 2 * WD_XPART + (2 + mb) * WD_WBLK);
          const uint32_t dd = tmem_base + (uint32_t)(mb * WD_FRAMES);
#pragma unroll
          for (int k = 0; k < WD_BK / 16; ++k) {
            if (WD_DBG(16)) break;  // experiment: no tensor work
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            // small cross terms first, the dominant hi*hi last
            umma_bf16_pair(dd, w_lo + ko, x_hi + ko, WD_IDESC, (kb | k) != 0);
            umma_bf16_pair(dd, w_hi + ko, x_lo + ko, WD_IDESC, 1);
            umma_bf16_pair(dd, w_hi + ko, x_hi + ko, WD_IDESC, 1);
          }
          if (release) umma_commit_pair(bar_empty + 8 * s);
          if (done) umma_commit_pair(bar_tfull + 8 * mb);
        }
        __syncwarp();
      };
      for (int64_t t = pair; t < n_tiles; t += n_pairs, ++it) {
        const uint32_t aph = (uint32_t)(it & 1);
        // ---- head: block 0 of the first stages while block 1 of the previous tile may still drain
        mbar_wait(bar_tempty + 0, aph ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < h1; ++kb) {
          wait_stage(kb);
          issue(kb, 0, false, kb == KB - 1);
        }
        mbar_wait(bar_tempty + 8, aph ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < h1; ++kb) issue(kb, 1, true, kb == KB - 1);
        // ---- middle: both blocks per stage
        for (int kb = h1; kb < t0; ++kb) {
          wait_stage(kb);
          issue(kb, 0, false, false);
          issue(kb, 1, true, false);
        }
        // ---- tail: block 0 of the last stages first, so its drain starts while block 1 still computes
        for (int kb = t0; kb < KB; ++kb) {
          wait_stage(kb);
          issue(kb, 0, false, kb == KB - 1);
        }
        for (int kb = t0; kb < KB; ++kb) issue(kb, 1, true, kb == KB - 1);
        gs += (uint32_t)KB;
      }
    } else {
      // ===================== stage relay (peer) =====================
      // tells the leader that this CTA's half of a stage (its 128 activation frames and its weight blocks) landed
      uint32_t gs = 0;
      for (int64_t t = pair; t < n_tiles; t += n_pairs) {
        for (int kb = 0; kb < KB; ++kb, ++gs) {
          const uint32_t s = gs & (WD_STAGES - 1), ph = (gs >> 2) & 1u;
          mbar_wait(bar_full + 8 * s, ph);
          if (elect_one()) mbar_arrive_remote(bar_pfull + 8 * s, 0);
          __syncwarp();
        }
      }
    }
  } else if (warp < 2 + WD_EPI_WARPS) {
    // ===================== epilogue (warps 2..9) =====================
    // Block by block (block 0 completes first).  Within a block, warp group g = (warp-2)/4 takes the 32-frame units with
    // u % 2 == g, so two warps drain each TMEM lane quarter concurrently.
    constexpr int G = WD_EPI_WARPS / 4;
    constexpr int NCH = WD_FRAMES / WD_CW;  // chunks per 256-channel block
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
      const WideTile tc = wd_tile((dbg & 0x200) ? n_tiles - 1 - t : t, n_rt, n_nh);
      const uint32_t aph = (uint32_t)(it & 1);
      const int64_t frame0 = (int64_t)tc.rt * WD_FRAMES;
      const int nvalid = (int)((d.rows - frame0) < WD_FRAMES ? (d.rows - frame0) : WD_FRAMES);
      const int64_t ch0 = (int64_t)tc.nh * 512 + rank * 128;  // + mb*256 + cl
      float* yb = d.Y + tc.b * d.y_batch_stride + frame0 * ystride + ch0 + cl;
      const float* rb = d.residual ? d.residual + tc.b * d.res_batch_stride + frame0 * rstride + ch0 + cl : nullptr;
      float piv[2], ssum[2], ssq[2], cnt[2];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) { piv[mb] = 0.f; ssum[mb] = 0.f; ssq[mb] = 0.f; cnt[mb] = 0.f; }
#pragma unroll 1
      for (int mb = 0; mb < 2; ++mb) {
        const int64_t ch = ch0 + mb * 256 + cl;
        if (rb) {
          // this block's MMAs are still in flight: pull the residual lines of this warp's chunks (32 channels = 128 B per
          // frame) into L2, one frame per lane
#pragma unroll
          for (int u = 0; u < NCH / 2; ++u) {
            if (u % G != g) continue;
            const int f = u * 32 + lane;
            if (f < nvalid) asm volatile("prefetch.global.L2 [%0];" ::"l"(rb - lane + f * rstride + mb * 256));
          }
        }
        float bsum = bias ? __ldg(bias + ch) : 0.f;
        if (bias_batch) bsum += __ldg(bias_batch + tc.b * d.M + ch);
        mbar_wait_relaxed(bar_tfull + 8 * mb, aph);
        tc_fence_after();
        float pv = 0.f, sm1 = 0.f, sq1 = 0.f, cn = 0.f;
        if (rb && !want_stats && epi_act == PS_ACT_NONE && nvalid == WD_FRAMES && !WD_DBG(4) && !(dbg & 0x100)) {
          // out_conv of a full tile (bias + residual, no activation, no statistics): the residual of chunk i + 1 is
          // requested BEFORE chunk i is read from TMEM, added and stored, so its (L2) latency passes behind that work
          // instead of being waited for once per chunk (the generic loop below issues a chunk's 16 loads and uses them at once)
          float r[WD_CW], rn[WD_CW];
          int c = g * 2;
          {
            const float* p = rb + (int64_t)(c * WD_CW) * rstride + mb * 256;
#pragma unroll
            for (int j = 0; j < WD_CW; ++j) { r[j] = __ldg(p); p += rstride; }
          }
#pragma unroll 1
          while (c < NCH) {
            const int c2 = c + ((c & 1) ? (G - 1) * 2 + 1 : 1);
            if (c2 < NCH) {
              const float* p = rb + (int64_t)(c2 * WD_CW) * rstride + mb * 256;
#pragma unroll
              for (int j = 0; j < WD_CW; ++j) { rn[j] = __ldg(p); p += rstride; }
            }
            float v[WD_CW];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * WD_FRAMES + c * WD_CW), v);
            float* yp = yb + (int64_t)(c * WD_CW) * ystride + mb * 256;
#pragma unroll
            for (int j = 0; j < WD_CW; ++j) {
              *yp = v[j] + bsum + r[j];
              yp += ystride;
            }
#pragma unroll
            for (int j = 0; j < WD_CW; ++j) r[j] = rn[j];
            c = c2;
          }
        } else
        // real loops (not unrolled): the chunk body exists once per variant, which keeps the kernel inside the
        // instruction cache
#pragma unroll 1
        for (int c = g * 2; c < NCH; c += (c & 1) ? (G - 1) * 2 + 1 : 1) {
          const int nj = nvalid - c * WD_CW;  // valid frames of this chunk (warp-uniform)
          if (nj <= 0) break;
          float v[WD_CW];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * WD_FRAMES + c * WD_CW), v);
          float* yp = yb + (int64_t)(c * WD_CW) * ystride + mb * 256;
          const float* rp = rb ? rb + (int64_t)(c * WD_CW) * rstride + mb * 256 : nullptr;
          const bool first = cn == 0.f;
          cn += (float)(nj < WD_CW ? nj : WD_CW);
          if (WD_DBG(4)) continue;  // experiment: no epilogue stores / residual loads
          if (nj >= WD_CW) {
            if (want_stats) {
              if (rp) wd_epi_chunk<-1, true, true, true>(v, bsum, eslope, yp, ystride, rp, rstride, WD_CW, first, pv, sm1, sq1, epi_act);
              else if (epi_act == PS_ACT_NONE) wd_epi_chunk<PS_ACT_NONE, false, true, true>(v, bsum, eslope, yp, ystride, rp, rstride, WD_CW, first, pv, sm1, sq1);
              else wd_epi_chunk<-1, false, true, true>(v, bsum, eslope, yp, ystride, rp, rstride, WD_CW, first, pv, sm1, sq1, epi_act);
            } else {
              if (rp && epi_act == PS_ACT_NONE) wd_epi_chunk<PS_ACT_NONE, true, true, false>(v, bsum, eslope, yp, ystride, rp, rstride, WD_CW, first, pv, sm1, sq1);
              else if (rp) wd_epi_chunk<-1, true, true, false>(v, bsum, eslope, yp, ystride, rp, rstride, WD_CW, first, pv, sm1, sq1, epi_act);
              else wd_epi_chunk<-1, false, true, false>(v, bsum, eslope, yp, ystride, rp, rstride, WD_CW, first, pv, sm1, sq1, epi_act);
            }
          } else {
            if (rp) wd_epi_chunk<-1, true, false, true>(v, bsum, eslope, yp, ystride, rp, rstride, nj, first, pv, sm1, sq1, epi_act);
            else wd_epi_chunk<-1, false, false, true>(v, bsum, eslope, yp, ystride, rp, rstride, nj, first, pv, sm1, sq1, epi_act);
          }
        }
        // all TMEM reads of this block are complete (tcgen05.wait::ld): hand it back to the leader's MMA thread
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(bar_tempty + 8 * mb, 0);
        piv[mb] = pv; ssum[mb] = sm1; ssq[mb] = sq1; cnt[mb] = cn;
      }
      if (want_stats) {
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          Wf mine;
          mine.n = cnt[mb]; mine.mean = 0.f; mine.m2 = 0.f;
          if (mine.n > 0.f) {
            const float md = ssum[mb] / mine.n;
            mine.mean = piv[mb] + md;
            mine.m2 = fmaxf(ssq[mb] - ssum[mb] * md, 0.f);
          }
          Wf w = wf_warp_reduce(mine);
          if (lane == 0) wf_s[mb * WD_EPI_WARPS + (warp - 2)] = w;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WD_EPI) : "memory");
        // slot layout of ps_gemm_stats_slots (128-frame rows x 128-channel columns): this tile covers frame rows 2*rt
        // (gets the whole tile's partial) and 2*rt + 1 (gets an empty one)
        const int64_t slots_m = d.M / 128;
        const int64_t n_rt128 = (d.rows + 127) / 128;
        const int64_t slots = n_rt128 * slots_m;
        if (et < 2) {
          const int mb = et;
          Wf tot = wf_s[mb * WD_EPI_WARPS];  // fixed merge order over the epilogue warps -> deterministic
#pragma unroll
          for (int w = 1; w < WD_EPI_WARPS; ++w) tot = wf_merge(tot, wf_s[mb * WD_EPI_WARPS + w]);
          const int64_t col = tc.nh * 4 + mb * 2 + rank;
          float* o = d.stats_partials + (tc.b * slots + (int64_t)(2 * tc.rt) * slots_m + col) * 3;
          o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
          if ((int64_t)(2 * tc.rt + 1) < n_rt128) {
            o += slots_m * 3;
            o[0] = 0.f; o[1] = 0.f; o[2] = 0.f;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WD_EPI) : "memory");
      }
    }
  } else {
    // ===================== activation producers (last 8 warps) =====================
    // All 256 threads fill one stage (this CTA's 128 frames x 32 k): 4 threads cover the 32 k of a frame (coalesced
    // 128 B), a warp covers 8 consecutive frames, so each 8-lane phase of a 128-bit shared store lands on two consecutive
    // swizzle rows = 32 distinct banks.  Every thread handles its 8-k chunk for frames r and r + 64, so the affine is read
    // once per 16 elements.
    const int pt = tid - (64 + WD_EPI);
    const uint32_t chunk = (uint32_t)(pt & 3);
    const int r0 = pt >> 2;  // 0..63
    const int kofs = (int)chunk * 8;
    const int K = (int)d.K;
    const float pslope = (d.pro_act == PS_ACT_PRELU && d.pro_slope) ? __ldg(d.pro_slope) : 1.f;  // AFFINE w/o act = slope 1
    // PReLU with 0 < slope <= 1 is max(u, slope * u) - with the NaN-propagating max, so the reference's Inf/NaN probe
    // (base_nn.py:740-777) still sees what it should; other slopes take the select
    const bool slope01 = pslope > 0.f && pslope <= 1.f;  // (slope 0: Inf * 0 would turn +Inf into NaN)
    uint32_t gs = 0;

    struct XBuf {
      float4 v[2][2];
    };
    struct Cur {
      int64_t t;
      int kb;
      const float* x0;   // &X[b][0][kofs]
      int64_t row_base;  // rt*256 + rank*128 + r0
      int64_t b;
    };
    const int64_t last_row = d.rows - 1;
    auto decode = [&](Cur& c) {
      if (c.t < n_tiles) {
        const WideTile tc = wd_tile((dbg & 0x200) ? n_tiles - 1 - c.t : c.t, n_rt, n_nh);
        c.b = tc.b;
        c.row_base = (int64_t)tc.rt * WD_FRAMES + rank * WD_FR_CTA + r0;
        c.x0 = d.X + tc.b * d.x_batch_stride + kofs;
      }
    };
    auto advance = [&](Cur& c) {
      if (++c.kb == KB) {
        c.kb = 0;
        c.t += n_pairs;
        decode(c);
      }
    };
    // frames past the end of the item re-read its last frame: their accumulator columns are never stored
    auto issue = [&](XBuf& x, const Cur& c) {
      if (WD_DBG(2)) return;  // experiment: no activation loads
      const float* xb = c.x0 + c.kb * WD_BK;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        int64_t row = c.row_base + p * 64;
        row = row < last_row ? row : last_row;
        const float4* src = reinterpret_cast<const float4*>(xb + row * d.x_row_stride);
        x.v[p][0] = __ldg(src);
        x.v[p][1] = __ldg(src + 1);
      }
    };
    int64_t staged_b = -1;
    auto stage_affine = [&](int64_t b) {
      if constexpr (PRO == PS_PRO_AFFINE) {
        if (b != staged_b) {
          asm volatile("bar.sync 2, %0;" ::"n"(WD_PRODUCERS) : "memory");  // every producer is done with the old rows
          const float* pa = d.pro_a + b * d.pro_batch_stride;
          const float* pb = d.pro_b + b * d.pro_batch_stride;
          for (int k = pt * 4; k < K; k += WD_PRODUCERS * 4) {
            *reinterpret_cast<float4*>(aff_s + k) = __ldg(reinterpret_cast<const float4*>(pa + k));
            *reinterpret_cast<float4*>(aff_s + K + k) = __ldg(reinterpret_cast<const float4*>(pb + k));
          }
          asm volatile("bar.sync 2, %0;" ::"n"(WD_PRODUCERS) : "memory");
          staged_b = b;
        }
      }
    };
    auto process = [&](const XBuf& x, int kb) {
      float sc[8], sh[8];
      if constexpr (PRO == PS_PRO_AFFINE) {
        const int k0 = kb * WD_BK + kofs;
        const float4 a0 = *reinterpret_cast<const float4*>(aff_s + k0), a1 = *reinterpret_cast<const float4*>(aff_s + k0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(aff_s + K + k0), b1 = *reinterpret_cast<const float4*>(aff_s + K + k0 + 4);
        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
      }
      const uint32_t s = gs & (WD_STAGES - 1), ph = (gs >> 2) & 1u;
      ++gs;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      uint8_t* x_hi = sm + s * WD_STAGE;
      uint8_t* x_lo = x_hi + WD_XPART;
#pragma unroll
      for (int p = 0; p < 2 && !WD_DBG(8); ++p) {
        const int r = p * 64 + r0;
        float v[8] = {x.v[p][0].x, x.v[p][0].y, x.v[p][0].z, x.v[p][0].w, x.v[p][1].x, x.v[p][1].y, x.v[p][1].z, x.v[p][1].w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float u0 = v[i], u1 = v[i + 1];
          if constexpr (PRO == PS_PRO_AFFINE) {
            ffma2(u0, u1, sc[i], sc[i + 1], sh[i], sh[i + 1]);
            if (slope01) {
              float m0 = u0 * pslope, m1 = u1 * pslope;
              asm("max.NaN.f32 %0, %1, %2;" : "=f"(u0) : "f"(u0), "f"(m0));
              asm("max.NaN.f32 %0, %1, %2;" : "=f"(u1) : "f"(u1), "f"(m1));
            } else {
              u0 = u0 > 0.f ? u0 : u0 * pslope;
              u1 = u1 > 0.f ? u1 : u1 * pslope;
            }
          }
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(u0, u1);
          const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
          const float r0f = u0 - __uint_as_float(hb << 16);
          const float r1f = u1 - __uint_as_float(hb & 0xFFFF0000u);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0f, r1f);
          hi[i >> 1] = hb;
          lo[i >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        const uint32_t off = wd_swz((uint32_t)r, chunk);
        *reinterpret_cast<uint4*>(x_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(x_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor cores of both SMs (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);  // one arrival per producer warp
    };

    // Register double buffer: the loads of stage n+1 are in flight while stage n is transformed.
    XBuf x0, x1;
    Cur pr, ld;
    pr.t = pair; pr.kb = 0;
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

// ---------------------------------------------------------------- dispatch
// Shapes the wide kernel takes over from gemm_pair_kernel: the 512-channel-multiple convs with enough frames for at least
// one full wave of 256-frame tiles.  PS_TC_WIDE=0 keeps everything on the 128-frame kernel (A/B runs).
bool gemm_wide_eligible(const ps_gemm_t& d, int sms) {
  static EnvInt env;
  if (env.get("PS_TC_WIDE", 1) == 0) return false;
  if (d.M % 512 != 0 || d.K % WD_BK != 0 || d.K < 64 || d.K > WD_MAXK) return false;
  if (!(d.pro_mode == PS_PRO_NONE || (d.pro_mode == PS_PRO_AFFINE && (d.pro_act == PS_ACT_NONE || d.pro_act == PS_ACT_PRELU)))) return false;
  if (d.ln_eps > 0.f || d.fin_scale) return false;
  const int64_t n_tiles = d.batch * cdiv(d.rows, WD_FRAMES) * (d.M / 512);
  return n_tiles >= sms / 2 && n_tiles < (1LL << 31);
}

template <int PRO>
static int launch_wide(const ps_gemm_t& d, cudaStream_t s, int dev, int64_t grid, int64_t n_rt, int64_t n_nh, int64_t n_tiles) {
  static SmemOnce<1> once;  // per instantiation and device
  if (int rc = once.ensure(dev, 0, gemm_wide_kernel<PRO>, WD_SMEM, "cudaFuncSetAttribute(gemm_wide_kernel)")) return rc;
#ifdef PS_EXPERIMENTS
  static EnvInt dbg_e;
  int dbg = dbg_e.get("PS_WIDE_DBG", 0);
#else
  int dbg = 0;
#endif
  static EnvInt respf_e;  // A/B: PS_WIDE_RESPF=0 keeps the generic epilogue for the out_conv form (bit 0x100 of dbg)
  if (respf_e.get("PS_WIDE_RESPF", 1) == 0) dbg |= 0x100;
  if (order_reversed()) dbg |= 0x200;  // walk the tiles from the last to the first (see ps_common.cuh: order_reversed)
  cudaError_t le = launch_pdl(gemm_wide_kernel<PRO>, dim3((unsigned)grid), dim3(WD_THREADS), WD_SMEM, s, d, n_rt, n_nh, n_tiles, dbg);
  if (le != cudaSuccess) { set_cuda_error(le, "gemm_wide_kernel"); return PS_ERR_CUDA; }
  return PS_OK;
}

int gemm_wide_launch(const ps_gemm_t& d, cudaStream_t s, int dev, int sms) {
  const int64_t n_rt = cdiv(d.rows, WD_FRAMES), n_nh = d.M / 512;
  const int64_t n_tiles = d.batch * n_rt * n_nh;
  const int64_t max_pairs = sms / 2;
  const int64_t grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
  if (d.pro_mode == PS_PRO_AFFINE) return launch_wide<PS_PRO_AFFINE>(d, s, dev, grid, n_rt, n_nh, n_tiles);
  return launch_wide<PS_PRO_NONE>(d, s, dev, grid, n_rt, n_nh, n_tiles);
}

}  // namespace ps
