// Wide CTA-pair tcgen05 / TMEM GEMM, tensor-map TMA variant (A/B alternative to ps_gemm_wide.cu; PS_TC_WIDE=2).
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
#include "ps_tma.cuh"

// Bottleneck experiments (PS_WIDE_TMA_DBG: remove one stream of work at a time; results are garbage) exist only in builds with
// -DPS_EXPERIMENTS (python -m puresound_b200.build --experiments -> libpuresound_b200_exp.so, never loaded by default).
#ifdef PS_EXPERIMENTS
#define WT_DBG(bit) ((dbg & (bit)) != 0)
#else
#define WT_DBG(bit) false
#endif

namespace ps {

constexpr int WT_FRAMES = 256;      // frames per pair tile (MMA N)
constexpr int WT_FR_CTA = 128;      // frames staged by each CTA
constexpr int WT_BK = 32;           // k per stage (64-byte swizzle rows of bf16; 128-byte rows of raw fp32)
constexpr int WT_WST = 4;           // W ring depth
constexpr int WT_XST = 3;           // X (bf16 operand) ring depth
constexpr int WT_RST = 3;           // R (raw fp32, TMA) ring depth
constexpr int WT_HOLD = 2;          // stages held across the block skew at the tile boundaries (< WT_XST)
constexpr int WT_XPART = WT_FR_CTA * WT_BK * 2;   // 8 KB: activation hi (or lo) of one stage
constexpr int WT_XBYTES = 2 * WT_XPART;           // 16 KB
constexpr int WT_RBYTES = WT_FR_CTA * WT_BK * 4;  // 16 KB raw fp32
constexpr int WT_WBLK = 128 * WT_BK * 2;          // 8 KB: one 128-channel weight block, hi (or lo), of one stage
constexpr int WT_WBYTES = 4 * WT_WBLK;            // 32 KB: weight bytes per stage per CTA
constexpr int WT_EPI_WARPS = 8;
constexpr int WT_PRODUCERS = 256, WT_EPI = WT_EPI_WARPS * 32;
constexpr int WT_THREADS = 64 + WT_EPI + WT_PRODUCERS;  // + loader warp (lane 0: weights, lane 1: activation TMA), MMA / relay warp
constexpr int WT_MAXK = 4096;
constexpr int WT_CW = 16;           // frames (TMEM columns) per epilogue chunk
constexpr int WT_OFF_W = 0, WT_OFF_X = WT_WST * WT_WBYTES, WT_OFF_R = WT_OFF_X + WT_XST * WT_XBYTES, WT_OFF_BAR = WT_OFF_R + WT_RST * WT_RBYTES;
// (the kernel also has 1 KB of static shared memory reserved by the toolchain: 227 KB - 1 KB is the dynamic limit)
// and the 227 KB per-block limit counts that static kilobyte and one more the system reserves: 225 KB of dynamic memory
// is what a launch accepts (measured: 230912 B is refused).  No alignment slack: the dynamic window starts 1 KB-aligned
// (checked at kernel entry, trap otherwise).
constexpr int WT_SMEM = WT_OFF_BAR + 512 /*barriers, scratch*/;
static_assert(WT_SMEM <= 230400, "shared memory budget of one CTA per SM");

__host__ __device__ constexpr uint32_t wt_swz(uint32_t r, uint32_t c) { return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4); }

// K-major, SWIZZLE_64B, 8-row atoms 512 B apart
__device__ __forceinline__ uint64_t wt_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// D=f32, A=B=bf16, K-major both, M=256 (pair), N=256
constexpr uint32_t WT_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(WT_FRAMES >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

struct WideTmaTile {
  uint32_t b, rt, nh;
};
__device__ __forceinline__ WideTmaTile wt_tile(int64_t t, int64_t n_rt, int64_t n_nh) {
  const uint32_t tt = (uint32_t)t, nrt = (uint32_t)n_rt, nnh = (uint32_t)n_nh;
  const uint32_t q = tt / nnh;
  WideTmaTile c;
  c.nh = tt - q * nnh;
  c.b = q / nrt;
  c.rt = q - c.b * nrt;
  return c;
}

__device__ __forceinline__ uint64_t wt_pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void wt_unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// (a, b) * (s0, s1) + (t0, t1) in one packed fp32 FMA (FFMA2); each half rounds like fmaf
__device__ __forceinline__ void wt_ffma2(float& a, float& b, float s0, float s1, float t0, float t1) {
  uint64_t d = wt_pack2(t0, t1);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(wt_pack2(a, b)), "l"(wt_pack2(s0, s1)));
  wt_unpack2(d, a, b);
}

// ring position helpers: g = stages consumed so far on that ring
__device__ __forceinline__ uint32_t wt_ring3_slot(uint32_t g) { return g % 3u; }
__device__ __forceinline__ uint32_t wt_ring3_phase(uint32_t g) { return (g / 3u) & 1u; }

// Epilogue of one WT_CW-frame chunk of one channel (thread): out = act(acc + bias) (+ residual r[]), strided scalar stores
// that are coalesced across the warp's 32 channels, and (kStats) shifted sums about a pivot for the gLN statistics.
template <int ACT, bool kRes, bool kFull, bool kStats>
__device__ __forceinline__ void wt_epi_chunk(const float (&v)[WT_CW], const float (&r)[WT_CW], float bsum, float eslope, float* yp, int64_t ystride,
                                             int nj, bool first, float& piv, float& ssum, float& ssq, int act = ACT) {
#pragma unroll
  for (int j = 0; j < WT_CW; ++j) {
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
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WT_THREADS, 1)  // (registers are granted per 4 warps: 576 threads count as 640 -> 96 per thread)
    gemm_wide_tma_kernel(const ps_gemm_t d, const __grid_constant__ CUtensorMap xmap, const int64_t n_rt, const int64_t n_nh, const int64_t n_tiles,
                     const int dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();  // the swizzled TMA / UMMA tiles need 1 KB alignment and there is no room for slack
  uint8_t* sm = smem_raw;
  const uint32_t bars = base + WT_OFF_BAR;
  // barrier map (8 B each): w_full @0 (4), w_empty @32 (4), x_full @64 (3), x_empty @96 (3), r_full @128 (3), r_empty @160 (3),
  // pfull @192 (4), tfull @224 (2), tempty @240 (2)
  const uint32_t bar_wfull = bars, bar_wempty = bars + 32, bar_xfull = bars + 64, bar_xempty = bars + 96, bar_rfull = bars + 128,
                 bar_rempty = bars + 160, bar_pfull = bars + 192, bar_tfull = bars + 224, bar_tempty = bars + 240;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + WT_OFF_BAR + 256);
  Wf* wf_s = reinterpret_cast<Wf*>(sm + WT_OFF_BAR + 288);  // [2 blocks][epilogue warps]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int64_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int KB = (int)(d.K / WT_BK);

  if (tid == 0) {
    for (int s = 0; s < WT_WST; ++s) {
      mbar_init(bar_wfull + 8 * s, 1);
      mbar_init(bar_wempty + 8 * s, 1);
      mbar_init(bar_pfull + 8 * s, 1);
    }
    for (int s = 0; s < WT_XST; ++s) {
      mbar_init(bar_xfull + 8 * s, WT_PRODUCERS / 32);   // the 8 producer warps
      mbar_init(bar_xempty + 8 * s, 1);
      mbar_init(bar_rfull + 8 * s, 1);
      mbar_init(bar_rempty + 8 * s, WT_PRODUCERS / 32);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 2 * WT_EPI_WARPS);  // one arrival per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32((const void*)tmem_ptr_s), 512);
  if (warp == 0 && lane == 1) tma_prefetch_desc(&xmap);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // Two independent loader threads in one warp (independent thread scheduling: each spins on its own ring).
    if (lane == 0) {
      // ===================== weight loader (W ring) =====================
      uint32_t g = 0;
      const uint8_t* wp = reinterpret_cast<const uint8_t*>(d.W_packed);
      for (int64_t t = pair; t < n_tiles; t += n_pairs) {
        const WideTmaTile tc = wt_tile(t, n_rt, n_nh);
        const uint8_t* src = wp + (size_t)(tc.nh * 2 + rank) * KB * WT_WBYTES;
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const uint32_t s = g & (WT_WST - 1), ph = (g >> 2) & 1u;
          mbar_wait(bar_wempty + 8 * s, ph ^ 1);
          mbar_arrive_expect_tx(bar_wfull + 8 * s, WT_WBYTES);
          if (WT_DBG(1))  // experiment: no weight traffic
            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(bar_wfull + 8 * s), "r"((uint32_t)WT_WBYTES) : "memory");
          else
            bulk_g2s(base + WT_OFF_W + s * WT_WBYTES, src + (size_t)kb * WT_WBYTES, WT_WBYTES, bar_wfull + 8 * s);
        }
      }
    } else if (lane == 1) {
      // ===================== activation loader (R ring): tensor-map TMA of the raw fp32 tile =====================
      // box = 32 k x 128 frames of item b; frames past the item's end are zero-filled (their columns are never stored)
      uint32_t g = 0;
      for (int64_t t = pair; t < n_tiles; t += n_pairs) {
        const WideTmaTile tc = wt_tile(t, n_rt, n_nh);
        const int row0 = (int)(tc.rt * WT_FRAMES + rank * WT_FR_CTA);
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const uint32_t s = wt_ring3_slot(g), ph = wt_ring3_phase(g);
          mbar_wait(bar_rempty + 8 * s, ph ^ 1);
          mbar_arrive_expect_tx(bar_rfull + 8 * s, WT_RBYTES);
          if (WT_DBG(2))  // experiment: no activation loads
            asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(bar_rfull + 8 * s), "r"((uint32_t)WT_RBYTES) : "memory");
          else
            tma_load_3d(base + WT_OFF_R + s * WT_RBYTES, &xmap, kb * WT_BK, row0, (int)tc.b, bar_rfull + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader) =====================
      uint32_t gs = 0;
      int64_t it = 0;
      const int h1 = KB < WT_HOLD ? KB : WT_HOLD;                 // head: stages [0, h1)
      const int t0 = (KB - WT_HOLD) > h1 ? (KB - WT_HOLD) : h1;  // tail: stages [t0, KB); mid: [h1, t0)
      auto wait_stage = [&](int kb) {
        const uint32_t g = gs + (uint32_t)kb;
        const uint32_t sw = g & (WT_WST - 1), phw = (g >> 2) & 1u, sx = wt_ring3_slot(g), phx = wt_ring3_phase(g);
        mbar_wait(bar_wfull + 8 * sw, phw);  // this CTA's weight stage
        mbar_wait(bar_xfull + 8 * sx, phx);  // this CTA's operand stage (8 producer warps)
        mbar_wait(bar_pfull + 8 * sw, phw);  // the peer's stage (relayed)
        tc_fence_after();
      };
      // MMAs of block mb for stage kb; `release` frees the stage's X and W slots, `done` hands the block's accumulators over
      auto issue = [&](int kb, int mb, bool release, bool done) {
        const uint32_t g = gs + (uint32_t)kb, sw = g & (WT_WST - 1), sx = wt_ring3_slot(g);
        if (elect_one()) {
          const uint32_t xa = base + WT_OFF_X + sx * WT_XBYTES, wa = base + WT_OFF_W + sw * WT_WBYTES;
          const uint64_t x_hi = wt_desc(xa), x_lo = wt_desc(xa + WT_XPART);
          const uint64_t w_hi = wt_desc(wa + mb * WT_WBLK);
          const uint64_t w_lo = wt_desc(wa + (2 + mb) * WT_WBLK);
          const uint32_t dd = tmem_base + (uint32_t)(mb * WT_FRAMES);
#pragma unroll
          for (int k = 0; k < WT_BK / 16; ++k) {
            if (WT_DBG(16)) break;  // experiment: no tensor work
            const uint64_t ko = (uint64_t)((k * 32) >> 4);
            // small cross terms first, the dominant hi*hi last
            umma_bf16_pair(dd, w_lo + ko, x_hi + ko, WT_IDESC, (kb | k) != 0);
            umma_bf16_pair(dd, w_hi + ko, x_lo + ko, WT_IDESC, 1);
            umma_bf16_pair(dd, w_hi + ko, x_hi + ko, WT_IDESC, 1);
          }
          if (release) {
            umma_commit_pair(bar_xempty + 8 * sx);
            umma_commit_pair(bar_wempty + 8 * sw);
          }
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
      // tells the leader that this CTA's half of a stage (its 128 operand frames and its weight blocks) landed
      uint32_t g = 0;
      for (int64_t t = pair; t < n_tiles; t += n_pairs) {
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const uint32_t sw = g & (WT_WST - 1), phw = (g >> 2) & 1u, sx = wt_ring3_slot(g), phx = wt_ring3_phase(g);
          mbar_wait(bar_wfull + 8 * sw, phw);
          mbar_wait(bar_xfull + 8 * sx, phx);
          if (elect_one()) mbar_arrive_remote(bar_pfull + 8 * sw, 0);
          __syncwarp();
        }
      }
    }
  } else if (warp < 2 + WT_EPI_WARPS) {
    // ===================== epilogue (warps 2..9) =====================
    // Block by block (block 0 completes first).  Within a block, warp group g2 takes the 32-frame units with u % 2 == g2, so
    // two warps drain each TMEM lane quarter concurrently.  The residual of the NEXT chunk is loaded into registers while the
    // current one is converted and stored, and the first chunk's before the accumulators are even complete.
    constexpr int G = WT_EPI_WARPS / 4;
    constexpr int NCH = WT_FRAMES / WT_CW;  // chunks per 256-channel block
    const int q = warp & 3;            // TMEM lane quarter this warp may read (hardware rule: warp id % 4)
    const int g2 = (warp - 2) >> 2;
    const int cl = q * 32 + lane;      // channel within this CTA's 128-channel block = TMEM lane
    const int et = tid - 64;
    const float eslope = d.epi_slope ? __ldg(d.epi_slope) : 0.f;
    const int epi_act = d.epi_act;
    const int64_t ystride = d.y_row_stride, rstride = d.res_row_stride;
    const float* const bias = d.bias;
    const float* const bias_batch = d.bias_batch;
    const bool want_stats = d.stats_partials != nullptr;
    const bool has_res = d.residual != nullptr;
    int64_t it = 0;
    // chunk walk of this warp: c = 2*g2, 2*g2+1, 2*g2+4, 2*g2+5, ... (step 1 then 2*G-1)
    auto next_chunk = [&](int c) { return c + ((c & 1) ? (G - 1) * 2 + 1 : 1); };
    for (int64_t t = pair; t < n_tiles; t += n_pairs, ++it) {
      const WideTmaTile tc = wt_tile(t, n_rt, n_nh);
      const uint32_t aph = (uint32_t)(it & 1);
      const int64_t frame0 = (int64_t)tc.rt * WT_FRAMES;
      const int nvalid = (int)((d.rows - frame0) < WT_FRAMES ? (d.rows - frame0) : WT_FRAMES);
      const int64_t ch0 = (int64_t)tc.nh * 512 + rank * 128;  // + mb*256 + cl
      float* yb = d.Y + tc.b * d.y_batch_stride + frame0 * ystride + ch0 + cl;
      const float* rb = has_res ? d.residual + tc.b * d.res_batch_stride + frame0 * rstride + ch0 + cl : nullptr;
      float piv[2], ssum[2], ssq[2], cnt[2];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) { piv[mb] = 0.f; ssum[mb] = 0.f; ssq[mb] = 0.f; cnt[mb] = 0.f; }
      float rn[WT_CW];  // residual of the chunk about to be processed
      auto load_res = [&](int mb, int c) {
        const int nj = nvalid - c * WT_CW;
        const float* p = rb + (int64_t)(c * WT_CW) * rstride + mb * 256;
#pragma unroll
        for (int j = 0; j < WT_CW; ++j) {
          rn[j] = (j < nj) ? __ldg(p) : 0.f;
          p += rstride;
        }
      };
      if (has_res && !WT_DBG(4)) {
        load_res(0, g2 * 2);
      }
#pragma unroll 1
      for (int mb = 0; mb < 2; ++mb) {
        const int64_t ch = ch0 + mb * 256 + cl;
        float bsum = bias ? __ldg(bias + ch) : 0.f;
        if (bias_batch) bsum += __ldg(bias_batch + tc.b * d.M + ch);
        mbar_wait_relaxed(bar_tfull + 8 * mb, aph);
        tc_fence_after();
        float pv = 0.f, sm1 = 0.f, sq1 = 0.f, cn = 0.f;
        // real loops (not unrolled): the chunk body exists once per variant, which keeps the kernel inside the
        // instruction cache
#pragma unroll 1
        for (int c = g2 * 2; c < NCH; c = next_chunk(c)) {
          const int nj = nvalid - c * WT_CW;  // valid frames of this chunk (warp-uniform)
          if (nj <= 0) break;
          float v[WT_CW], r[WT_CW];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * WT_FRAMES + c * WT_CW), v);
          if (WT_DBG(4)) { cn += 1.f; continue; }  // experiment: no epilogue stores / residual loads
          if (has_res) {
#pragma unroll
            for (int j = 0; j < WT_CW; ++j) r[j] = rn[j];
            // prefetch the next chunk this warp will process (next chunk of this block, else the first of the next block)
            int cn2 = next_chunk(c), mb2 = mb;
            if (cn2 >= NCH || nvalid - cn2 * WT_CW <= 0) { cn2 = g2 * 2; mb2 = mb + 1; }
            if (mb2 < 2 && nvalid - cn2 * WT_CW > 0) load_res(mb2, cn2);
          }
          float* yp = yb + (int64_t)(c * WT_CW) * ystride + mb * 256;
          const bool first = cn == 0.f;
          cn += (float)(nj < WT_CW ? nj : WT_CW);
          if (nj >= WT_CW) {
            if (want_stats) {
              if (has_res) wt_epi_chunk<-1, true, true, true>(v, r, bsum, eslope, yp, ystride, WT_CW, first, pv, sm1, sq1, epi_act);
              else if (epi_act == PS_ACT_NONE) wt_epi_chunk<PS_ACT_NONE, false, true, true>(v, r, bsum, eslope, yp, ystride, WT_CW, first, pv, sm1, sq1);
              else wt_epi_chunk<-1, false, true, true>(v, r, bsum, eslope, yp, ystride, WT_CW, first, pv, sm1, sq1, epi_act);
            } else {
              if (has_res && epi_act == PS_ACT_NONE) wt_epi_chunk<PS_ACT_NONE, true, true, false>(v, r, bsum, eslope, yp, ystride, WT_CW, first, pv, sm1, sq1);
              else if (has_res) wt_epi_chunk<-1, true, true, false>(v, r, bsum, eslope, yp, ystride, WT_CW, first, pv, sm1, sq1, epi_act);
              else wt_epi_chunk<-1, false, true, false>(v, r, bsum, eslope, yp, ystride, WT_CW, first, pv, sm1, sq1, epi_act);
            }
          } else {
            if (has_res) wt_epi_chunk<-1, true, false, true>(v, r, bsum, eslope, yp, ystride, nj, first, pv, sm1, sq1, epi_act);
            else wt_epi_chunk<-1, false, false, true>(v, r, bsum, eslope, yp, ystride, nj, first, pv, sm1, sq1, epi_act);
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
          if (lane == 0) wf_s[mb * WT_EPI_WARPS + (warp - 2)] = w;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WT_EPI) : "memory");
        // slot layout of ps_gemm_stats_slots (128-frame rows x 128-channel columns): this tile covers frame rows 2*rt
        // (gets the whole tile's partial) and 2*rt + 1 (gets an empty one)
        const int64_t slots_m = d.M / 128;
        const int64_t n_rt128 = (d.rows + 127) / 128;
        const int64_t slots = n_rt128 * slots_m;
        if (et < 2) {
          const int mb = et;
          Wf tot = wf_s[mb * WT_EPI_WARPS];  // fixed merge order over the epilogue warps -> deterministic
#pragma unroll
          for (int w = 1; w < WT_EPI_WARPS; ++w) tot = wf_merge(tot, wf_s[mb * WT_EPI_WARPS + w]);
          const int64_t col = tc.nh * 4 + mb * 2 + rank;
          float* o = d.stats_partials + (tc.b * slots + (int64_t)(2 * tc.rt) * slots_m + col) * 3;
          o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
          if ((int64_t)(2 * tc.rt + 1) < n_rt128) {
            o += slots_m * 3;
            o[0] = 0.f; o[1] = 0.f; o[2] = 0.f;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WT_EPI) : "memory");
      }
    }
  } else {
    // ===================== operand producers (last 8 warps): R ring -> X ring =====================
    // All 256 threads convert one stage (this CTA's 128 frames x 32 k): thread (r0 = pt / 4, chunk = pt % 4) takes the
    // 8-k chunk of frames r0 and r0 + 64.  Raw tile: 128-byte rows, 16-byte pieces XOR-swizzled by (row & 7)
    // (CU_TENSOR_MAP_SWIZZLE_128B): an 8-lane phase of the 128-bit loads covers two rows x four chunks = eight distinct
    // piece positions, no bank conflict.  Operand tile: 64-byte rows (bf16), wt_swz, same argument for the stores.
    const int pt = tid - (64 + WT_EPI);
    const uint32_t chunk = (uint32_t)(pt & 3);
    const int r0 = pt >> 2;  // 0..63
    const int kofs = (int)chunk * 8;
    const float pslope = (d.pro_act == PS_ACT_PRELU && d.pro_slope) ? __ldg(d.pro_slope) : 1.f;  // AFFINE w/o act = slope 1
    // PReLU with 0 < slope <= 1 is max(u, slope * u) - with the NaN-propagating max, so the reference's Inf/NaN probe
    // (base_nn.py:740-777) still sees what it should; other slopes take the select
    const bool slope01 = pslope > 0.f && pslope <= 1.f;  // (slope 0: Inf * 0 would turn +Inf into NaN)
    uint32_t g = 0;
    for (int64_t t = pair; t < n_tiles; t += n_pairs) {
      const WideTmaTile tc = wt_tile(t, n_rt, n_nh);
      const float* pa = nullptr;
      const float* pb = nullptr;
      if constexpr (PRO == PS_PRO_AFFINE) {
        pa = d.pro_a + tc.b * d.pro_batch_stride + kofs;
        pb = d.pro_b + tc.b * d.pro_batch_stride + kofs;
      }
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const uint32_t s = wt_ring3_slot(g), ph = wt_ring3_phase(g);
        float sc[8], sh[8];
        if constexpr (PRO == PS_PRO_AFFINE) {  // per-item folded norm: 2 KB per item and operand, L1-resident
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(pa + kb * WT_BK)), a1 = __ldg(reinterpret_cast<const float4*>(pa + kb * WT_BK + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(pb + kb * WT_BK)), b1 = __ldg(reinterpret_cast<const float4*>(pb + kb * WT_BK + 4));
          sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
          sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
        }
        mbar_wait(bar_rfull + 8 * s, ph);  // the TMA box landed
        const uint8_t* rt_s = sm + WT_OFF_R + s * WT_RBYTES;
        uint32_t hi[2][4], lo[2][4];
#pragma unroll
        for (int p = 0; p < 2 && !WT_DBG(8); ++p) {
          const uint32_t r = (uint32_t)(p * 64 + r0);
          const float4 v0 = *reinterpret_cast<const float4*>(rt_s + r * 128u + (((2u * chunk) ^ (r & 7u)) << 4));
          const float4 v1 = *reinterpret_cast<const float4*>(rt_s + r * 128u + (((2u * chunk + 1u) ^ (r & 7u)) << 4));
          float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float u0 = v[i], u1 = v[i + 1];
            if constexpr (PRO == PS_PRO_AFFINE) {
              wt_ffma2(u0, u1, sc[i], sc[i + 1], sh[i], sh[i + 1]);
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
            hi[p][i >> 1] = hb;
            lo[p][i >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
          }
        }
        // the raw slot is free as soon as every lane has its values in registers (the conversions above consumed them)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_rempty + 8 * s);
        mbar_wait(bar_xempty + 8 * s, ph ^ 1);  // the MMAs that read this operand slot last have retired
        uint8_t* x_hi = sm + WT_OFF_X + s * WT_XBYTES;
        uint8_t* x_lo = x_hi + WT_XPART;
#pragma unroll
        for (int p = 0; p < 2 && !WT_DBG(8); ++p) {
          const uint32_t off = wt_swz((uint32_t)(p * 64 + r0), chunk);
          *reinterpret_cast<uint4*>(x_hi + off) = make_uint4(hi[p][0], hi[p][1], hi[p][2], hi[p][3]);
          *reinterpret_cast<uint4*>(x_lo + off) = make_uint4(lo[p][0], lo[p][1], lo[p][2], lo[p][3]);
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor cores of both SMs (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_xfull + 8 * s);  // one arrival per producer warp
      }
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
// Selected with PS_TC_WIDE=2 only.  Measured on cfg2 (round 2, runs 7-9): same speed as gemm_wide_kernel in isolation
// (0.36 / 0.37 / 0.49 ms for in_conv / pointwise / out_conv) but SLOWER inside the power-capped step (0.437 vs 0.420 ms per
// launch) - the raw-tile staging adds a shared-memory write + read per operand element, and the step is energy-bound.
bool gemm_wide_tma_eligible(const ps_gemm_t& d, int sms) {
  static EnvInt env;
  if (env.get("PS_TC_WIDE", 1) != 2) return false;
  if (d.M % 512 != 0 || d.K % WT_BK != 0 || d.K < 64 || d.K > WT_MAXK) return false;
  if (!(d.pro_mode == PS_PRO_NONE || (d.pro_mode == PS_PRO_AFFINE && (d.pro_act == PS_ACT_NONE || d.pro_act == PS_ACT_PRELU)))) return false;
  if (d.ln_eps > 0.f || d.fin_scale) return false;
  // tensor-map limits: strides are 16-byte multiples below 2^40 (the 16-byte alignment is gemm_tc_eligible's), dims fit 32 bits
  if (d.rows >= (1LL << 31) || d.batch >= (1LL << 31) || d.x_row_stride * 4 >= (1LL << 40) || d.x_batch_stride * 4 >= (1LL << 40)) return false;
  if (d.batch > 1 && d.x_batch_stride <= 0) return false;
  const int64_t n_tiles = d.batch * cdiv(d.rows, WT_FRAMES) * (d.M / 512);
  return n_tiles >= sms / 2 && n_tiles < (1LL << 31);
}

template <int PRO>
static int launch_wide_tma(const ps_gemm_t& d, const CUtensorMap& xmap, cudaStream_t s, int dev, int64_t grid, int64_t n_rt, int64_t n_nh, int64_t n_tiles) {
  static SmemOnce<1> once;  // per instantiation and device
  if (int rc = once.ensure(dev, 0, gemm_wide_tma_kernel<PRO>, WT_SMEM, "cudaFuncSetAttribute(gemm_wide_tma_kernel)")) return rc;
#ifdef PS_EXPERIMENTS
  static EnvInt dbg_e;
  const int dbg = dbg_e.get("PS_WIDE_TMA_DBG", 0);
#else
  const int dbg = 0;
#endif
  gemm_wide_tma_kernel<PRO><<<(unsigned)grid, WT_THREADS, WT_SMEM, s>>>(d, xmap, n_rt, n_nh, n_tiles, dbg);
  PS_CHECK_LAUNCH("gemm_wide_tma_kernel");
  return PS_OK;
}

int gemm_wide_tma_launch(const ps_gemm_t& d, cudaStream_t s, int dev, int sms) {
  const int64_t n_rt = cdiv(d.rows, WT_FRAMES), n_nh = d.M / 512;
  const int64_t n_tiles = d.batch * n_rt * n_nh;
  const int64_t max_pairs = sms / 2;
  const int64_t grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
  // X as a 3-d fp32 tensor (k, frame, item): box = 32 k x 128 frames x 1 item, 128-byte swizzle, out-of-bounds frames read 0
  CUtensorMap xmap;
  const uint64_t dims[3] = {(uint64_t)d.K, (uint64_t)d.rows, (uint64_t)d.batch};
  const uint64_t strides[2] = {(uint64_t)d.x_row_stride * 4, (uint64_t)(d.batch > 1 ? d.x_batch_stride : d.x_row_stride * d.rows) * 4};
  const uint32_t box[3] = {WT_BK, WT_FR_CTA, 1};
  if (int rc = tma_encode_f32(&xmap, d.X, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (d.pro_mode == PS_PRO_AFFINE) return launch_wide_tma<PS_PRO_AFFINE>(d, xmap, s, dev, grid, n_rt, n_nh, n_tiles);
  return launch_wide_tma<PS_PRO_NONE>(d, xmap, s, dev, grid, n_rt, n_nh, n_tiles);
}

}  // namespace ps
