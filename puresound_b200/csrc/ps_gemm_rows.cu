// tcgen05 / TMEM GEMM for FEW output channels (M <= 128) with the frames on the MMA's M side (sm_100a, one CTA per SM).
//
//   Y[b,f,c] = epi( sum_k pro(X[b,f,k]) * W[c,k] ),  fp32 in HBM, fp32-grade result (3xBF16 split, see ps_gemm_tc.cu).
//
// The CTA-pair kernels (ps_gemm_pair.cu, ps_gemm_wide.cu) put the CHANNELS on the 256-row M side of the MMA: right for
// the 512-channel 1x1 convs, but a 32 ... 128-channel layer (the decoder filterbank, lobe/encoder.py:62-68; Linear ->
// LayerNorm -> + of the DPRNN / SkiM / DPCRN blocks, dprnn.py:161-163,173-175; the U-Net shell, unet.py) is then zero-
// padded to 256 channels, half of the pair idles, and the weight stage is re-streamed from L2 for every 128 frames.
// Measured (round 2, run 18): 0.28-0.44 of the HBM copy rate on shapes that are pure streams of their fp32 operand.
//
// Here the roles are swapped: A = activations (128 frames = the 128 TMEM lanes), B = weights (N = BN = 32 / 64 / 128
// channels), D = [128 frames x BN channels]:
//   * the WHOLE packed weight matrix (BN x K bf16 hi + lo <= 128 KB) is copied into shared memory ONCE per CTA and stays
//     there - no weight stream, no weight hand-shake; a shared-memory stage carries only the operand (16 KB per 32 k);
//   * no padding: the tensor work and the accumulator drain are exactly BN channels wide;
//   * a TMEM lane is a FRAME, so an epilogue thread owns one output row: the row LayerNorm (mean / variance over the M
//     channels) is thread-local - two extra passes over the thread's own TMEM lane, no shuffles, no barriers - and the
//     stores go through a per-warp shared-memory transpose so that a warp instruction writes 4 rows x 128 B.
// Warps: 0 = weight copy (once), 1 = MMA issuer (+ TMEM allocation), 2-5 = epilogue (one per TMEM lane quarter),
// 6-13 = activation producers (fp32 global -> prologue transform -> bf16 hi / lo -> swizzled UMMA tile).
// (Tried and dropped, run 49: one 256-wide tile for 129 ... 256 channels - slower than the CTA-pair kernel with 64-k stages on
// every shape of the TSE / SkiM models: 31808 x 256 x 256 0.039 -> 0.042 ms, 191968 x 256 x 256 0.123 -> 0.149 ms, cfg4 4.86 -> 5.35 ms.)
// Prologues: none, folded norm affine + PReLU, mask product x * act(x2) (mask apply in front of the decoder,
// base_nn.py:41-79).  Epilogues: bias, per-item bias, ReLU / PReLU, residual, Welford partials; or the row LayerNorm.
#include "ps_tc_ptx.cuh"

namespace ps {

constexpr int RW_BM = 128;                       // frames per tile (MMA M = TMEM lanes)
constexpr int RW_BK = 32;                        // k per shared-memory stage (64-byte swizzle rows)
constexpr int RW_STAGES = 4;                     // operand ring depth with the weights resident
constexpr int RW_STAGES_WS = 6;                  // ring depth with the weights streamed (a stage then also carries a weight block)
constexpr int RW_APART = RW_BM * RW_BK * 2;      // 8 KB: operand hi (or lo) of one stage
constexpr int RW_STAGE = 2 * RW_APART;           // 16 KB
constexpr int RW_PRODUCERS = 256, RW_EPI = 128;
constexpr int RW_THREADS = 64 + RW_EPI + RW_PRODUCERS;
constexpr int RW_EPI_PITCH = 36;                 // floats per staged row: 16-byte aligned, conflict-free for 128-bit access
constexpr int RW_EPI_STAGE = 4 * 32 * RW_EPI_PITCH * 4;
constexpr int RW_MAXK = 1024;
constexpr int RW_AFF_BYTES = 2 * RW_MAXK * 4;
constexpr int RW_VEC_BYTES = 3 * 128 * 4;        // bias | ln gamma | ln beta of the (<= 128) channels
constexpr int RW_WMAX = 128 * 1024;              // resident weight image: BN * K * 4 bytes
constexpr int RW_PF_DIST = 4;                    // L2 prefetch distance of the operand stream, in 64-k blocks
constexpr int RW_MISC = RW_EPI_STAGE + RW_AFF_BYTES + RW_VEC_BYTES + 256 /*barriers*/ + 1024 /*align*/;
constexpr int RW_TAIL = RW_STAGES * RW_STAGE + RW_MISC;
constexpr int RW_KMAX_WS = 4096;                 // streamed weights: any K up to here (the affine prologue still needs K <= RW_MAXK)

__host__ __device__ constexpr uint32_t rw_swz(uint32_t r, uint32_t c) { return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4); }

// K-major, SWIZZLE_64B, 8-row atoms 512 B apart
__device__ __forceinline__ uint64_t rw_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

__host__ __device__ constexpr int rw_bn(int64_t M) { return M <= 32 ? 32 : (M <= 64 ? 64 : 128); }
__host__ __device__ constexpr uint32_t rw_tmem_cols(int BN) { return BN <= 32 ? 64u : (BN <= 64 ? 128u : 256u); }

// (128 registers is the ceiling at 14 warps: SM sub-partitions 0 and 1 hold four warps each, 4 x 32 x 128 = their 16 K
// registers; a 144-register build fails to launch.  The affine variant spills 172 B, the mask variant 80 B.)
// kWS: the packed weight matrix does not fit next to the operand ring (BN * K * 4 > 128 KB: the U-Net shell's tap windows of
// 768-1536 floats, out_conv of the 128-channel Conv-TasNet reading) - its 32-k blocks then travel through the ring with the
// operand stages (one cp.async.bulk per stage from warp 0, L2-resident source), six stages deep.
template <int PRO, int BN, bool kLN, bool kWS = false>
__global__ void __launch_bounds__(RW_THREADS, 1) gemm_rows_kernel(const ps_gemm_t d, const uint8_t* __restrict__ wimg, const int64_t n_rt,
                                                                  const int64_t n_tiles) {
  // D = f32, A = B = bf16, both K-major, M = 128, N = BN
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(RW_BM >> 4) << 24);
  constexpr int WBLK = BN * RW_BK * 2;  // weight hi (or lo) of one 32-k block
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int K = (int)d.K, KB = K / RW_BK;
  constexpr int NST = kWS ? RW_STAGES_WS : RW_STAGES;            // ring depth
  constexpr int STAGE = RW_STAGE + (kWS ? 2 * WBLK : 0);         // operand hi | lo (| weight hi | lo)
  const uint32_t wbytes = kWS ? 0u : (uint32_t)KB * 2u * WBLK;   // resident image: a multiple of 4 KB
  const uint32_t ring = base + wbytes;
  uint8_t* ring_p = sm + wbytes;
  float* epi_stage = reinterpret_cast<float*>(ring_p + NST * STAGE);
  float* aff_s = reinterpret_cast<float*>(ring_p + NST * STAGE + RW_EPI_STAGE);
  float* vec_s = reinterpret_cast<float*>(ring_p + NST * STAGE + RW_EPI_STAGE + RW_AFF_BYTES);  // bias | gamma | beta
  const uint32_t bars = ring + NST * STAGE + RW_EPI_STAGE + RW_AFF_BYTES + RW_VEC_BYTES;
  // barrier map (8 B each): full[0..5] empty[0..5] tfull[0..1] tempty[0..1] w, then the TMEM pointer, then statistics scratch
  const uint32_t bar_full = bars, bar_empty = bars + 48, bar_tfull = bars + 96, bar_tempty = bars + 112, bar_w = bars + 128;
  uint8_t* bars_p = ring_p + NST * STAGE + RW_EPI_STAGE + RW_AFF_BYTES + RW_VEC_BYTES;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(bars_p + 144);
  Wf* wf_s = reinterpret_cast<Wf*>(bars_p + 160);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = (int)d.M;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full + 8 * s, 2 * RW_PRODUCERS / 32 + (kWS ? 1 : 0));  // one arrival per producer warp and row half (every warp fills a part of every stage)
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, RW_EPI);
    }
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_ptr_s), rw_tmem_cols(BN));
  if (tid >= 64 && tid < 64 + 128) {  // per-channel vectors of the epilogue (zero beyond M)
    const int c = tid - 64;
    vec_s[c] = (d.bias && c < M) ? __ldg(d.bias + c) : 0.f;
    if constexpr (kLN) {
      vec_s[128 + c] = (d.ln_gamma && c < M) ? __ldg(d.ln_gamma + c) : 1.f;
      vec_s[256 + c] = (d.ln_beta && c < M) ? __ldg(d.ln_beta + c) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if constexpr (!kWS) {
      // ===================== the weight image, once =====================
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_w, wbytes);
        for (uint32_t o = 0; o < wbytes; o += 4096) bulk_g2s(base + o, wimg + o, 4096, bar_w);
      }
      __syncwarp();
    } else {
      // ===================== weight blocks through the ring =====================
      int s = 0;
      uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_full + 8 * s, 2 * WBLK);
            bulk_g2s(ring + s * STAGE + RW_STAGE, wimg + (size_t)kb * (2 * WBLK), 2 * WBLK, bar_full + 8 * s);
          }
          __syncwarp();
          if (++s == NST) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int s = 0;
    uint32_t ph = 0;
    int64_t it = 0;
    if constexpr (!kWS) mbar_wait(bar_w, 0);
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int a = (int)(it & 1);
      const uint32_t aph = (uint32_t)((it >> 1) & 1);
      mbar_wait(bar_tempty + 8 * a, aph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(a * BN);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = ring + s * STAGE, wa = kWS ? sa + RW_STAGE : base + (uint32_t)kb * 2u * WBLK;
          const uint64_t a_hi = rw_desc(sa), a_lo = rw_desc(sa + RW_APART);
          const uint64_t b_hi = rw_desc(wa), b_lo = rw_desc(wa + WBLK);
#pragma unroll
          for (int k = 0; k < RW_BK / 16; ++k) {
            const uint64_t ko = (uint64_t)((k * 32) >> 4);  // +32 B per K = 16 step inside the swizzle row
            // small cross terms first, the dominant hi*hi last
            umma_bf16(tmem_d, a_lo + ko, b_hi + ko, IDESC, (kb | k) != 0);
            umma_bf16(tmem_d, a_hi + ko, b_lo + ko, IDESC, 1);
            umma_bf16(tmem_d, a_hi + ko, b_hi + ko, IDESC, 1);
          }
          umma_commit(bar_empty + 8 * s);  // frees the operand stage when these MMAs retire
          if (kb == KB - 1) umma_commit(bar_tfull + 8 * a);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int et = tid - 64;
    const float eslope = d.epi_slope ? __ldg(d.epi_slope) : 0.f;
    float* stg = epi_stage + (warp - 2) * 32 * RW_EPI_PITCH;
    const int c4 = (lane & 7) * 4;  // this thread's 4 columns inside a 32-column chunk (after the transpose)
    const int rsub = lane >> 3;     // and its row within each group of 4 rows
    const int nch = M / 32;         // 32-column chunks that hold real channels
    int64_t it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int64_t b = t / n_rt, rt = t - b * n_rt;
      const int a = (int)(it & 1);
      const uint32_t aph = (uint32_t)((it >> 1) & 1);
      const int64_t row0 = rt * RW_BM + q * 32;
      float* yb = d.Y + b * d.y_batch_stride + c4;
      const float* rb = d.residual ? d.residual + b * d.res_batch_stride + c4 : nullptr;
      if (rb && row0 + lane < d.rows) {
        // this tile's MMAs are still in flight: pull the residual rows into L2 now
        const char* pr = reinterpret_cast<const char*>(d.residual + b * d.res_batch_stride + (row0 + lane) * d.res_row_stride);
        for (int l = 0; l < M * 4; l += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + l));
      }
      mbar_wait_relaxed(bar_tfull + 8 * a, aph);
      tc_fence_after();
      const float* bbp = d.bias_batch ? d.bias_batch + b * d.M + c4 : nullptr;
      const int nvalid = (int)((d.rows - row0) < 32 ? (d.rows - row0) : 32);  // valid rows of this warp's quarter (may be <= 0)
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
      float mean = 0.f, rstd = 1.f;
      if constexpr (kLN) {
        // row LayerNorm: this thread's TMEM lane IS the row - two passes over it (mean, then the centred second moment)
        float s1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          float v[32];
          tmem_ld32(tacc + (uint32_t)(c * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) s1 += v[j] + vec_s[c * 32 + j];
        }
        mean = s1 / (float)M;
        float s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          float v[32];
          tmem_ld32(tacc + (uint32_t)(c * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float dv = v[j] + vec_s[c * 32 + j] - mean;
            s2 = fmaf(dv, dv, s2);
          }
        }
        rstd = rsqrtf(s2 / (float)M + d.ln_eps);
      }
      float piv = 0.f, ssum = 0.f, ssq = 0.f;
      bool have_piv = false;
#pragma unroll 1
      for (int c = 0; c < nch; ++c) {
        // the eight residual loads of this chunk go out first: their (L2) latency passes behind the TMEM read, the
        // normalisation and the transpose below
        float4 res4[8];
        if (rb) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + rsub;
            res4[i] = rr < nvalid ? __ldg(reinterpret_cast<const float4*>(rb + (row0 + rr) * d.res_row_stride + c * 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float v[32];
        tmem_ld32(tacc + (uint32_t)(c * 32), v);
        if constexpr (kLN) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (v[j] + vec_s[c * 32 + j] - mean) * rstd;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stg + lane * RW_EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        // after the transpose a thread holds 4 consecutive channels of 8 rows
        float4 bs, gm = make_float4(1.f, 1.f, 1.f, 1.f);
        if constexpr (kLN) {
          gm = *reinterpret_cast<const float4*>(vec_s + 128 + c * 32 + c4);
          bs = *reinterpret_cast<const float4*>(vec_s + 256 + c * 32 + c4);
        } else {
          bs = *reinterpret_cast<const float4*>(vec_s + c * 32 + c4);
          if (bbp) {
            const float4 b2 = __ldg(reinterpret_cast<const float4*>(bbp + c * 32));
            bs.x += b2.x; bs.y += b2.y; bs.z += b2.z; bs.w += b2.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + rsub;
          const int64_t row = row0 + rr;
          if (rr < nvalid) {
            const float4 x4 = *reinterpret_cast<const float4*>(stg + rr * RW_EPI_PITCH + c4);
            float o[4];
            if constexpr (kLN) {
              o[0] = fmaf(x4.x, gm.x, bs.x); o[1] = fmaf(x4.y, gm.y, bs.y); o[2] = fmaf(x4.z, gm.z, bs.z); o[3] = fmaf(x4.w, gm.w, bs.w);
            } else {
              o[0] = x4.x + bs.x; o[1] = x4.y + bs.y; o[2] = x4.z + bs.z; o[3] = x4.w + bs.w;
              if (d.epi_act == PS_ACT_RELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = (o[e] != o[e]) ? o[e] : fmaxf(o[e], 0.f);
              } else if (d.epi_act == PS_ACT_PRELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = o[e] > 0.f ? o[e] : o[e] * eslope;
              }
            }
            if (rb) { o[0] += res4[i].x; o[1] += res4[i].y; o[2] += res4[i].z; o[3] += res4[i].w; }
            *reinterpret_cast<float4*>(yb + row * d.y_row_stride + c * 32) = make_float4(o[0], o[1], o[2], o[3]);
            if constexpr (!kLN) {
              if (!have_piv) { piv = o[0]; have_piv = true; }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float dv = o[e] - piv;
                ssum += dv;
                ssq = fmaf(dv, dv, ssq);
              }
            }
          }
        }
        __syncwarp();
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld inside tmem_ld32): hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * a);
      if constexpr (!kLN) {
        if (d.stats_partials) {
          // rows rsub, rsub+4, ... below nvalid, 4 columns x nch chunks each
          const int nrows_t = nvalid > rsub ? (nvalid - rsub + 3) / 4 : 0;
          Wf mine;
          mine.n = (float)(nrows_t * 4 * nch);
          mine.mean = 0.f; mine.m2 = 0.f;
          if (mine.n > 0.f) {
            const float md = ssum / mine.n;
            mine.mean = piv + md;
            mine.m2 = fmaxf(ssq - ssum * md, 0.f);
          }
          Wf w = wf_warp_reduce(mine);
          if (lane == 0) wf_s[q] = w;
          asm volatile("bar.sync 1, %0;" ::"n"(RW_EPI) : "memory");
          if (et == 0) {
            // fixed merge order over the four row quarters -> deterministic; one 128-row x (<= 128)-channel slot per tile
            Wf tot = wf_merge(wf_merge(wf_s[0], wf_s[1]), wf_merge(wf_s[2], wf_s[3]));
            float* o = d.stats_partials + (b * n_rt + rt) * 3;
            o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(RW_EPI) : "memory");
        }
      }
    }
  } else {
    // ===================== activation producers (warps 6..13) =====================
    // Work unit = 64 rows x 64 k of a tile: 8 threads cover the 64 k of a row (coalesced 256 B), 32 rows per pass, 2 passes;
    // threads with kc < 4 write stage s, the others stage s + 1 (a 64-k block = two consecutive 32-k stages, its two units
    // fill rows 0-63 and 64-127 of both).  The global loads of a unit (4 x LDG.128 per thread) are issued THREE units ahead
    // of their use (four register buffers): with one 128-row block in flight behind the one being transformed, an
    // iteration took one HBM latency per 32 KB - 26 GB/s per SM, 0.44 of the HBM rate at the DPRNN projection (run 21).
    const int pt = tid - 192;
    const int kc = pt & 7;              // k = kc*8 .. kc*8+7 of the 64-k block
    const int r0 = pt >> 3;             // 0..31
    const int sub = kc >> 2;            // which of the block's two stages this thread fills
    const uint32_t cch = (uint32_t)(kc & 3);  // 16-byte chunk inside that stage's row
    const int KB64 = K / 64;
    const float pslope = (PRO == PS_PRO_AFFINE && d.pro_act == PS_ACT_PRELU && d.pro_slope) ? __ldg(d.pro_slope) : 1.f;  // no act = slope 1
    const int mask_act = d.pro_act;
    const int64_t x2_delta = (PRO == PS_PRO_MASK) ? (d.X2 - d.X) : 0;  // the mask has the strides of X
    int s = 0;
    uint32_t ph = 0;

    // position in this CTA's sequence of units, kept in 32-bit (tile and row counts are < 2^31, checked at launch): three
    // of these walk ahead of each other (transform / load / prefetch) and every register counts at 128 per thread
    struct Cur {
      int t;         // tile
      int u;         // unit inside the tile: 64-k block u >> 1, row half u & 1
      int row_base;  // rt*128 + r0
      int b;         // batch item
    };
    const int n_tiles_i = (int)n_tiles, last_row = (int)d.rows - 1, n_rt_i = (int)n_rt;
    const int units = 2 * KB64, tstep = (int)gridDim.x;
    auto decode = [&](Cur& c) {
      if (c.t < n_tiles_i) {
        c.b = c.t / n_rt_i;
        c.row_base = (c.t - c.b * n_rt_i) * RW_BM + r0;
      }
    };
    auto advance = [&](Cur& c) {  // next unit
      if (++c.u == units) {
        c.u = 0;
        c.t = c.t + tstep < c.t ? n_tiles_i : c.t + tstep;  // (no wrap-around)
        decode(c);
      }
    };
    // &X[b][row][kc*8 + 64 * block] of pass p of a unit; rows past the end of the item re-read its last row: their
    // accumulator rows are computed but never stored
    auto src_of = [&](const Cur& c, int p) -> const float* {
      int row = c.row_base + (c.u & 1) * 64 + p * 32;
      row = row < last_row ? row : last_row;
      return d.X + (int64_t)c.b * d.x_batch_stride + (int64_t)row * d.x_row_stride + (kc * 8 + (c.u >> 1) * 64);
    };
    auto issue = [&](float4(&x)[2][2], const Cur& c) {
      if (c.t >= n_tiles_i) return;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const float4* src = reinterpret_cast<const float4*>(src_of(c, p));
        x[p][0] = __ldg(src);
        x[p][1] = __ldg(src + 1);
      }
    };
    // software prefetch into L2, RW_PF_DIST 64-k blocks ahead of the register loads: one 128-byte line per 4 threads
    auto prefetch = [&](const Cur& c) {
      if ((kc & 3) == 0 && c.t < n_tiles_i) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const float* src = src_of(c, p);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
          if constexpr (PRO == PS_PRO_MASK) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + x2_delta));
        }
      }
    };
    int staged_b = -1;
    auto stage_affine = [&](int b) {
      if constexpr (PRO == PS_PRO_AFFINE) {
        if (b != staged_b) {
          asm volatile("bar.sync 2, %0;" ::"n"(RW_PRODUCERS) : "memory");  // every producer is done reading the old rows
          const float* pa = d.pro_a + (int64_t)b * d.pro_batch_stride;
          const float* pb = d.pro_b + (int64_t)b * d.pro_batch_stride;
          for (int k = pt * 4; k < K; k += RW_PRODUCERS * 4) {
            *reinterpret_cast<float4*>(aff_s + k) = __ldg(reinterpret_cast<const float4*>(pa + k));
            *reinterpret_cast<float4*>(aff_s + K + k) = __ldg(reinterpret_cast<const float4*>(pb + k));
          }
          asm volatile("bar.sync 2, %0;" ::"n"(RW_PRODUCERS) : "memory");
          staged_b = b;
        }
      }
    };
    // transform + bf16 hi/lo split + swizzled store of one unit into stages s (kc < 4) and s + 1
    // the mask operand of a unit (MASK prologue; its lines were prefetched into L2): two register buffers that alternate,
    // each refilled for the unit two steps ahead as soon as process() has consumed it - the operand buffers stay four deep
    auto load_mask = [&](float4(&mk)[2][2], const Cur& c) {
      if constexpr (PRO == PS_PRO_MASK) {
        if (c.t >= n_tiles_i) return;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const float4* src = reinterpret_cast<const float4*>(src_of(c, p) + x2_delta);
          mk[p][0] = __ldg(src);
          mk[p][1] = __ldg(src + 1);
        }
      }
    };
    auto process = [&](const float4(&x)[2][2], float4(&mk)[2][2], const Cur& c, const Cur& c2) {
      stage_affine(c.b);
      float sc[8], sh[8];
      if constexpr (PRO == PS_PRO_AFFINE) {
        const int k0 = (c.u >> 1) * 64 + kc * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(aff_s + k0), a1 = *reinterpret_cast<const float4*>(aff_s + k0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(aff_s + K + k0), b1 = *reinterpret_cast<const float4*>(aff_s + K + k0 + 4);
        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
      }
      const int my_s = s + sub;
      uint8_t* a_hi = ring_p + my_s * STAGE;
      uint8_t* a_lo = a_hi + RW_APART;
      auto split_store = [&](const float(&v)[8], int p) {
        const int r = (c.u & 1) * 64 + p * 32 + r0;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float u0 = v[i], u1 = v[i + 1];
          if constexpr (PRO == PS_PRO_AFFINE) {
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
        const uint32_t off = rw_swz((uint32_t)r, cch);
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      };
      if constexpr (PRO == PS_PRO_MASK) {
        // the products first, so that the mask registers are dead and can be refilled at once for the unit TWO steps ahead
        float v[2][8];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const float x8[8] = {x[p][0].x, x[p][0].y, x[p][0].z, x[p][0].w, x[p][1].x, x[p][1].y, x[p][1].z, x[p][1].w};
          const float m8[8] = {mk[p][0].x, mk[p][0].y, mk[p][0].z, mk[p][0].w, mk[p][1].x, mk[p][1].y, mk[p][1].z, mk[p][1].w};
#pragma unroll
          for (int i = 0; i < 8; ++i) v[p][i] = x8[i] * apply_act(m8[i], mask_act, 0.f);
        }
        load_mask(mk, c2);
        mbar_wait(bar_empty + 8 * my_s, ph ^ 1);
        split_store(v[0], 0);
        split_store(v[1], 1);
      } else {
        mbar_wait(bar_empty + 8 * my_s, ph ^ 1);  // (passes at once for the second unit of a block)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const float x8[8] = {x[p][0].x, x[p][0].y, x[p][0].z, x[p][0].w, x[p][1].x, x[p][1].y, x[p][1].z, x[p][1].w};
          split_store(x8, p);
        }
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {  // a warp holds kc = 0..7, i.e. a part of BOTH stages of the block
        mbar_arrive(bar_full + 8 * s);
        mbar_arrive(bar_full + 8 * (s + 1));
      }
      if (c.u & 1) {
        s += 2;
        if (s == NST) { s = 0; ph ^= 1; }
      }
    };

    float4 x0[2][2], x1[2][2], x2[2][2], x3[2][2], mA[2][2], mB[2][2];
    Cur cur, ld, pf;  // unit being transformed, unit being loaded, unit being prefetched
    cur.t = (int)blockIdx.x; cur.u = 0; cur.row_base = 0; cur.b = 0;
    decode(cur);
    pf = cur;
    for (int i = 0; i < 2 * RW_PF_DIST && pf.t < n_tiles_i; ++i) {
      prefetch(pf);
      advance(pf);
    }
    ld = cur;
    issue(x0, ld); advance(ld);
    issue(x1, ld); advance(ld);
    if constexpr (PRO != PS_PRO_MASK) { issue(x2, ld); advance(ld); }
    {
      Cur n1 = cur;
      advance(n1);
      load_mask(mA, cur);
      load_mask(mB, n1);
    }
    // each step: request the unit three ahead (two with the MASK prologue, whose two mask buffers need the registers: a
    // four-deep build spilled 192 B per thread and ran 0.44 instead of 0.29 ms at the decoder), prefetch further ahead,
    // transform the oldest buffer (which re-requests its mask buffer for the unit two ahead)
#define RW_STEP(XNEW, XOLD, MBUF)                         \
  {                                                       \
    issue(XNEW, ld); advance(ld);                         \
    prefetch(pf); if (pf.t < n_tiles_i) advance(pf);      \
    Cur n1 = cur; advance(n1);                            \
    Cur n2 = n1; advance(n2);                             \
    process(XOLD, MBUF, cur, n2); cur = n1;               \
    if (cur.t >= n_tiles_i) break;                        \
  }
    if constexpr (PRO == PS_PRO_MASK) {
      while (cur.t < n_tiles_i) {  // 3 operand buffers x 2 mask buffers: period 6
        RW_STEP(x2, x0, mA)
        RW_STEP(x0, x1, mB)
        RW_STEP(x1, x2, mA)
        RW_STEP(x2, x0, mB)
        RW_STEP(x0, x1, mA)
        RW_STEP(x1, x2, mB)
      }
    } else {
      while (cur.t < n_tiles_i) {
        RW_STEP(x3, x0, mA)
        RW_STEP(x0, x1, mB)
        RW_STEP(x1, x2, mA)
        RW_STEP(x2, x3, mB)
      }
    }
#undef RW_STEP
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, rw_tmem_cols(BN));
  }
}

// ---------------------------------------------------------------- weight packing
// W [M, K] fp32 -> per 32-k block: [hi BN x 32 bf16 | lo BN x 32 bf16], each in the K-major 64-byte-swizzled shared-memory
// image (rows beyond M are zero): BN * K * 4 bytes, copied linearly into shared memory by the kernel.
__global__ void pack_weights_rows_kernel(const float* __restrict__ W, int64_t ldw, int64_t M, int64_t K, int BN, uint8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)BN * K) return;
  const int64_t n = i / K, k = i % K;
  const int64_t kb = k / RW_BK, kk = k % RW_BK;
  const float x = n < M ? W[n * ldw + k] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  const size_t blk = (size_t)kb * 2 * BN * RW_BK * 2;
  const size_t off = (size_t)rw_swz((uint32_t)n, (uint32_t)(kk >> 3)) + (size_t)(kk & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(out + blk + off) = h;
  *reinterpret_cast<__nv_bfloat16*>(out + blk + (size_t)BN * RW_BK * 2 + off) = l;
}

// bytes of the resident image for this shape, 0 when the kernel does not serve it (PS_GEMM_ROWS=0 switches it off: A/B)
int64_t gemm_rows_image_bytes(int64_t M, int64_t K) {
  static EnvInt env;
  if (env.get("PS_GEMM_ROWS", 1) == 0) return 0;
  if (M < 32 || M > 128 || M % 32 != 0 || K < 64 || K % 64 != 0 || K > RW_KMAX_WS) return 0;
  const int64_t bytes = (int64_t)rw_bn(M) * K * 4;  // resident when <= 128 KB, else streamed through the operand ring
  static EnvInt ws_env;
  if (bytes > RW_WMAX && ws_env.get("PS_GEMM_ROWS_WS", 1) == 0) return 0;  // A/B: streamed-weight shapes back on the CTA-pair kernel
  return bytes;
}

int gemm_rows_pack(const float* W, int64_t ldw, int64_t M, int64_t K, void* packed, cudaStream_t s) {
  const int BN = rw_bn(M);
  pack_weights_rows_kernel<<<(unsigned)cdiv((int64_t)BN * K, 256), 256, 0, s>>>(W, ldw, M, K, BN, reinterpret_cast<uint8_t*>(packed));
  PS_CHECK_LAUNCH("pack_weights_rows_kernel");
  return PS_OK;
}

static inline bool rw_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// d has passed gemm_tc_eligible (alignment of X / Y / bias / residual / prologue vectors, activation kinds)
bool gemm_rows_eligible(const ps_gemm_t& d) {
  if (gemm_rows_image_bytes(d.M, d.K) == 0) return false;
  if (d.rows >= (1LL << 30) || d.batch * cdiv(d.rows, RW_BM) >= (1LL << 30)) return false;  // 32-bit cursors in the producers
  if (d.pro_mode == PS_PRO_AFFINE && d.K > RW_MAXK) return false;  // scale | shift rows of an item are staged in shared memory
  if (!(d.pro_mode == PS_PRO_NONE || d.pro_mode == PS_PRO_MASK || (d.pro_mode == PS_PRO_AFFINE && (d.pro_act == PS_ACT_NONE || d.pro_act == PS_ACT_PRELU))))
    return false;
  if (d.ln_eps > 0.f && (d.pro_mode != PS_PRO_NONE || d.epi_act != PS_ACT_NONE || d.stats_partials || d.bias_batch)) return false;
  if ((d.ln_gamma && !rw_al16(d.ln_gamma)) || (d.ln_beta && !rw_al16(d.ln_beta))) return false;
  return true;
}

template <int PRO, int BN, bool kLN>
static int launch_rows(const ps_gemm_t& d, const uint8_t* wimg, cudaStream_t s, int dev, int64_t grid, int64_t n_rt, int64_t n_tiles) {
  static SmemOnce<2> once;  // per instantiation and device: raised to the largest image once
  if ((int64_t)BN * d.K * 4 <= RW_WMAX) {
    if (int rc = once.ensure(dev, 0, gemm_rows_kernel<PRO, BN, kLN, false>, RW_WMAX + RW_TAIL, "cudaFuncSetAttribute(gemm_rows_kernel)")) return rc;
    const int smem = (int)(BN * d.K * 4) + RW_TAIL;
    gemm_rows_kernel<PRO, BN, kLN, false><<<(unsigned)grid, RW_THREADS, smem, s>>>(d, wimg, n_rt, n_tiles);
  } else {
    constexpr int smem = RW_STAGES_WS * (RW_STAGE + 2 * BN * RW_BK * 2) + RW_MISC;
    if (int rc = once.ensure(dev, 1, gemm_rows_kernel<PRO, BN, kLN, true>, smem, "cudaFuncSetAttribute(gemm_rows_kernel)")) return rc;
    gemm_rows_kernel<PRO, BN, kLN, true><<<(unsigned)grid, RW_THREADS, smem, s>>>(d, wimg, n_rt, n_tiles);
  }
  PS_CHECK_LAUNCH("gemm_rows_kernel");
  return PS_OK;
}

template <int BN>
static int launch_rows_bn(const ps_gemm_t& d, const uint8_t* wimg, cudaStream_t s, int dev, int64_t grid, int64_t n_rt, int64_t n_tiles) {
  if (d.ln_eps > 0.f) return launch_rows<PS_PRO_NONE, BN, true>(d, wimg, s, dev, grid, n_rt, n_tiles);
  if (d.pro_mode == PS_PRO_AFFINE) return launch_rows<PS_PRO_AFFINE, BN, false>(d, wimg, s, dev, grid, n_rt, n_tiles);
  if (d.pro_mode == PS_PRO_MASK) return launch_rows<PS_PRO_MASK, BN, false>(d, wimg, s, dev, grid, n_rt, n_tiles);
  return launch_rows<PS_PRO_NONE, BN, false>(d, wimg, s, dev, grid, n_rt, n_tiles);
}

// wimg = the rows image inside d.W_packed (it follows the CTA-pair image, see ps_gemm_pack_weights)
int gemm_rows_launch(const ps_gemm_t& d, const void* wimg, cudaStream_t s) {
  int dev = 0, sms = 0;
  if (int rc = current_device(&dev)) return rc;
  if (int rc = sm_count_of(dev, &sms)) return rc;
  const int64_t n_rt = cdiv(d.rows, RW_BM);
  const int64_t n_tiles = d.batch * n_rt;
  if (n_tiles >= (1LL << 30) || d.rows >= (1LL << 30)) return PS_ERR_UNSUPPORTED;  // 32-bit cursors in the producers
  const int64_t grid = n_tiles < sms ? n_tiles : sms;
  const uint8_t* w = reinterpret_cast<const uint8_t*>(wimg);
  const int BN = rw_bn(d.M);
  if (BN == 32) return launch_rows_bn<32>(d, w, s, dev, grid, n_rt, n_tiles);
  if (BN == 64) return launch_rows_bn<64>(d, w, s, dev, grid, n_rt, n_tiles);
  return launch_rows_bn<128>(d, w, s, dev, grid, n_rt, n_tiles);
}

}  // namespace ps
