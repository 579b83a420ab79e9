// Tensor-core LSTM recurrence for H <= 128, H % 32 == 0 (DPRNN intra- / inter-chunk passes; reference dprnn.py:67-103 via
// nn.LSTM, gate order i,f,g,o).  The text below describes H = 128; smaller H run in the same layout with the missing
// units / k-columns of W_hh zero-padded by the pack kernel: their lanes stay h = c = 0, the k loop stops at H, the gate
// warps of all-padding lane quarters only keep the barriers in step.  One CTA per SM owns up to 64 sequences of one
// direction for all L steps (fewer per CTA when the whole problem fits one wave of CTAs anyway: a shorter gate phase
// is a shorter step, and the recurrence is a latency chain of L steps); the recurrent matrix never leaves the SM:
//
//   * W_hh (512 x 128 fp32 = 256 KB) does not fit the 227 KB of shared memory, so it is kept as a bf16 hi/lo split
//     (the 3xBF16 scheme of ps_gemm_tc.cu: hi*hi + hi*lo + lo*hi in the fp32 accumulator, ~2^-17 per product) with
//     W_lo (128 KB) resident in SHARED MEMORY and W_hi (128 KB) resident in TENSOR MEMORY, where tcgen05.mma takes it
//     as its A operand (TS form; two of the three passes use W_hi, and a TMEM A operand is cheaper than a smem tile);
//   * per step the gate pre-activations  G[512 x 64] = W_hh[512 x 128] * h_{t-1}^T[128 x 64]  are 96 tcgen05.mma
//     (M=128 = one gate of all 128 units, N=64 sequences, K=16; 4 gates x 8 k-steps x 3 split passes) into the other
//     256 TMEM columns; B = h_{t-1} as bf16 hi/lo, K-major 64-byte-swizzled tiles in shared memory;
//   * TMEM lane = hidden unit, column = sequence, one 64-column block per gate: a thread reads i,f,g,o of ITS unit for
//     a sequence from the four blocks, so the cell update is thread-local (c in registers), and writes h_t back as
//     the next step's B operand (bf16 hi/lo) and to the output tensor (a warp = 32 consecutive units = one coalesced
//     128-byte store per sequence);
//   * gx = W_ih x + b (ps_gemm) is read in its native [position, D*4H] layout: for one gate and sequence a warp reads
//     128 contiguous bytes; next step's lines are prefetched into L2 while this step computes.
//
// Warps: 0 = MMA issuer (+ TMEM allocation), 4..19 = gate warps (four per TMEM lane quarter; each owns 8 sequences in
// EACH half-batch).
// Barriers, one pair per half-batch of 32 sequences: mma_done (tcgen05.commit -> gate warps), h_ready (gate warps ->
// MMA issuer); all gate warps work through half A, then half B: the tensor core computes one half's gates while the
// cells of the other are updated.
// The strided position function of ps_lstm_t is honoured, so one [N,S,K,*] tensor serves both passes without a permute.
#include <stdlib.h>

#include "ps_tc_ptx.cuh"

#ifdef PS_EXPERIMENTS
#define LT_DBG(bit) ((dbg & (bit)) != 0)
#else
#define LT_DBG(bit) false
#endif

namespace ps {

constexpr int LT_H = 128, LT_N = 64;           // hidden units, sequences per CTA
constexpr int LT_GWQ = 2;                      // gate warps per (half-batch, TMEM lane quarter)
constexpr int LT_THREADS = (4 + 8 * LT_GWQ) * 32;  // warp 0 MMA, 1-3 idle (registers are granted per 4 warps), then the gate warps
constexpr int LT_WTILE = 128 * 64;             // one [128 rows x 32 k] bf16 tile, 64-byte swizzle: 8 KB
constexpr int LT_WHI_BYTES = 4 * 4 * LT_WTILE; // 4 gates x 4 k-tiles = 128 KB
constexpr int LT_HTILE = LT_N * 64;            // [64 seqs x 32 k] bf16: 4 KB
constexpr int LT_H_BYTES = 2 * 4 * LT_HTILE;   // hi | lo, 4 k-tiles each = 32 KB
constexpr int LT_POS_BYTES = 2 * LT_N * 8;     // per sequence slot: its first row (position) as int32 (the rest is spare)
constexpr int LT_C_BYTES = LT_N * LT_H * 4;     // cell state [sequence][unit] fp32: 32 KB (registers go to the gx double buffer)
constexpr int LT_SMEM = LT_WHI_BYTES + LT_H_BYTES + LT_POS_BYTES + 64 /*barriers*/ + LT_C_BYTES + 1024 /*align*/;
constexpr uint32_t LT_ACC_COL = 256;           // TMEM: W_lo in columns [0,256), accumulators in [256,512)
// D=f32, A=B=bf16, K-major, M=128, N=32 (one half-batch of sequences per MMA)
constexpr int LT_NH = LT_N / 2;
constexpr uint32_t LT_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LT_NH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__host__ __device__ constexpr uint32_t lt_swz(uint32_t r, uint32_t c) { return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4); }

__device__ __forceinline__ uint64_t lt_desc(uint32_t saddr) {  // K-major, SWIZZLE_64B, 8-row atoms 512 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One MUFU.EX2 + one MUFU.RCP per activation (2^-22 / 1 ulp), ~1e-7 absolute on the gate; saturates cleanly
// (ex2(+big) = inf -> rcp(inf) = 0) and propagates NaN.  The IEEE-rounded __frcp_rn / exp2f forms cost ~25 instructions
// per activation and made the gate phase 90 % of the step (ncu, profiles/).
__device__ __forceinline__ float lt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lt_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lt_sigmoid(float x) { return lt_rcp(1.f + lt_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float lt_tanh(float x) { return fmaf(2.f, lt_rcp(1.f + lt_ex2(-2.8853900817779268f * x)), -1.f); }
// min that PROPAGATES NaN (fminf would swallow it; the reference's look-ahead probe feeds inf/NaN)
__device__ __forceinline__ float lt_min_nan(float x, float lim) {
  float y;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(lim));
  return y;
}
// One LSTM cell with 7 MUFU instead of 10: with Ex = exp(-x) the three sigmoids and two tanh share reciprocals,
//   c' = f c + i g = [c (1+Ei)(1+Eg) + (1-Eg)(1+Ef)] / [(1+Ef)(1+Ei)(1+Eg)],   Eg = exp(-2 g)
//   h  = o tanh(c') = (1 - Ec) / [(1+Eo)(1+Ec)],                                  Ec = exp(-2 c')
// The exponents are capped at 2^36 = exp(25) (a pre-activation below -25, or -12.5 under tanh): exp(25)^3 = 3.7e32 stays
// finite in fp32, and the cap moves a gate by < 1.4e-11; towards -inf ex2 underflows to 0 on its own.  One FMNMX per
// exponential.  The XU pipe (16 lanes / cycle / SM) is the floor of the gate phase.
// The arguments are the exponents themselves: xi = -log2(e) * pre-activation (xg = -2 log2(e) * ...): the pack kernel
// folds these factors into W_hh, and gx joins with one FFMA (LT_SCALE).
__device__ __forceinline__ float lt_cell(float xi, float xf, float xg, float xo, float& c) {
  constexpr float L2E = 1.4426950408889634f, CAP = 36.0673760222f;
  const float Ei = lt_ex2(lt_min_nan(xi, CAP)), Ef = lt_ex2(lt_min_nan(xf, CAP));
  const float Eg = lt_ex2(lt_min_nan(xg, CAP)), Eo = lt_ex2(lt_min_nan(xo, CAP));
  const float A = 1.f + Ei, B = 1.f + Eg, F = 1.f + Ef;
  const float AB = A * B;
  const float cn = fmaf(c, AB, (2.f - B) * F) * lt_rcp(F * AB);
  c = cn;
  const float Cc = 1.f + lt_ex2(lt_min_nan(-2.f * L2E * cn, CAP));
  return (2.f - Cc) * lt_rcp((1.f + Eo) * Cc);
}

// spq = real sequences per (gate warp, half-batch) (1..8): sequence slot s = half*32 + wq*8 + j is real iff j < spq, and a
// CTA owns 8*spq consecutive sequences.  kCh2: spq > 4, i.e. two chunks of four sequences per warp and half.
// exponent scale of gate g (i, f, g, o): sigmoid(p) = 1 / (1 + 2^(-log2(e) p)), tanh(p) = 2 / (1 + 2^(-2 log2(e) p)) - 1
__host__ __device__ constexpr float lt_scale(int g) { return g == 2 ? -2.8853900817779268f : -1.4426950408889634f; }

template <bool kGxi, bool kCh2>
__global__ void __launch_bounds__(LT_THREADS, 1) lstm_tc_kernel(const ps_lstm_t d, const int spq, const int dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint8_t* h_sm = sm + LT_WHI_BYTES;                       // [hi: 4 tiles][lo: 4 tiles]
  int32_t* row_s = reinterpret_cast<int32_t*>(sm + LT_WHI_BYTES + LT_H_BYTES);  // first row (position) of every sequence slot
  const uint32_t bars = base + LT_WHI_BYTES + LT_H_BYTES + LT_POS_BYTES;
  const uint32_t bar_w = bars, bar_mma = bars + 8 /*[2]*/, bar_h = bars + 24 /*[2]*/;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + LT_WHI_BYTES + LT_H_BYTES + LT_POS_BYTES + 48);
  float* c_sm = reinterpret_cast<float*>(sm + LT_WHI_BYTES + LT_H_BYTES + LT_POS_BYTES + 64);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const int Hr = (int)d.H;  // real hidden size; the on-chip layout is always LT_H wide
  const int64_t q0 = (int64_t)blockIdx.x * (8 * spq);
  const int G = d.D * 4 * Hr, OW = d.D * Hr;  // row widths of gx and out
  const uint8_t* wimg = reinterpret_cast<const uint8_t*>(d.w_packed) + (size_t)dir * (2 * LT_WHI_BYTES);

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int hf = 0; hf < 2; ++hf) {
      mbar_init(bar_mma + 8 * hf, 1);
      mbar_init(bar_h + 8 * hf, 8 * LT_GWQ);  // every gate warp serves both half-batches
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32((const void*)tmem_ptr_s), 512);
  if (tid < LT_N) {
    int64_t q = q0 + (int64_t)(tid >> 3) * spq + (tid & 7);
    if ((tid & 7) >= spq || q >= d.n_seq) q = d.n_seq - 1;  // unused slots shadow the last real sequence; never stored
    row_s[tid] = (int32_t)((q / d.inner) * d.outer_stride + (q % d.inner) * d.inner_stride);  // < 2^31: checked by the launcher
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== W_hi -> shared memory, then the MMA issue loop =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w, LT_WHI_BYTES);
#pragma unroll
      for (int i = 0; i < 4; ++i) bulk_g2s(base + i * (LT_WHI_BYTES / 4), wimg + (size_t)i * (LT_WHI_BYTES / 4), LT_WHI_BYTES / 4, bar_w);
    }
    __syncwarp();
    mbar_wait(bar_w, 0);
    // The 64 sequences run as two independent half-batches of 32 (own barriers, own accumulator columns): while the
    // gate warps of one half compute its cell update, the tensor core is already busy with the other half's step.
    const int nk = Hr >> 4;  // k16 steps over the real K = H
    for (int64_t step = 0; step < d.L; ++step) {
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        // h_{t-1} of this half (and, for step 0, h0 and W_lo in TMEM) is in place; its accumulators have been read
        mbar_wait(bar_h + 8 * hf, (uint32_t)(step & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t hs = base + LT_WHI_BYTES + (uint32_t)(hf * LT_NH * 64);  // rows [32*hf, 32*hf+32) of every h tile
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (LT_DBG(1)) break;  // experiment: gate phase alone (PS_LSTM_DBG=1, results are garbage)
            const uint32_t dd = tmem_base + LT_ACC_COL + (uint32_t)(g * LT_N + hf * LT_NH);
#pragma unroll
            for (int k = 0; k < 8; ++k) {  // k16 steps over K = 128
              if (k >= nk) break;
              const int kt = k >> 1;
              const uint64_t ko = (uint64_t)(((k & 1) * 32) >> 4);
              const uint64_t h_hi = lt_desc(hs + kt * LT_HTILE) + ko, h_lo = lt_desc(hs + (4 + kt) * LT_HTILE) + ko;
              const uint64_t w_lo = lt_desc(base + (g * 4 + kt) * LT_WTILE) + ko;
              const uint32_t w_hi = tmem_base + (uint32_t)(g * 64 + k * 8);  // 8 columns = 16 packed bf16
              // two of the three passes take A from tensor memory: a shared-memory A tile costs >= 32 cycles per MMA
              // (4 KB at 128 B/cycle) however small N is
              umma_bf16(dd, w_lo, h_hi, LT_IDESC, k != 0);
              umma_bf16_ts(dd, w_hi, h_lo, LT_IDESC, 1);
              umma_bf16_ts(dd, w_hi, h_hi, LT_IDESC, 1);
            }
          }
          umma_commit(bar_mma + 8 * hf);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== gate warps =====================
    // warp = 4 + wq*4 + q: q = TMEM lane quarter (hardware rule: warp id % 4), wq = which SPH-sequence slice it owns in
    // EACH half-batch.  Every gate warp serves both halves in turn: while it updates the cells of half A the tensor core
    // computes half B's gates and vice versa, so with G = gate time of all 64 sequences and M = MMA time of one half the
    // step costs max(G, G/2 + M, 2M) instead of the G/2' + M of warps bound to one half (which idle during their M).
    constexpr int NW = LT_GWQ * 2, SPH = LT_NH / NW, CH = 4;  // 4 warps per lane quarter, 8 slots per (warp, half)
    const int q = warp & 3;
    const int wq = (warp - 4) >> 2;
    const int u = q * 32 + lane;         // hidden unit = TMEM lane
    const bool live = q * 32 < Hr;       // false: every unit of this lane quarter is padding
    // ---- one-time: W_hi rows of this lane into TMEM columns [g*64, g*64+64) (each 32-bit column = 2 consecutive k)
    if (wq == 0) {
      const uint32_t* wlo = reinterpret_cast<const uint32_t*>(wimg + LT_WHI_BYTES);  // [512 rows][64 words]
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[16];
          const uint4* src = reinterpret_cast<const uint4*>(wlo + (size_t)(g * LT_H + u) * 64 + c * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 t4 = __ldg(src + i);
            r[4 * i] = t4.x; r[4 * i + 1] = t4.y; r[4 * i + 2] = t4.z; r[4 * i + 3] = t4.w;
          }
          tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64 + c * 16), r);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // Sequence slot of (half, j): half*32 + wq*8 + j, holding sequence q0 + (half*4 + wq)*spq + j when j < spq.
    // ---- initial state: c into its shared-memory slots (a thread only ever touches its own), h0 into the B tiles
    float* cu = c_sm + (wq * SPH) * LT_H + u;  // c of (half, j) at cu[(half*32 + j) * LT_H]
    // This thread's 16-bit slot in the swizzled h tiles: k-tile q (its warp's 32 units), 16-byte chunk lane / 8; the row of
    // slot (half, j) is 64 B further per sequence and swaps chunks by (j / 2) % 4 (the slot base is a multiple of 8), so
    // four base pointers cover every (half, j) with compile-time offsets.
    uint8_t* hb[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) hb[x] = h_sm + q * LT_HTILE + (wq * SPH) * 64 + ((((lane >> 3) ^ x) & 3) << 4) + (lane & 7) * 2;
    // bf16 hi/lo split of two values with packed conversions (F2FP on the ALU pipe; the scalar F2F.BF16 runs on the
    // XU pipe, which the ex2/rcp of the gates already saturate) and 16-bit stores into the swizzled h tiles
    auto store_h2 = [&](float ha, float hb_, int half, int j) {  // slots (half, j) and (half, j + 1); j even, compile-time
      const __nv_bfloat162 ph = __floats2bfloat162_rn(ha, hb_);
      const uint32_t hbits = *reinterpret_cast<const uint32_t*>(&ph);
      const __nv_bfloat162 pl = __floats2bfloat162_rn(ha - __uint_as_float(hbits << 16), hb_ - __uint_as_float(hbits & 0xFFFF0000u));
      const uint32_t lbits = *reinterpret_cast<const uint32_t*>(&pl);
      uint8_t* o = hb[(j >> 1) & 3] + (half * LT_NH + j) * 64;
      *reinterpret_cast<uint16_t*>(o) = (uint16_t)(hbits & 0xFFFFu);
      *reinterpret_cast<uint16_t*>(o + 64) = (uint16_t)(hbits >> 16);
      *reinterpret_cast<uint16_t*>(o + 4 * LT_HTILE) = (uint16_t)(lbits & 0xFFFFu);
      *reinterpret_cast<uint16_t*>(o + 4 * LT_HTILE + 64) = (uint16_t)(lbits >> 16);
    };
    int nv[2];  // real sequences of this warp in each half (may be <= 0)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int64_t qb = q0 + (int64_t)(half * NW + wq) * spq;
      nv[half] = (int)((d.n_seq - qb) < spq ? (d.n_seq - qb) : spq);
#pragma unroll
      for (int j = 0; j < SPH; j += 2) {
        float h2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          // unused slots (j >= spq, or beyond the last sequence) are exact shadows of sequence n_seq - 1
          const int64_t qq = (j + e < nv[half]) ? qb + j + e : d.n_seq - 1;
          const int64_t so = ((int64_t)dir * d.n_seq + qq) * Hr + u;
          cu[(half * LT_NH + j + e) * LT_H] = (live && d.c0) ? __ldg(d.c0 + so) : 0.f;
          h2[e] = (live && d.h0) ? __ldg(d.h0 + so) : 0.f;
        }
        store_h2(h2[0], h2[1], half, j);
      }
    }
    auto release_half = [&](int half) {  // h_t of this half is in shared memory (async proxy), its accumulators are read
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_h + 8 * half);
    };
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { mbar_arrive(bar_h); mbar_arrive(bar_h + 8); }

    // gx row layout: native [dir][gate][unit] (4 loads of one float, a warp reads 128 B each) or, when the host permuted
    // the rows of W_ih (gx_interleaved), [dir][unit][gate]: one 16-byte load per sequence, a warp reads 512 contiguous B.
    // Addresses: byte pointer of this thread's column at step t, plus row * (row width in bytes): one IMAD.WIDE per row.
    // (the empty asm statements make these values opaque: under register pressure the compiler otherwise re-derives them
    // from the descriptor inside every chunk, ~25 integer instructions per cell)
    constexpr bool gxi = kGxi;
    uint32_t Gb = (uint32_t)G * 4u, OWb = (uint32_t)OW * 4u;
    asm volatile("" : "+r"(Gb), "+r"(OWb), "+r"(nv[0]), "+r"(nv[1]));
    const char* gxu = reinterpret_cast<const char*>(d.gx + (int64_t)dir * 4 * Hr + (gxi ? 4 * u : u));
    char* outu = reinterpret_cast<char*>(d.out + (int64_t)dir * Hr + u);
    const int64_t stepg = d.step_stride * (int64_t)Gb, stepo = d.step_stride * (int64_t)OWb;
    const uint32_t* rows = reinterpret_cast<const uint32_t*>(row_s) + wq * SPH;  // unsigned: row * width is one IMAD.WIDE.U32
    // gx of a chunk travels in one of two register buffers and is requested TWO chunks ahead of its use, across the step
    // boundary: the first two chunks of step t + 1 are requested while the last two of step t are computed, so every
    // load has a chunk of cell updates plus a barrier wait to land (ncu, run 27: with the first chunk requested just
    // before the wait on the tensor core and the others one chunk ahead, the first FFMA on a loaded value held 16 % of
    // the kernel's stall samples)
    float gxa[4][CH], gxb[4][CH];
    auto load_gx = [&](float(&gx)[4][CH], const char* gxt, int half, int j0) {
      const uint4 r4 = *reinterpret_cast<const uint4*>(rows + half * LT_NH + j0);
      const uint32_t r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const float* p = reinterpret_cast<const float*>(gxt + (uint64_t)r[j] * Gb);
        if constexpr (gxi) {
          const float4 v4 = __ldg(reinterpret_cast<const float4*>(p));
          gx[0][j] = v4.x; gx[1][j] = v4.y; gx[2][j] = v4.z; gx[3][j] = v4.w;
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) gx[g][j] = __ldg(p + g * Hr);
        }
      }
    };
    const int64_t dstep = dir ? -stepg : stepg;
    if (live && d.L > 0) {
      const char* g0 = gxu + (dir ? d.L - 1 : 0) * stepg;
      load_gx(gxa, g0, 0, 0);
      if constexpr (kCh2) load_gx(gxb, g0, 0, CH);
      else load_gx(gxb, g0, 1, 0);
    }
    for (int64_t step = 0; step < d.L; ++step) {
      const int64_t t = dir ? d.L - 1 - step : step;
      const char* gxt = gxu + t * stepg;
      char* outt = outu + t * stepo;
      asm volatile("" : "+l"(gxt), "+l"(outt));
      if (!live) {  // padding lanes: h stays 0 in the B tiles; only keep the barriers in step
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(bar_mma + 8 * half, (uint32_t)(step & 1));
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_h + 8 * half);
        }
        continue;
      }
      // the lines of step t + 2 -> L2 (lanes 0-7 / 8-15 cover the slots of half A / B; 4 gates x 128 B per warp quarter)
      if (step + 2 < d.L && lane < 16 && (lane & 7) < spq) {
        const char* pn = gxt + 2 * dstep + (uint64_t)rows[(lane >> 3) * LT_NH + (lane & 7)] * Gb - (gxi ? 4 * u : u) * 4 +
                         (kGxi ? q * 512 : q * 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + g * (kGxi ? 128 : Hr * 4)));
      }
      auto chunk = [&](const float(&gx)[4][CH], int half, int j0) {
        float a[4][CH];
        const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + LT_ACC_COL + (uint32_t)(half * LT_NH + wq * SPH + j0);
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld4(tb + (uint32_t)(g * LT_N), a[g]);
        tmem_ld_wait();
        const uint4 r4 = *reinterpret_cast<const uint4*>(rows + half * LT_NH + j0);
        const uint32_t r[4] = {r4.x, r4.y, r4.z, r4.w};
        float h[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          float cv = cu[(half * LT_NH + j0 + j) * LT_H];
          h[j] = lt_cell(fmaf(gx[0][j], lt_scale(0), a[0][j]), fmaf(gx[1][j], lt_scale(1), a[1][j]),
                         fmaf(gx[2][j], lt_scale(2), a[2][j]), fmaf(gx[3][j], lt_scale(3), a[3][j]), cv);
          cu[(half * LT_NH + j0 + j) * LT_H] = cv;
          // unconditional: an unused slot shadows the last real sequence from its initial state on, so it stores the
          // same value to the same address
          *reinterpret_cast<float*>(outt + (uint64_t)r[j] * OWb) = h[j];
        }
#pragma unroll
        for (int j = 0; j < CH; j += 2) store_h2(h[j], h[j + 1], half, j0 + j);
      };
      const uint32_t par = (uint32_t)(step & 1);
      const bool more = step + 1 < d.L;
      const char* gxn = gxt + dstep;  // next step's rows (only dereferenced when there is a next step)
      mbar_wait(bar_mma, par);
      tc_fence_after();
      if constexpr (kCh2) {
        chunk(gxa, 0, 0);
        load_gx(gxa, gxt, 1, 0);
        chunk(gxb, 0, CH);
        load_gx(gxb, gxt, 1, CH);
        release_half(0);
        mbar_wait(bar_mma + 8, par);
        tc_fence_after();
        chunk(gxa, 1, 0);
        if (more) load_gx(gxa, gxn, 0, 0);
        chunk(gxb, 1, CH);
        if (more) load_gx(gxb, gxn, 0, CH);
      } else {
        chunk(gxa, 0, 0);
        if (more) load_gx(gxa, gxn, 0, 0);
        release_half(0);
        mbar_wait(bar_mma + 8, par);
        tc_fence_after();
        chunk(gxb, 1, 0);
        if (more) load_gx(gxb, gxn, 1, 0);
      }
      release_half(1);
    }
    // final states: c from its slot, h_n = this thread's own last output row (read back; same thread, same address)
    const char* outl = outu + (dir ? 0 : d.L - 1) * stepo;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll 1
      for (int j = 0; j < SPH; ++j) {
        if (!live || j >= nv[half]) break;
        const int64_t so = ((int64_t)dir * d.n_seq + q0 + (int64_t)(half * NW + wq) * spq + j) * Hr + u;
        if (d.cn) d.cn[so] = cu[(half * LT_NH + j) * LT_H];
        if (d.hn) d.hn[so] = *reinterpret_cast<const float*>(outl + (uint64_t)rows[half * LT_NH + j] * OWb);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// w_hh_t [D, H, 4H] (W_hh transposed, the layout ps_lstm takes) -> per direction [W_lo shared-memory image 128 KB |
// W_hi row-major bf16 [4*128][128] 128 KB (stored to TMEM by the kernel)]; units / k-columns beyond H are zero
__global__ void lstm_pack_kernel(const float* __restrict__ w_hh_t, int H, int D, uint8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)D * 4 * LT_H * LT_H) return;
  const int dir = (int)(i / (4 * LT_H * LT_H));
  const int r = (int)((i / LT_H) % (4 * LT_H)), k = (int)(i % LT_H);  // padded W_hh[gate*128 + unit][k]
  const bool real = (r % LT_H) < H && k < H;
  // W_hh[g*H+unit][k], times the gate's exponent scale (lt_cell takes exponents)
  const float w = real ? lt_scale(r / LT_H) * w_hh_t[((int64_t)dir * H + k) * 4 * H + (r / LT_H) * H + (r % LT_H)] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(w);
  const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
  uint8_t* o = out + (size_t)dir * (2 * LT_WHI_BYTES);
  const int g = r / LT_H, row = r % LT_H, kt = k / 32, kk = k % 32;
  const size_t off = (size_t)(g * 4 + kt) * LT_WTILE + lt_swz((uint32_t)row, (uint32_t)(kk >> 3)) + (size_t)(kk & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(o + off) = l;                                                // shared-memory tile image: lo
  *reinterpret_cast<__nv_bfloat16*>(o + LT_WHI_BYTES + ((size_t)r * LT_H + k) * 2) = h;          // row-major (-> TMEM): hi
}

bool lstm_tc_eligible(const ps_lstm_t& d) {
  static std::atomic<int> off_c{-1};  // PS_LSTM=simt forces the fp32 kernel
  int off = off_c.load(std::memory_order_relaxed);
  if (off < 0) { const char* e = getenv("PS_LSTM"); off = (e && e[0] == 's') ? 1 : 0; off_c.store(off, std::memory_order_relaxed); }
  // the kernel keeps 32-bit row indices: the last position touched must stay below 2^31
  const int64_t last_pos = ((d.n_seq - 1) / d.inner) * d.outer_stride + (d.inner - 1) * d.inner_stride + (d.L - 1) * d.step_stride;
  if (last_pos >= 2147483647LL || d.outer_stride < 0 || d.inner_stride < 0 || d.step_stride < 0) return false;
  return !off && d.w_packed != nullptr && d.H <= LT_H && d.H % 32 == 0 && (reinterpret_cast<uintptr_t>(d.w_packed) & 15) == 0 &&
         (!d.gx_interleaved || (reinterpret_cast<uintptr_t>(d.gx) & 15) == 0);
}

int lstm_tc_launch(const ps_lstm_t& d, cudaStream_t s) {
  static SmemOnce<4> once;
  int dev = 0, n_sm = 0;
  if (int rc = current_device(&dev)) return rc;
  if (int rc = sm_count_of(dev, &n_sm)) return rc;
  const char* where = "cudaFuncSetAttribute(lstm_tc_kernel)";
  if (int rc = once.ensure(dev, 0, lstm_tc_kernel<false, false>, LT_SMEM, where)) return rc;
  if (int rc = once.ensure(dev, 1, lstm_tc_kernel<false, true>, LT_SMEM, where)) return rc;
  if (int rc = once.ensure(dev, 2, lstm_tc_kernel<true, false>, LT_SMEM, where)) return rc;
  if (int rc = once.ensure(dev, 3, lstm_tc_kernel<true, true>, LT_SMEM, where)) return rc;
  // Sequences per CTA = 8 * spq.  A step costs about the same up to 32 sequences (the 192 MMAs of a step are the floor) and
  // grows with the number of 4-sequence chunks per warp beyond (measured on B200, H = 128: 3.7 / 4.7 / 5.1 us for
  // <= 32 / 48 / 64 sequences, profiles/r01_s3_lstm_notes.md); the launch takes ceil(CTAs / SMs) waves of L such steps.
  // Pick the cheapest, larger CTAs on a tie.
  static EnvInt spq_e;
  const int spq_env = spq_e.get("PS_LSTM_SPQ", 0);
  int spq = 8;
  if (spq_env >= 1 && spq_env <= 8) {
    spq = spq_env;
  } else {
    static const int step_cost[9] = {0, 37, 37, 37, 37, 47, 47, 51, 51};
    int64_t best = 0;
    for (int c = 8; c >= 1; --c) {
      const int64_t waves = cdiv(cdiv(d.n_seq, 8 * c) * d.D, n_sm);
      const int64_t cost = waves * step_cost[c];
      if (c == 8 || cost < best) { best = cost; spq = c; }
    }
  }
#ifdef PS_EXPERIMENTS
  static EnvInt dbg_e;  // bottleneck experiments (results are garbage): compiled out of the release build
  const int dbg = dbg_e.get("PS_LSTM_DBG", 0);
#else
  const int dbg = 0;
#endif
  const int64_t nblk = cdiv(d.n_seq, 8 * spq);
  if (nblk > 2147483647LL) return PS_ERR_UNSUPPORTED;
  dim3 grid((unsigned)nblk, (unsigned)d.D);
  if (d.gx_interleaved) {
    if (spq > 4) lstm_tc_kernel<true, true><<<grid, LT_THREADS, LT_SMEM, s>>>(d, spq, dbg);
    else lstm_tc_kernel<true, false><<<grid, LT_THREADS, LT_SMEM, s>>>(d, spq, dbg);
  } else {
    if (spq > 4) lstm_tc_kernel<false, true><<<grid, LT_THREADS, LT_SMEM, s>>>(d, spq, dbg);
    else lstm_tc_kernel<false, false><<<grid, LT_THREADS, LT_SMEM, s>>>(d, spq, dbg);
  }
  PS_CHECK_LAUNCH("lstm_tc_kernel");
  return PS_OK;
}

}  // namespace ps

namespace ps {
int lstm_simt_pack(const float* w_hh_t, int64_t H, int32_t D, void* packed, cudaStream_t s);
}

// sizes the tensor-core kernel does not serve get the CUDA-core kernel's gate-minor fp32 image instead
bool ps_lstm_packed_is_simt(int64_t H) { return H >= 1 && H <= 256 && (H < 32 || H > ps::LT_H || H % 32 != 0); }

extern "C" int64_t ps_lstm_packed_bytes(int64_t H, int32_t D) {
  if (D != 1 && D != 2) return 0;
  if (ps_lstm_packed_is_simt(H)) return (int64_t)D * H * 4 * H * (int64_t)sizeof(float);
  if (H < 32 || H > ps::LT_H || H % 32 != 0) return 0;
  return (int64_t)D * 2 * ps::LT_WHI_BYTES;
}

extern "C" int ps_lstm_pack_weights(const float* w_hh_t, int64_t H, int32_t D, void* packed, void* stream) {
  PS_REQUIRE(w_hh_t && packed);
  if (ps_lstm_packed_bytes(H, D) == 0) return PS_ERR_UNSUPPORTED;
  if (ps_lstm_packed_is_simt(H)) return ps::lstm_simt_pack(w_hh_t, H, D, packed, (cudaStream_t)stream);
  const int64_t n = (int64_t)D * 4 * ps::LT_H * ps::LT_H;
  ps::lstm_pack_kernel<<<(unsigned)ps::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w_hh_t, (int)H, D, reinterpret_cast<uint8_t*>(packed));
  PS_CHECK_LAUNCH("lstm_pack_kernel");
  return PS_OK;
}
