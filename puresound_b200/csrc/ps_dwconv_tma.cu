// Dilated depthwise Conv1d (3 taps) as a TMA-fed sliding window (sm_100a) - the kernel BASELINE.json's north_star names:
// "depthwise dilated conv with a TMA/shared-memory halo ... fused gLN as a single-pass Welford ... fused with PReLU".
//
//   y[b,t,c] = bias[c] + sum_p w[c,p] * u[b, t + (p - P/2) d, c]      (causal: t - (P-1-p) d),   u = PReLU(norm(x))
//
// dwconv_tile_kernel (ps_dwconv.cu) gives every CTA one time tile plus its halo: the halo rows are read again by the
// neighbouring tile (1.5x reads at d = 128, 1.19x averaged over the eight dilations of a TCN repeat), the tile is staged with
// register loads (8 x LDG.128 per thread in flight, then a barrier, then the taps - load and compute never overlap inside a
// CTA), and 16 bytes of spill at the 64-register build that four CTAs per SM need.  Here a CTA owns 32 channels (one
// 128-byte row segment) of one item for a LONG run of frames and slides a ring of frame rows through shared memory:
//
//   * rows enter the ring by TENSOR-MAP TMA (cp.async.bulk.tensor, 3-d map (channel, frame, item), box 32 x 64 x 1): one
//     elected thread, no registers, no address math, frames before / after the item zero-filled by the hardware; chunks are
//     requested 2-3 iterations ahead, so the copy of chunk k+3 overlaps the taps of chunk k inside the CTA;
//   * every row is read from HBM ONCE per run (the halo is re-read only where two runs of an item meet: 1.03x on average);
//   * a row is normalised + PReLU'd in place once, when it enters (the zero padding is applied after the prologue, like
//     the reference pads the activated tensor: lobe/cnn.py:58-74), and then serves its three taps from shared memory;
//   * outputs leave as coalesced 128-bit stores, the Welford partial of the run is one slot of the statistics array.
//
// Algorithmic bytes: 2 * 4 * B * T * C (one read, one write).  Serves P = 3, C % 32 == 0, halo <= 320 frames, prologue none
// or folded affine (gLN / gGN / bN1d) + PReLU; everything else stays on ps_dwconv.cu's kernels.
#include "ps_tc_ptx.cuh"
#include "ps_tma.cuh"

namespace ps {

constexpr int DM_CH = 64;        // frames per ring chunk (one TMA box)
constexpr int DM_CG = 32;        // channels per CTA
constexpr int DM_THREADS = 256;  // 8 float4 columns x 32 rows
constexpr int DM_MAXSLOTS = 8;
constexpr int DM_CHUNK_BYTES = DM_CH * DM_CG * 4;  // 8 KB

// Packed fp32 arithmetic (sm_100: fma / add / mul .f32x2, SASS FFMA2 / FADD2 / FMUL2): two lanes of a float4 per instruction,
// each half rounded like the scalar operation.  In the cfg2 step the kernel is issue-bound at the power-capped clock (ncu,
// run 34: issue slots 78 % busy), so halving the FMA-pipe instructions of the taps, the statistics and the prologue is time.
__device__ __forceinline__ uint64_t dm_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void dm_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t dm_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t dm_add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t dm_mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int PRO>
__global__ void __launch_bounds__(DM_THREADS) dwconv_tma_kernel(const ps_dwconv_t d, const __grid_constant__ CUtensorMap xmap, const int rows_per_cta,
                                                               const int nslots, const int HC, const int rev) {
  extern __shared__ __align__(128) uint8_t ring[];  // [nslots][64 rows][32 channels] fp32
  __shared__ __align__(8) uint64_t full_bar[DM_MAXSLOTS];
  __shared__ Wf red[DM_THREADS / 32];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  // rev: the launch walks items and runs from the last to the first (CTAs are dispatched in blockIdx order)
  const int cg = blockIdx.x, split = rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int64_t b = rev ? (int64_t)(gridDim.z - 1 - blockIdx.z) : (int64_t)blockIdx.z;
  const int T = (int)d.T, C = (int)d.C;
  const int c0 = cg * DM_CG + tx * 4;
  const int dil = d.dilation;
  const int halo = 2 * dil;
  const int halo_l = d.causal ? halo : dil;
  const int ta = split * rows_per_cta;
  const int tb = (ta + rows_per_cta) < T ? (ta + rows_per_cta) : T;
  const uint32_t ring_u = smem_u32(ring), bar_u = smem_u32(full_bar);

  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) mbar_init(bar_u + 8 * s, 1);
    fence_barrier_init();
    tma_prefetch_desc(&xmap);
  }
  __syncthreads();
  pdl_wait();  // the producer of x and the statistics merge that made pro_a / pro_b are complete
  // statistics slots no run of this launch writes must read as empty partials (count 0): the first CTA of every item zeroes
  // them (the launcher used to memset the whole array, a graph node that cut the programmatic launch chain)
  if (d.stats_partials && cg == 0 && split == 0) {
    const int64_t used = (int64_t)gridDim.x * gridDim.y;
    float* sp = d.stats_partials + b * d.stats_slots * 3;
    for (int64_t i = used * 3 + tid; i < d.stats_slots * 3; i += DM_THREADS) sp[i] = 0.f;
  }
  if (ta >= T) return;  // (cannot happen: the grid has ceil(T / rows_per_cta) runs)

  const int n_out = (tb - ta + DM_CH - 1) / DM_CH;  // output chunks of this run
  const int total_in = n_out + HC;                  // input chunks: rows [ta - halo_l, ...)
  const int in_row0 = ta - halo_l;
  // (every ring index below is a running counter with a wrap: `i % nslots` / `i / nslots` with a run-time nslots were two
  // integer divisions per chunk and thread - the kernel executed 146 instructions per 16-byte output, ncu run 39)
  int issued = 0, s_is = 0;  // (thread 0) input chunks requested so far, slot of the next request
  auto issue_until = [&](int limit) {  // request chunks [issued, min(limit, total_in)): chunk i -> slot i mod nslots
    while (issued < limit && issued < total_in) {
      mbar_arrive_expect_tx(bar_u + 8 * s_is, DM_CHUNK_BYTES);
      tma_load_3d(ring_u + s_is * DM_CHUNK_BYTES, &xmap, cg * DM_CG, in_row0 + issued * DM_CH, (int)b, bar_u + 8 * s_is);
      ++issued;
      s_is = s_is + 1 == nslots ? 0 : s_is + 1;
    }
  };
  if (tid == 0) issue_until(nslots);

  const float slope = (d.pro_act == PS_ACT_PRELU && d.pro_slope) ? __ldg(d.pro_slope) : 1.f;
  const bool slope01 = slope > 0.f && slope <= 1.f;  // PReLU(z) = max(z, slope * z) then: one FMNMX instead of compare + select
  float pa[4] = {1.f, 1.f, 1.f, 1.f}, pb[4] = {0.f, 0.f, 0.f, 0.f}, bias[4] = {0.f, 0.f, 0.f, 0.f}, w[3][4];
  if constexpr (PRO == 1) {
    const float4 a4 = __ldg(reinterpret_cast<const float4*>(d.pro_a + b * d.pro_batch_stride + c0));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(d.pro_b + b * d.pro_batch_stride + c0));
    pa[0] = a4.x; pa[1] = a4.y; pa[2] = a4.z; pa[3] = a4.w;
    pb[0] = b4.x; pb[1] = b4.y; pb[2] = b4.z; pb[3] = b4.w;
  }
  if (d.bias) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(d.bias + c0));
    bias[0] = b4.x; bias[1] = b4.y; bias[2] = b4.z; bias[3] = b4.w;
  }
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int i = 0; i < 4; ++i) w[p][i] = __ldg(d.w + (c0 + i) * 3 + p);

  const uint64_t pa01 = dm_pack(pa[0], pa[1]), pa23 = dm_pack(pa[2], pa[3]), pb01 = dm_pack(pb[0], pb[1]), pb23 = dm_pack(pb[2], pb[3]);
  const uint64_t bias01 = dm_pack(bias[0], bias[1]), bias23 = dm_pack(bias[2], bias[3]), slope2 = dm_pack(slope, slope);
  uint64_t w01[3], w23[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) { w01[p] = dm_pack(w[p][0], w[p][1]); w23[p] = dm_pack(w[p][2], w[p][3]); }
  // The ring is [nslots * 64 rows][32 channels] contiguous: tap p of this thread's row h reads ring row
  // (64 sk + ty + 32 h + p dil) mod (64 nslots).  Byte offsets that do not depend on the chunk are kept in registers (made
  // opaque: the compiler otherwise re-derives them from ty, tx, dil in every iteration - ~25 integer instructions per row).
  uint32_t tapb[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    tapb[p] = (uint32_t)(((ty + p * dil) * DM_CG + tx * 4) * 4);
    asm volatile("" : "+r"(tapb[p]));
  }
  const uint32_t ring_bytes = (uint32_t)nslots * DM_CHUNK_BYTES;
  // statistics: shifted sums about a pivot (this thread's first output), kept as two packed pairs of partial sums
  float piv = 0.f;
  uint64_t npiv2 = dm_pack(0.f, 0.f), ssum01 = npiv2, ssum23 = npiv2, ssq01 = npiv2, ssq23 = npiv2;
  int nout = 0;
  float* yrow[2];  // output pointer of this thread's two rows of the current chunk
#pragma unroll
  for (int h = 0; h < 2; ++h) yrow[h] = d.y + (b * T + ta + ty + 32 * h) * (int64_t)C + c0;
  int tr = 0, s_tr = 0;  // input chunks transformed so far, slot of the next one
  uint32_t ph_tr = 0;    // its mbarrier phase
  int t_in = in_row0 + ty;  // frame of this thread's first row of input chunk tr
  int sk = 0;               // slot of input chunk k
  for (int k = 0; k < n_out; ++k) {
    // ---- input chunks up to k + HC have landed and carry PReLU(norm(x)), zero outside the item
    while (tr <= k + HC && tr < total_in) {
      mbar_wait(bar_u + 8 * s_tr, ph_tr);
      if constexpr (PRO == 1) {
        float* cp = reinterpret_cast<float*>(ring + s_tr * DM_CHUNK_BYTES) + ty * DM_CG + tx * 4;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int t = t_in + 32 * h;
          float4* q = reinterpret_cast<float4*>(cp + 32 * h * DM_CG);
          if (t >= 0 && t < T) {
            const float4 v = *q;
            uint64_t z01 = dm_fma2(dm_pack(v.x, v.y), pa01, pb01), z23 = dm_fma2(dm_pack(v.z, v.w), pa23, pb23);
            float o[4], m[4];
            dm_unpack(z01, o[0], o[1]);
            dm_unpack(z23, o[2], o[3]);
            dm_unpack(dm_mul2(z01, slope2), m[0], m[1]);
            dm_unpack(dm_mul2(z23, slope2), m[2], m[3]);
            if (slope01) {
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = fmaxf(o[i], m[i]);  // (a NaN z gives a NaN slope * z: max(NaN, NaN) = NaN)
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = o[i] > 0.f ? o[i] : m[i];
            }
            *q = make_float4(o[0], o[1], o[2], o[3]);
          } else {
            *q = make_float4(0.f, 0.f, 0.f, 0.f);  // zero padding applies AFTER the prologue: the reference pads the activated tensor
          }
        }
      }
      ++tr;
      t_in += DM_CH;
      if (++s_tr == nslots) { s_tr = 0; ph_tr ^= 1; }
    }
    __syncthreads();  // transformed rows visible to every thread; every thread is done with the taps of chunk k - 1
    if (tid == 0) {
      fence_proxy_async();        // this CTA's generic-proxy writes to the slots about to be overwritten by the async proxy
      issue_until(k + nslots);    // chunks < k are dead: their slots take the chunks up to k + nslots - 1
    }
    // ---- taps of output chunk k from the ring
    const int t0 = ta + k * DM_CH + ty;
    const uint32_t skb = (uint32_t)sk * DM_CHUNK_BYTES;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (t0 + 32 * h < tb) {
        uint64_t a01 = bias01, a23 = bias23;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          uint32_t off = skb + tapb[p] + (uint32_t)(h * 32 * DM_CG * 4);  // < 2 * ring_bytes: the halo is at most nslots - 3 chunks
          off = off >= ring_bytes ? off - ring_bytes : off;
          const float4 u4 = *reinterpret_cast<const float4*>(ring + off);
          a01 = dm_fma2(w01[p], dm_pack(u4.x, u4.y), a01);
          a23 = dm_fma2(w23[p], dm_pack(u4.z, u4.w), a23);
        }
        float acc[4];
        dm_unpack(a01, acc[0], acc[1]);
        dm_unpack(a23, acc[2], acc[3]);
        *reinterpret_cast<float4*>(yrow[h]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (nout == 0) { piv = acc[0]; npiv2 = dm_pack(-piv, -piv); }
        ++nout;
        const uint64_t d01 = dm_add2(a01, npiv2), d23 = dm_add2(a23, npiv2);
        ssum01 = dm_add2(ssum01, d01);
        ssum23 = dm_add2(ssum23, d23);
        ssq01 = dm_fma2(d01, d01, ssq01);
        ssq23 = dm_fma2(d23, d23, ssq23);
      }
      yrow[h] += (int64_t)DM_CH * C;
    }
    sk = sk + 1 == nslots ? 0 : sk + 1;
  }
  if (d.stats_partials) {
    float s0, s1, s2, s3, q0, q1, q2, q3;
    dm_unpack(ssum01, s0, s1); dm_unpack(ssum23, s2, s3);
    dm_unpack(ssq01, q0, q1); dm_unpack(ssq23, q2, q3);
    const float ssum = (s0 + s1) + (s2 + s3), ssq = (q0 + q1) + (q2 + q3);
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
    if (tid == 0) {
      const int64_t slot = (int64_t)split * gridDim.x + cg;
      float* o = d.stats_partials + (b * d.stats_slots + slot) * 3;
      o[0] = tot.n; o[1] = tot.mean; o[2] = tot.m2;
      if (d.fin_scale) {
        // fused gLN / gGN finalize (few-frame launches, where a separate merge launch costs as much as this kernel): the
        // CTA that writes an item's last partial merges them all, in slot order (see ps_gemm_pair.cu).  The zeroed tail
        // slots were written by this item's first CTA before the barrier inside wf_block_reduce, i.e. before ITS count.
        __threadfence();
        const unsigned int total = gridDim.x * gridDim.y;
        fin_last = (atomicAdd(d.fin_counter + b, 1u) + 1u == total) ? 1 : 0;
      }
    }
    if (d.fin_scale) {
      static_assert(DM_THREADS == 256, "the fused finalize runs on a 256-thread CTA");
      __shared__ double fin_s[3 * 256];
      __syncthreads();
      if (fin_last) {
        __threadfence();
        stats_finalize_item<0>(d.stats_partials + b * d.stats_slots * 3, d.stats_slots, d.fin_gamma, d.fin_beta, d.fin_eps, d.C,
                               d.fin_scale + b * d.C, d.fin_shift + b * d.C, nullptr, tid, fin_s, fin_s + 256, fin_s + 512);
        if (tid == 0) d.fin_counter[b] = 0;
      }
    }
  }
}

// ---------------------------------------------------------------- dispatch (called by ps_dwconv)
bool dwconv_tma_eligible(const ps_dwconv_t& d) {
  static EnvInt env;
  if (env.get("PS_DW_TMA", 1) == 0) return false;  // A/B: PS_DW_TMA=0 keeps dwconv_tile_kernel
  if (d.P != 3 || d.C % DM_CG != 0 || d.dilation < 1 || d.T < DM_CH) return false;
  if (!(d.pro_mode == PS_PRO_NONE || (d.pro_mode == PS_PRO_AFFINE && (d.pro_act == PS_ACT_PRELU || d.pro_act == PS_ACT_NONE)))) return false;
  const int halo = 2 * d.dilation;
  const int HC = (halo + DM_CH - 1) / DM_CH;
  if (HC + 3 > DM_MAXSLOTS) return false;
  if (d.T >= (1LL << 30) || d.batch >= (1LL << 31) || d.T * d.C * 4 >= (1LL << 40)) return false;
  if ((reinterpret_cast<uintptr_t>(d.x) | reinterpret_cast<uintptr_t>(d.y)) & 15) return false;
  if (d.pro_mode == PS_PRO_AFFINE && (((reinterpret_cast<uintptr_t>(d.pro_a) | reinterpret_cast<uintptr_t>(d.pro_b)) & 15) || (d.pro_batch_stride & 3))) return false;
  if (d.bias && (reinterpret_cast<uintptr_t>(d.bias) & 15)) return false;
  return true;
}

// rows of one CTA's run: long enough that the halo re-read where two runs meet stays below ~12 %, short enough for >= ~4
// CTAs per SM worth of runs; a multiple of the chunk
static int dm_rows_per_cta(const ps_dwconv_t& d, int sms) {
  const int halo = 2 * d.dilation;
  int64_t rows = halo * 8 > 1024 ? halo * 8 : 1024;
  const int64_t groups = d.batch * (d.C / DM_CG);
  // small batches: split further until there are ~4 runs per SM (a run is never shorter than 4 chunks)
  while (rows > 4 * DM_CH && groups * cdiv(d.T, rows) < 4LL * sms) rows /= 2;
  rows = cdiv(rows, DM_CH) * DM_CH;
  return (int)rows;
}

int64_t dwconv_tma_slots(int64_t T, int64_t C) {
  // upper bound of runs per item over every dilation / batch: the shortest run is 4 chunks
  return cdiv(T, 4 * DM_CH) * (C / DM_CG);
}

template <int PRO>
static int launch_dm(const ps_dwconv_t& d, const CUtensorMap& xmap, int rows, int nslots, int HC, dim3 grid, cudaStream_t s, int dev) {
  static SmemOnce<1> once;
  if (int rc = once.ensure(dev, 0, dwconv_tma_kernel<PRO>, DM_MAXSLOTS * DM_CHUNK_BYTES, "cudaFuncSetAttribute(dwconv_tma_kernel)")) return rc;
  cudaError_t le = launch_pdl(dwconv_tma_kernel<PRO>, grid, dim3(DM_THREADS), (size_t)nslots * DM_CHUNK_BYTES, s, d, xmap, rows, nslots, HC, order_reversed() ? 1 : 0);
  if (le != cudaSuccess) { set_cuda_error(le, "dwconv_tma_kernel"); return PS_ERR_CUDA; }
  return PS_OK;
}

int dwconv_tma_launch(const ps_dwconv_t& d, cudaStream_t s) {
  int dev = 0, sms = 0;
  if (int rc = current_device(&dev)) return rc;
  if (int rc = sm_count_of(dev, &sms)) return rc;
  const int halo = 2 * d.dilation;
  const int HC = (halo + DM_CH - 1) / DM_CH;
  const int nslots = HC + 3;
  const int rows = dm_rows_per_cta(d, sms);
  CUtensorMap xmap;
  const uint64_t dims[3] = {(uint64_t)d.C, (uint64_t)d.T, (uint64_t)d.batch};
  const uint64_t strides[2] = {(uint64_t)d.C * 4, (uint64_t)d.T * d.C * 4};
  const uint32_t box[3] = {DM_CG, DM_CH, 1};
  if (int rc = tma_encode_f32(&xmap, d.x, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  dim3 grid((unsigned)(d.C / DM_CG), (unsigned)cdiv(d.T, rows), (unsigned)d.batch);
  if (grid.y > 65535 || grid.z > 65535) return PS_ERR_UNSUPPORTED;
  if (d.pro_mode == PS_PRO_AFFINE) return launch_dm<1>(d, xmap, rows, nslots, HC, grid, s, dev);
  return launch_dm<0>(d, xmap, rows, nslots, HC, grid, s, dev);
}

}  // namespace ps
