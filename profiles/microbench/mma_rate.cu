// Microbenchmark (round 2): what bounds the issue rate of cta_group::2 tcgen05.mma when operands stream from shared memory?
//   variant N=128: per k16 step and 256-channel block 3 MMAs of M=256 x N=128 x K=16 (the round-1 gemm_pair_kernel shape)
//   variant N=256: the same FLOPs as M=256 x N=256 MMAs (half as many instructions, A re-read half as often)
// with and without the two other shared-memory clients of the real kernel: the weight stream (cp.async.bulk global -> smem,
// 32 KB per 32-k stage) and the activation producers (STS.128 of 8 or 16 KB per stage).  Operands are whatever the buffers
// hold; only time matters.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../puresound_b200/csrc mma_rate.cu
#include <stdio.h>
#include <stdlib.h>

#include "ps_tc_ptx.cuh"

namespace ps {
void set_cuda_error(cudaError_t, const char*) {}
}
using namespace ps;

constexpr int STAGES = 4;

__device__ __forceinline__ uint64_t desc64(uint32_t saddr) {  // K-major, SWIZZLE_64B, 8-row atoms 512 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// N = MMA N (128 or 256).  Stage layout: X hi | X lo (N/2 rows x 32 k bf16 each) | W hi b0 | W hi b1 | W lo b0 | W lo b1 (128 x 32 each)
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
    mma_rate_kernel(const uint8_t* __restrict__ wsrc, int n_stages, int with_bulk, int with_sts, long long* cycles) {
  constexpr int XPART = (N / 2) * 32 * 2;
  constexpr int WBLK = 128 * 32 * 2;
  constexpr int STAGE = 2 * XPART + 4 * WBLK;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE;
  const uint32_t bar_full = bars, bar_empty = bars + 64;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(sm + STAGES * STAGE + 128);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32((const void*)tmem_ptr_s), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const long long t0 = clock64();
  if (warp == 0) {
    // weight stream: 32 KB per stage per CTA from an L2-resident buffer, paced by the stage ring
    if (with_bulk) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < n_stages; ++it) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_full + 8 * s, 4 * WBLK);
          bulk_g2s(base + s * STAGE + 2 * XPART, wsrc + (size_t)((blockIdx.x * 16 + (it & 15)) % 512) * (4 * WBLK), 4 * WBLK, bar_full + 8 * s);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < n_stages; ++it) {
        if (with_bulk) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
        } else if (it >= STAGES) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);  // bound the MMA queue like the real pipeline does
          tc_fence_after();
        }
        if (elect_one()) {
          const uint32_t sa = base + s * STAGE;
          const uint64_t x_hi = desc64(sa), x_lo = desc64(sa + XPART);
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            const uint64_t w_hi = desc64(sa + 2 * XPART + mb * WBLK), w_lo = desc64(sa + 2 * XPART + (2 + mb) * WBLK);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t ko = (uint64_t)((k * 32) >> 4);
              const uint32_t dd = tmem_base + (uint32_t)(mb * N);
              umma_bf16_pair(dd, w_lo + ko, x_hi + ko, IDESC, 1);
              umma_bf16_pair(dd, w_hi + ko, x_lo + ko, IDESC, 1);
              umma_bf16_pair(dd, w_hi + ko, x_hi + ko, IDESC, 1);
            }
          }
          umma_commit_pair(bar_empty + 8 * s);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      // drain
      for (int i = 0; i < STAGES; ++i) {
        const int it = n_stages - STAGES + i;
        if (it >= 0) mbar_wait(bar_empty + 8 * (it % STAGES), (uint32_t)((it / STAGES) & 1));
      }
    } else if (!with_bulk) {
      // nothing: the leader's commits also arrive here
    }
  } else if (with_sts) {
    // activation producers: XPART*2 bytes of STS.128 per stage, free-running at roughly the stage pace (no handshake)
    const int pt = tid - 64;  // 0..319 (first 256 used)
    if (pt < 256) {
      const uint32_t chunk = pt & 3, r0 = pt >> 2;
      for (int it = 0; it < n_stages; ++it) {
        uint8_t* x_hi = sm + (it % STAGES) * STAGE;
#pragma unroll
        for (int p = 0; p < N / 128; ++p) {
          const uint32_t r = r0 + p * 64;
          const uint32_t off = r * 64u + ((chunk ^ ((r >> 1) & 3u)) << 4);
          *reinterpret_cast<uint4*>(x_hi + off) = make_uint4(it, pt, r, 0);
          *reinterpret_cast<uint4*>(x_hi + XPART + off) = make_uint4(pt, it, r, 1);
        }
        // pace: one stage of MMAs is 6 * N/128 * 128 clocks
        const long long until = t0 + (long long)(it + 1) * with_sts;
        while (clock64() < until) {}
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int N>
static void run(const uint8_t* wsrc, long long* cyc_d, int n_stages, int with_bulk, int with_sts) {
  constexpr int STAGE = 2 * (N / 2) * 32 * 2 + 4 * 128 * 32 * 2;
  const int smem = STAGES * STAGE + 1024 + 256;
  cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    mma_rate_kernel<N><<<148, 384, smem>>>(wsrc, n_stages, with_bulk, with_sts, cyc_d);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { printf("N=%d bulk=%d sts=%d: %s\n", N, with_bulk, with_sts, cudaGetErrorString(e)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  long long cyc[148];
  cudaMemcpy(cyc, cyc_d, sizeof(cyc), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
  // FLOPs issued: per stage per pair 12 MMAs (N=128) of 256*128*16 MACs; N=256 stages carry twice the MACs
  const double flops = 2.0 * 74 * (double)n_stages * 12 * 256.0 * N * 16;
  const double ideal_clk = (double)n_stages * 12 * (N / 128) * 64;
  printf("N=%3d bulk=%d sts_pace=%4d: %.3f ms, %.0f TFLOP/s issued, max CTA cycles %lld (ideal MMA cycles %.0f -> tensor busy %.1f %%), %.2f GHz\n", N,
         with_bulk, with_sts, best, flops / (best * 1e-3) / 1e12, mx, ideal_clk, 100.0 * ideal_clk / (double)mx, (double)mx / (best * 1e-3) / 1e9);
}

int main() {
  uint8_t* w;
  long long* cyc;
  cudaMalloc(&w, 512 * 32768);
  cudaMemset(w, 0x11, 512 * 32768);
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int S128 = 8000, S256 = 4000;  // same FLOPs
  for (int bulk = 0; bulk < 2; ++bulk)
    for (int sts = 0; sts < 2; ++sts) {
      run<128>(w, cyc, S128, bulk, sts ? 768 : 0);
      run<256>(w, cyc, S256, bulk, sts ? 1536 : 0);
    }
  return 0;
}
