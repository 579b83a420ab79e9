// resource probe for gemm_wide_kernel: attributes + occupancy at the requested dynamic shared memory
#include <stdio.h>
#include "../../puresound_b200/csrc/ps_gemm_wide.cu"
namespace ps { void set_cuda_error(cudaError_t e, const char* w) { printf("cuda error %s at %s\n", cudaGetErrorString(e), w); } }
int main() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, ps::gemm_wide_kernel<1>);
  printf("regs %d static smem %zu maxThreadsPerBlock %d maxDynamicSharedSizeBytes %d WD_SMEM %d threads %d\n", a.numRegs, a.sharedSizeBytes, a.maxThreadsPerBlock,
         a.maxDynamicSharedSizeBytes, ps::WD_SMEM, ps::WD_THREADS);
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  printf("device: smem/block optin %zu, smem/SM %zu, reserved/block %zu, regs/SM %d\n", p.sharedMemPerBlockOptin, p.sharedMemPerMultiprocessor, p.reservedSharedMemPerBlock, p.regsPerMultiprocessor);
  for (int smem = 200 * 1024; smem <= 232448; smem += 1024) {
    cudaError_t e = cudaFuncSetAttribute(ps::gemm_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int nb = -1;
    cudaError_t e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ps::gemm_wide_kernel<1>, ps::WD_THREADS, smem);
    if (e != cudaSuccess || e2 != cudaSuccess || nb < 1) { printf("smem %d: set %s occ %s blocks %d\n", smem, cudaGetErrorString(e), cudaGetErrorString(e2), nb); cudaGetLastError(); }
  }
  cudaFuncSetAttribute(ps::gemm_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ps::WD_SMEM);
  int nb = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ps::gemm_wide_kernel<1>, ps::WD_THREADS, ps::WD_SMEM);
  printf("at WD_SMEM: blocks/SM %d\n", nb);
  for (int thr = 512; thr <= 640; thr += 32) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ps::gemm_wide_kernel<1>, thr, 200 * 1024);
    printf("threads %d smem 200K: blocks/SM %d\n", thr, nb);
  }
  return 0;
}
