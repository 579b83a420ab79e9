// Tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG) helpers: host-side descriptor encoding through the driver entry point
// (no link against libcuda: the symbol is fetched with cudaGetDriverEntryPoint) and the device-side load wrappers.
#pragma once
#include <cuda.h>

#include "ps_common.cuh"

namespace ps {

typedef CUresult (*ps_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline ps_encode_tiled_fn tma_encoder() {
  static std::atomic<void*> cached{nullptr};
  void* p = cached.load(std::memory_order_acquire);
  if (p == nullptr) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    cached.store(p, std::memory_order_release);
  }
  return reinterpret_cast<ps_encode_tiled_fn>(p);
}

// fp32 tensor of up to 3 dims (dim 0 contiguous), strides in BYTES for dims 1.. (multiples of 16), box per dim, OOB -> 0
inline int tma_encode_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, CUtensorMapSwizzle swz) {
  ps_encode_tiled_fn enc = tma_encoder();
  if (!enc) { set_cuda_error(cudaErrorNotSupported, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)"); return PS_ERR_CUDA; }
  cuuint64_t gd[3], gs[2];
  cuuint32_t bx[3], es[3];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled"); return PS_ERR_CUDA; }
  return PS_OK;
}

#ifdef __CUDACC__
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// box load into this CTA's shared memory; completion (bytes) on the mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
#endif

}  // namespace ps
