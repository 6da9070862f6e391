// TMA (cp.async.bulk.tensor) + mbarrier helpers shared by the sm_100a scan kernels.
//
// Host side: 3-D tiled tensor maps over (channel, time, batch) views of channel-last activations, encoded with the
// driver's cuTensorMapEncodeTiled (resolved through cudaGetDriverEntryPoint: the library links only libcudart).
// Device side: one elected lane arms an mbarrier with the byte count of a tile and issues the tensor copies; the data
// lands in shared memory through the async proxy and completes the barrier's transaction count.  Rows outside the tensor
// (time < 0 or >= L) are zero-filled by the hardware, which is what the ragged first / last tiles of a scan need.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm {
namespace tma {

// ---- host ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  // function-local static: initialised once, thread-safe (C++11); holds a driver entry point, no device state
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Map over a (batch, time, channel) view: element (b, l, c) at base + (b*sb + l*sl + c) * es bytes (unit channel stride).
// Box = (box_c channels, box_l steps, 1 batch).  Returns false when the view is not TMA-addressable (alignment / strides).
inline bool make_map_blc(CUtensorMap* m, const void* base, int es, int64_t channels, int64_t L, int64_t batch, int64_t sl_elems,
                         int64_t sb_elems, int box_c, int box_l) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr || base == nullptr) return false;
  const uint64_t sl = (uint64_t)sl_elems * es, sb = (uint64_t)sb_elems * es;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (sl & 15) != 0 || sl_elems <= 0) return false;
  if (batch > 1 && ((sb & 15) != 0 || sb_elems <= 0)) return false;
  if ((box_c * es) % 16 != 0 || box_c > 256 || box_l > 256) return false;
  if (sl >= (1ull << 40) || sb >= (1ull << 40)) return false;
  const CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;   // bit copies
  cuuint64_t dims[3] = {(cuuint64_t)channels, (cuuint64_t)L, (cuuint64_t)batch};
  cuuint64_t strides[2] = {sl, batch > 1 ? sb : sl * (uint64_t)L};
  if ((strides[1] & 15) != 0) return false;
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_l, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(m, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// ---- device ----------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
// make barrier initialisation visible to the async proxy before the first TMA that signals it
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TMA_MBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TMA_MBAR_DONE;\n\t"
      "bra TMA_MBAR_WAIT;\n\t"
      "TMA_MBAR_DONE:\n\t}"
      ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// generic-proxy accesses of shared memory (this thread's, and through a preceding barrier the warp's) are ordered before
// subsequent async-proxy accesses: issued before a TMA load refills a stage the warp has just read, and before a TMA
// store reads a stage the warp has just written
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the same for global memory: a TMA load of data that threads of this CTA wrote with ordinary stores
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void prefetch_map(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// global -> shared tile load: coordinates (c, l, b) of the box origin, innermost first; completes `bar` by the box bytes
__device__ __forceinline__ void load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c, int l, int b) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(l), "r"(b)
      : "memory");
}
// shared -> global tile store (bulk async-group completion); rows outside the tensor are not written
__device__ __forceinline__ void store_3d(const CUtensorMap* m, const void* src, int c, int l, int b) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c), "r"(l), "r"(b)
               : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
#endif

}  // namespace tma
}  // namespace cm
