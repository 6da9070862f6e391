// Column sums of a (rows, cols) activation-gradient matrix: the bias gradient of a Linear / pointwise conv.
//
// Scope note ("fusable ... next bandwidth kernels"): every ConMamba layer has six Linear
// layers with a bias around the Mamba block (reference modules/Conmamba.py:595-621, speechbrain PositionalwiseFeedForward,
// ConvolutionModule); torch reduces each bias gradient with a generic reduce_kernel that runs at 1-1.5 TB/s on B200 -
// 5.3 of 69 ms of the ConMamba-large step (profiles/r01_step_profile_large_final.txt).  This kernel streams the matrix
// once with 4-byte / 8-byte accesses, a thread owning two adjacent columns and four independent running sums; CTAs write
// partial rows that cm_reduce_multi adds in a fixed order (deterministic, no atomics).  Roof: HBM; bytes: rows*cols*s.
#include "common.cuh"

namespace cm {

constexpr int kCsThreads = 128;   // 256 columns per CTA

template <typename T> struct Cs2;
template <> struct Cs2<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
};
template <> struct Cs2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
};
template <> struct Cs2<__half> {
  static __device__ __forceinline__ float2 ld(const __half* p) {
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(p));
    return __half22float2(*reinterpret_cast<const __half2*>(&r));
  }
};

template <typename T>
__global__ void __launch_bounds__(kCsThreads) colsum_kernel(const T* __restrict__ x, int64_t rows, int cols, int64_t stride,
                                                           float* __restrict__ part) {
  const int c = (blockIdx.x * kCsThreads + threadIdx.x) * 2;
  if (c >= cols) return;
  const int64_t step = gridDim.y;
  const T* p = x + c;
  float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  int64_t r = blockIdx.y;
  for (; r + 3 * step < rows; r += 4 * step) {
    const float2 v0 = Cs2<T>::ld(p + r * stride), v1 = Cs2<T>::ld(p + (r + step) * stride);
    const float2 v2 = Cs2<T>::ld(p + (r + 2 * step) * stride), v3 = Cs2<T>::ld(p + (r + 3 * step) * stride);
    a0 = fadd2(a0, v0); a1 = fadd2(a1, v1); a2 = fadd2(a2, v2); a3 = fadd2(a3, v3);
  }
  for (; r < rows; r += step) a0 = fadd2(a0, Cs2<T>::ld(p + r * stride));
  const float2 s = fadd2(fadd2(a0, a1), fadd2(a2, a3));
  *reinterpret_cast<float2*>(part + (int64_t)blockIdx.y * cols + c) = s;
}

}  // namespace cm

extern "C" int cm_colsum_num_part(int64_t rows) {
  if (rows <= 0) return CM_ERR_BAD_ARG;
  const int64_t want = (rows + 31) / 32;
  return (int)(want < 256 ? want : 256);
}

extern "C" int cm_colsum(const void* x, int64_t rows, int32_t cols, int64_t row_stride, int32_t dtype, float* part, void* stream) {
  if (x == nullptr || part == nullptr || rows <= 0 || cols <= 0) return CM_ERR_BAD_ARG;
  if (!cm::dtype_ok(dtype)) return CM_ERR_BAD_ARG;
  const int es = dtype == CM_F32 ? 4 : 2;
  if ((cols & 1) || (row_stride & 1) || (reinterpret_cast<uintptr_t>(x) % (2 * es)) != 0 ||
      (reinterpret_cast<uintptr_t>(part) & 7) != 0)
    return CM_ERR_UNSUPPORTED;          // pair accesses need even, aligned rows
  const dim3 grid((cols / 2 + cm::kCsThreads - 1) / cm::kCsThreads, cm_colsum_num_part(rows));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case CM_F32: cm::colsum_kernel<float><<<grid, cm::kCsThreads, 0, st>>>(static_cast<const float*>(x), rows, cols, row_stride, part); break;
    case CM_BF16: cm::colsum_kernel<__nv_bfloat16><<<grid, cm::kCsThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x), rows, cols, row_stride, part); break;
    default: cm::colsum_kernel<__half><<<grid, cm::kCsThreads, 0, st>>>(static_cast<const __half*>(x), rows, cols, row_stride, part); break;
  }
  CM_LAUNCH_CHECK();
  return 0;
}
