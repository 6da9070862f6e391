// Column sums of a (rows, cols) activation-gradient matrix: the bias gradient of a Linear / pointwise conv.
//
// Scope note ("fusable ... next bandwidth kernels"): every ConMamba layer has six Linear
// layers with a bias around the Mamba block (reference modules/Conmamba.py:595-621, speechbrain PositionalwiseFeedForward,
// ConvolutionModule); torch reduces each bias gradient with a generic reduce_kernel that runs at 1-1.5 TB/s on B200 -
// 5.3 of 69 ms of the ConMamba-large step (profiles/r01_step_profile_large_final.txt).  This kernel streams the matrix
// once with 4-byte / 8-byte accesses, a thread owning two adjacent columns and four independent running sums; CTAs write
// partial rows that cm_reduce_multi adds in a fixed order (deterministic, no atomics).  Roof: HBM; bytes: rows*cols*s.
#include <cstdlib>

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

// ---- round 2: 16-byte accesses, eight rows in flight per thread ------------------------------------------------------
// The kernel above keeps 16 bytes per thread in flight and, at 256 columns, only 256 CTAs of 128 threads exist: 3.5 KB per
// SM against the ~35 KB that HBM3e latency x bandwidth asks for - 0.11 of the HBM peak at 32064 x 256.  Here a thread owns
// the 16-byte chunk of a row (8 bf16 / 4 fp32 columns), a CTA of 256 threads covers 256 / (cols / VEC) rows per pass, and
// every thread issues 8 row loads before the first add (128 bytes in flight per thread).  The CTA's row groups are summed
// in shared memory in a fixed order: one partial row per CTA as before.
template <typename T> struct CsVec;
template <> struct CsVec<float> {
  static constexpr int VEC = 4;
  using Raw = float4;
  static __device__ __forceinline__ Raw ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void acc(const Raw& v, float (&a)[VEC]) { a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w; }
};
template <> struct CsVec<__nv_bfloat16> {
  static constexpr int VEC = 8;
  using Raw = uint4;
  static __device__ __forceinline__ Raw ld(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void acc(const Raw& v, float (&a)[VEC]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[2 * i] += __uint_as_float(w[i] << 16); a[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u); }
  }
};
template <> struct CsVec<__half> {
  static constexpr int VEC = 8;
  using Raw = uint4;
  static __device__ __forceinline__ Raw ld(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void acc(const Raw& v, float (&a)[VEC]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      a[2 * i] += f.x; a[2 * i + 1] += f.y;
    }
  }
};

constexpr int kCsvThreads = 256;
constexpr int kCsvUnroll = 8;

// tpr = threads per row (cols / VEC, a power of two <= 256 dividing 256); rpp = 256 / tpr rows per pass
template <typename T>
__global__ void __launch_bounds__(kCsvThreads) colsum_vec_kernel(const T* __restrict__ x, int64_t rows, int cols, int64_t stride,
                                                                float* __restrict__ part, int tpr) {
  constexpr int VEC = CsVec<T>::VEC;
  __shared__ float red[kCsvThreads][VEC + 1];
  const int tx = threadIdx.x % tpr, ty = threadIdx.x / tpr;
  const int rpp = kCsvThreads / tpr;
  const int c = (blockIdx.x * tpr + tx) * VEC;
  float a[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) a[i] = 0.f;
  const int64_t step = (int64_t)gridDim.y * rpp;
  const T* p = x + c;
  int64_t r = (int64_t)blockIdx.y * rpp + ty;
  if (c < cols) {
    for (; r + (kCsvUnroll - 1) * step < rows; r += kCsvUnroll * step) {
      typename CsVec<T>::Raw v[kCsvUnroll];
#pragma unroll
      for (int u = 0; u < kCsvUnroll; ++u) v[u] = CsVec<T>::ld(p + (r + u * step) * stride);
#pragma unroll
      for (int u = 0; u < kCsvUnroll; ++u) CsVec<T>::acc(v[u], a);
    }
    for (; r < rows; r += step) CsVec<T>::acc(CsVec<T>::ld(p + r * stride), a);
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) red[threadIdx.x][i] = a[i];
  __syncthreads();
  if (ty == 0 && c < cols) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float sacc = 0.f;
      for (int y = 0; y < rpp; ++y) sacc += red[y * tpr + tx][i];     // fixed order: deterministic
      part[(int64_t)blockIdx.y * cols + c + i] = sacc;
    }
  }
}

template <typename T>
static bool colsum_vec_launch(const T* x, int64_t rows, int32_t cols, int64_t stride, float* part, int n_part, cudaStream_t st) {
  constexpr int VEC = CsVec<T>::VEC;
  if (getenv("CM_COLSUM_NO_VEC") != nullptr) return false;
  if (cols % VEC != 0 || stride % VEC != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return false;
  // measured on B200 (32064 rows, with the reducer launch): 256 columns 19.4 vs 23.5 us, 1024 columns 45.9 vs 29.5 us - the
  // pair kernel already has 1024 CTAs there; the vector kernel takes the narrow matrices only
  int tpr = cols / VEC;
  const int gx = 1;
  if (tpr > 32 && getenv("CM_COLSUM_VEC_ALL") == nullptr) return false;
  if (tpr > kCsvThreads) return false;
  if ((tpr & (tpr - 1)) != 0) return false;        // power of two: divides 256
  colsum_vec_kernel<T><<<dim3(gx, n_part), kCsvThreads, 0, st>>>(x, rows, cols, stride, part, tpr);
  return true;
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
  {
    bool done = false;
    switch (dtype) {
      case CM_F32: done = cm::colsum_vec_launch<float>(static_cast<const float*>(x), rows, cols, row_stride, part, grid.y, st); break;
      case CM_BF16: done = cm::colsum_vec_launch<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), rows, cols, row_stride, part, grid.y, st); break;
      default: done = cm::colsum_vec_launch<__half>(static_cast<const __half*>(x), rows, cols, row_stride, part, grid.y, st); break;
    }
    if (done) {
      CM_LAUNCH_CHECK();
      return 0;
    }
  }
  switch (dtype) {
    case CM_F32: cm::colsum_kernel<float><<<grid, cm::kCsThreads, 0, st>>>(static_cast<const float*>(x), rows, cols, row_stride, part); break;
    case CM_BF16: cm::colsum_kernel<__nv_bfloat16><<<grid, cm::kCsThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x), rows, cols, row_stride, part); break;
    default: cm::colsum_kernel<__half><<<grid, cm::kCsThreads, 0, st>>>(static_cast<const __half*>(x), rows, cols, row_stride, part); break;
  }
  CM_LAUNCH_CHECK();
  return 0;
}
