// Shared pieces of the row-normalising kernels with an activation (ln_act.cu, stem.cu): a group of G threads owns a row,
// a thread kNQ quads of 4 consecutive elements at a stride of G quads; statistics in fp32.
#pragma once
#include "sp_common.cuh"

namespace cm {
namespace lna {

using cm::sp::Quad;

constexpr int kThreads = 128;
constexpr int kNQ = 5;
constexpr int kMaxCols = kThreads * kNQ * 4;   // 2560

template <typename T> __device__ __forceinline__ void st4(void* p, const float* v);
template <> __device__ __forceinline__ void st4<float>(void* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(void* p, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = o;
}
template <> __device__ __forceinline__ void st4<__half>(void* p, const float* v) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = o;
}

// Sum of (a, b) over the G threads of a row group.  For G > 32 the warps of a group meet through `red`; the buffer index
// toggles per call, so one __syncthreads() per reduction is enough (a warp can only reach the next-but-one reduction,
// which reuses this buffer, after every warp has passed the barrier of the next one, i.e. has finished reading).
template <int G>
__device__ __forceinline__ float2 group_sum(float2 v, float2 (*red)[kThreads / 32], int& par) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  if (G == 32) return v;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[par][warp] = v;
  __syncthreads();
  constexpr int WPG = G / 32;
  const int w0 = (warp / WPG) * WPG;
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < WPG; ++w) {
    const float2 t = red[par][w0 + w];
    s.x += t.x;
    s.y += t.y;
  }
  par ^= 1;
  return s;
}

// activation of the normalised value and its derivative
template <int ACT> __device__ __forceinline__ float act_fwd(float t, float slope) {
  if (ACT == CM_LN_ACT_GELU) return gelu_f(t);
  return t > 0.f ? t : t * slope;
}
template <int ACT> __device__ __forceinline__ float act_bwd(float t, float dy, float slope) {
  if (ACT == CM_LN_ACT_GELU) return dy * gelu_grad_f(t);
  return t > 0.f ? dy : dy * slope;
}
}  // namespace lna
}  // namespace cm
