// Helpers shared by the state-parallel scan kernels (scan_fwd_sp.cu, scan_bwd_sp.cu).
#pragma once
#include "common.cuh"

namespace cm {
namespace sp {

constexpr int kT = 16;          // steps per tile
constexpr int kNW = 2;          // recurrence warps per direction
constexpr int kGT = kNW * 32;   // recurrence threads per direction
constexpr int kCH = 32;         // channels per CTA: a warp covers 8 channel pairs, a lane 4 states of one pair
constexpr int kNP = kCH / 2;    // channel pairs per CTA

// ---- paired element I/O ------------------------------------------------------------------------------------------
template <typename T> struct Pair;
template <> struct Pair<float> {
  using Raw = float2;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r; asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw ld_cg(const void* p) {
    Raw r; asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw zero() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ float2 cvt(Raw r) { return r; }
  static __device__ __forceinline__ void st(void* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <> struct Pair<__nv_bfloat16> {
  using Raw = uint32_t;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r; asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw ld_cg(const void* p) {
    Raw r; asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(r) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw zero() { return 0u; }
  static __device__ __forceinline__ float2 cvt(Raw r) {
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
  static __device__ __forceinline__ void st(void* p, float2 v) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
  }
};
template <> struct Pair<__half> {
  using Raw = uint32_t;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r; asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw ld_cg(const void* p) {
    Raw r; asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(r) : "l"(p)); return r;
  }
  static __device__ __forceinline__ Raw zero() { return 0u; }
  static __device__ __forceinline__ float2 cvt(Raw r) {
    return __half22float2(*reinterpret_cast<const __half2*>(&r));
  }
  static __device__ __forceinline__ void st(void* p, float2 v) {
    *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v.x, v.y);
  }
};

// four consecutive elements (one quarter of a B or C row)
template <typename T> struct Quad;
template <> struct Quad<float> {
  using Raw = float4;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ Raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ void cvt(Raw r, float* o) { o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
};
template <> struct Quad<__nv_bfloat16> {
  using Raw = uint2;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r;
    asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ Raw zero() { return make_uint2(0u, 0u); }
  static __device__ __forceinline__ void cvt(Raw r, float* o) {
    o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
    o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};
template <> struct Quad<__half> {
  using Raw = uint2;
  static __device__ __forceinline__ Raw ld_nc(const void* p) {
    Raw r;
    asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ Raw zero() { return make_uint2(0u, 0u); }
  static __device__ __forceinline__ void cvt(Raw r, float* o) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
};

// private barrier of a direction group (immediate ids: a register id would make ptxas reserve all 16 barriers)
template <int DIR, int NTHR>
__device__ __forceinline__ void group_bar() {
  if (DIR == 0) asm volatile("bar.sync 1, %0;" ::"n"(NTHR) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"n"(NTHR) : "memory");
}

// 32-bit signed byte stride of one processed step; false if it does not fit
inline bool step_stride(int64_t sl_elems, int es, bool reverse, int64_t steps, int32_t* out) {
  const int64_t v = (reverse ? -sl_elems : sl_elems) * es;
  if (v > INT32_MAX / 2 || v < INT32_MIN / 2) return false;
  const int64_t span = (v < 0 ? -v : v) * steps;
  if (span > ((int64_t)1 << 40)) return false;
  *out = (int32_t)v;
  return true;
}

}  // namespace sp
}  // namespace cm
