// Shared device helpers for the ConMamba sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conmamba_b200.h"

namespace cm {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// 16-bit global loads that land zero-extended in a full 32-bit register.  The C++ route (unsigned short -> uint32_t)
// makes the compiler mask the value right after the LDG (LOP3 0xffff), which stalls the warp on every load and
// defeats the register prefetch; PTX allows a destination wider than the access type.
__device__ __forceinline__ uint32_t ld16_nc(const void* p) {
  uint32_t v;
  asm("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld16_cg(const void* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=r"(v) : "l"(p));   // coherent at L2: reads the partner warp's stash
  return v;
}

// ---- element I/O ---------------------------------------------------------------------------------
template <typename T>
struct Elem;
// ld_raw / cvt split a load from its conversion: kernels batch all ld_raw of a tile into registers first and
// convert at the point of use, so that no ALU instruction sits between consecutive loads (a shift right after
// each LDG serialises the loads on the scoreboard - measured: 6 long-scoreboard stalls per issued instruction).
template <>
struct Elem<float> {
  using Raw = float;
  static __device__ __forceinline__ Raw ld_raw(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ Raw ld_raw_cg(const float* p) { return __ldcg(p); }
  static __device__ __forceinline__ float cvt(Raw r) { return r; }
  static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float round(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
  using Raw = uint32_t;   // zero-extended 16 bits in a full register (a 16-bit Raw makes ptxas pack pairs with PRMT)
  static __device__ __forceinline__ Raw ld_raw(const __nv_bfloat16* p) { return ld16_nc(p); }
  static __device__ __forceinline__ Raw ld_raw_cg(const __nv_bfloat16* p) { return ld16_cg(p); }
  static __device__ __forceinline__ float cvt(Raw r) { return __uint_as_float(r << 16); }
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) {
    unsigned short r = __ldg(reinterpret_cast<const unsigned short*>(p));
    return __uint_as_float(static_cast<uint32_t>(r) << 16);
  }
  static __device__ __forceinline__ float ld_cg(const __nv_bfloat16* p) {
    unsigned short r = __ldcg(reinterpret_cast<const unsigned short*>(p));
    return __uint_as_float(static_cast<uint32_t>(r) << 16);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};
template <>
struct Elem<__half> {
  using Raw = uint32_t;
  static __device__ __forceinline__ Raw ld_raw(const __half* p) { return ld16_nc(p); }
  static __device__ __forceinline__ Raw ld_raw_cg(const __half* p) { return ld16_cg(p); }
  static __device__ __forceinline__ float cvt(Raw r) { return __half2float(__ushort_as_half(static_cast<unsigned short>(r))); }
  static __device__ __forceinline__ float ld(const __half* p) {
    unsigned short r = __ldg(reinterpret_cast<const unsigned short*>(p));
    return __half2float(__ushort_as_half(r));
  }
  static __device__ __forceinline__ float ld_cg(const __half* p) {
    unsigned short r = __ldcg(reinterpret_cast<const unsigned short*>(p));
    return __half2float(__ushort_as_half(r));
  }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
  static __device__ __forceinline__ float round(float v) { return __half2float(__float2half_rn(v)); }
};

// ---- MUFU wrappers (single SASS instruction each) ------------------------------------------------
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sigmoid(x) = 1 / (1 + 2^(-x*log2e)); saturates cleanly (ex2 -> inf -> rcp -> 0)
__device__ __forceinline__ float sigmoidf_fast(float x) { return rcp(1.0f + ex2(-x * kLog2e)); }

// softplus with torch semantics (identity above threshold 20).  PRECISE keeps the relative error of small
// results at fp32 level (series for exp(x) < 1/4) - needed for the rtol 1e-4 fp32 contract; the 16-bit I/O
// instantiations use the two-MUFU form.  *sig receives d softplus / dx = sigmoid(x).
template <bool PRECISE>
__device__ __forceinline__ float softplus_fwd(float x) {
  // branch-free: both forms are evaluated and selected
  const float e = ex2(x * kLog2e);
  float r = kLn2 * lg2(1.0f + e);
  if (PRECISE) {
    // log1p(e) = e - e^2/2 + e^3/3 - ... ; used for e < 1/4 (11 terms: truncation < 1e-8 relative), where the
    // lg2.approx absolute error would dominate a small result
    float p = -1.0f / 12.0f;
    p = fmaf(p, e, 1.0f / 11.0f);
    p = fmaf(p, e, -1.0f / 10.0f);
    p = fmaf(p, e, 1.0f / 9.0f);
    p = fmaf(p, e, -1.0f / 8.0f);
    p = fmaf(p, e, 1.0f / 7.0f);
    p = fmaf(p, e, -1.0f / 6.0f);
    p = fmaf(p, e, 1.0f / 5.0f);
    p = fmaf(p, e, -1.0f / 4.0f);
    p = fmaf(p, e, 1.0f / 3.0f);
    p = fmaf(p, e, -1.0f / 2.0f);
    p = fmaf(p, e, 1.0f);
    r = (e < 0.25f) ? p * e : r;
  }
  return (x > 20.0f) ? x : r;
}
__device__ __forceinline__ float softplus_grad(float x) { return x > 20.0f ? 1.0f : sigmoidf_fast(x); }

// sigmoid for 16-bit outputs: one MUFU (tanh.approx, |rel err| ~2^-11) instead of ex2 + rcp
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool PRECISE>
__device__ __forceinline__ float sigmoid_sel(float x) {
  if (PRECISE) return sigmoidf_fast(x);
  return fmaf(tanh_approx(0.5f * x), 0.5f, 0.5f);
}

// ---- exact (erf) GELU ------------------------------------------------------------------------------
constexpr float kInvSqrt2 = 0.7071067811865476f;
constexpr float kInvSqrt2Pi = 0.3989422804014327f;
// Phi(x) = 0.5 * (1 + erf(x / sqrt 2)) and phi(x) = exp(-x^2/2) / sqrt(2 pi) from ONE ex2: erf by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, the fp32 resolution of 1 + erf), 1 - erf(|y|) = poly(t) * exp(-y^2), t = 1 / (1 + p |y|).
// libdevice's erff costs ~25 instructions per element and made the fused kernel issue-bound (82 % issue-slot utilisation
// in profiles/r01_fused_addln_gelu_cfg3_ncu.txt); this form is 2 MUFU + 10 FMA-pipe instructions and shares the
// exponential with the derivative.
__device__ __forceinline__ void gelu_parts(float x, float* Phi, float* phi) {
  // constants folded (y = |x| / sqrt 2 never formed, the 0.5 inside the coefficients): 4 FMUL + 5 FFMA + 2 MUFU
  const float e = ex2(x * x * (-0.5f * kLog2e));        // exp(-x^2 / 2)
  const float t = rcp(fmaf(0.3275911f * kInvSqrt2, fabsf(x), 1.0f));
  float p = 0.5f * 1.061405429f;
  p = fmaf(p, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float q = p * t * e;                            // 0.5 * erfc(|x| / sqrt 2) = Phi(-|x|)
  *Phi = x >= 0.f ? 1.0f - q : q;
  *phi = kInvSqrt2Pi * e;
}
__device__ __forceinline__ float gelu_f(float x) {
  float P, d;
  gelu_parts(x, &P, &d);
  return x * P;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float P, d;
  gelu_parts(x, &P, &d);
  return fmaf(x, d, P);
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) ---------------
#ifdef CM_SCALAR_F32X2   // A/B experiment: the packed helpers as two scalar instructions each
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#else
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
#endif

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

inline int dtype_ok(int32_t dt) { return dt == CM_F32 || dt == CM_BF16 || dt == CM_F16; }

// time index of processed step s
__device__ __forceinline__ int step_to_time(int s, int L, bool reverse) { return reverse ? (L - 1 - s) : s; }

}  // namespace cm

// host-side helpers shared by the launchers -----------------------------------------------------------
#define CM_LAUNCH_CHECK()                                \
  do {                                                   \
    cudaError_t e__ = cudaGetLastError();                \
    if (e__ != cudaSuccess) return static_cast<int>(e__); \
  } while (0)

// split point of a bidirectional row: direction 0 stashes [0, M), direction 1 stashes [M, L)
static inline __host__ __device__ int cm_mid(int L) { return (L + 1) / 2; }
// Length (in processed steps) of the first range of a direction.  In a bidirectional launch the ascending
// direction first covers times [0, M) and the descending one [M, L); after a CTA barrier each finishes the
// other half, combining with what its partner stashed there.
static inline __host__ __device__ int cm_first_range(int L, int ndir, int reverse) {
  if (ndir == 1) return L;
  return reverse ? (L - cm_mid(L)) : cm_mid(L);
}
static inline __host__ __device__ int cm_ceil_div(int a, int b) { return (a + b - 1) / b; }
