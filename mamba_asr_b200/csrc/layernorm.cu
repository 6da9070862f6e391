// LayerNorm over the last dimension, forward and backward, for sm_100a.
//
// Scope note: this widens the hot path by one row of SURVEY.md section 8(f) (rank 2, "fusable LN ... next bandwidth
// kernels").  The ConMamba layer applies six LayerNorms of width d_model around each Mamba block
// (reference modules/Conmamba.py:595-621, 638-649); under bf16 autocast torch runs them as cast -> fp32 kernel ->
// separate gamma/beta-gradient kernel, which measured 25 % of the encoder step on B200 (profiles/r01_launches_*).
//
// Forward: one warp per row, the row lives in registers, two-pass mean / variance by warp shuffles, y = (x-mean)*rstd*g+b;
// mean and rstd (fp32 per row) are saved.  Backward: one pass over x and dy; each warp walks a strided set of rows,
// forms dx with two warp reductions per row and accumulates its lanes' columns of dgamma / dbeta in registers; CTAs
// write one partial row each, summed in fixed order by cm_reduce_multi (deterministic, no atomics).
// Roof: HBM; algorithmic bytes per row: forward C*(s_in + s_out), backward C*(s_in + 2*s_dy).
#include <cstdlib>
#include "common.cuh"

namespace cm {

constexpr int kLnWarps = 8;

template <typename T> __device__ __forceinline__ float ln_ld(const T* p) { return Elem<T>::ld(p); }

template <typename Tin, typename Tout, int NPL>
__global__ void __launch_bounds__(32 * kLnWarps) layernorm_fwd_kernel(const Tin* __restrict__ x, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, Tout* __restrict__ y,
                                                                    float* __restrict__ mean, float* __restrict__ rstd,
                                                                    int64_t rows, int C, int64_t x_stride, int64_t y_stride,
                                                                    float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const Tin* xr = x + row * x_stride;
  float v[NPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < C) ? ln_ld<Tin>(xr + c) : 0.f;
    s += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    const float dlt = (c < C) ? v[i] - mu : 0.f;
    q = fmaf(dlt, dlt, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rs = rsqrtf(q / (float)C + eps);
  Tout* yr = y + row * y_stride;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      const float g = gamma ? __ldg(gamma + c) : 1.f, bb = beta ? __ldg(beta + c) : 0.f;
      Elem<Tout>::st(yr + c, fmaf((v[i] - mu) * rs, g, bb));
    }
  }
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

template <typename Tin, typename Tdy, int NPL>
__global__ void __launch_bounds__(32 * kLnWarps) layernorm_bwd_kernel(const Tin* __restrict__ x, const Tdy* __restrict__ dy,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    Tin* __restrict__ dx, float* __restrict__ dgamma_part,
                                                                    float* __restrict__ dbeta_part, int64_t rows, int C,
                                                                    int64_t x_stride, int64_t dy_stride, int64_t dx_stride) {
  __shared__ float red[kLnWarps][32 * NPL + 1];   // reused for dgamma, then dbeta
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[NPL], dg[NPL], db[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    g[i] = (gamma && c < C) ? __ldg(gamma + c) : 1.f;
    dg[i] = 0.f; db[i] = 0.f;
  }
  const float invC = 1.f / (float)C;
  for (int64_t row = (int64_t)blockIdx.x * kLnWarps + warp; row < rows; row += (int64_t)gridDim.x * kLnWarps) {
    const Tin* xr = x + row * x_stride;
    const Tdy* dr = dy + row * dy_stride;
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float xh[NPL], gy[NPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      const float xv = (c < C) ? ln_ld<Tin>(xr + c) : 0.f;
      const float dv = (c < C) ? ln_ld<Tdy>(dr + c) : 0.f;
      xh[i] = (c < C) ? (xv - mu) * rs : 0.f;
      gy[i] = dv * g[i];
      dg[i] = fmaf(dv, xh[i], dg[i]);
      db[i] += dv;
      s1 += gy[i];
      s2 = fmaf(gy[i], xh[i], s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * invC, m2 = s2 * invC;
    Tin* dxr = dx + row * dx_stride;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) Elem<Tin>::st(dxr + c, rs * (gy[i] - m1 - xh[i] * m2));
    }
  }
  // CTA reduce of the per-warp column sums (fixed order), one partial row per CTA
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (pass) __syncthreads();
#pragma unroll
    for (int i = 0; i < NPL; ++i) red[warp][lane + 32 * i] = pass ? db[i] : dg[i];
    __syncthreads();
    float* dst = pass ? dbeta_part : dgamma_part;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < kLnWarps; ++w) a += red[w][c];
      dst[(int64_t)blockIdx.x * C + c] = a;
    }
  }
}

// ---- pair-vectorised variants (even C, even strides): a lane owns column pairs lane + 32 i, every warp works on TWO
// rows at a time, so twice the bytes are in flight per warp and every access is 4 or 8 bytes wide.
template <typename T> struct Ln2;
template <> struct Ln2<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <> struct Ln2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float2 v) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y); }
};
template <> struct Ln2<__half> {
  static __device__ __forceinline__ float2 ld(const __half* p) {
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(p));
    return __half22float2(*reinterpret_cast<const __half2*>(&r));
  }
  static __device__ __forceinline__ void st(__half* p, float2 v) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v.x, v.y); }
};

template <typename Tin, typename Tout, int NPP>
__global__ void __launch_bounds__(32 * kLnWarps) layernorm_fwd2_kernel(const Tin* __restrict__ x, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, Tout* __restrict__ y,
                                                                     float* __restrict__ mean, float* __restrict__ rstd,
                                                                     int64_t rows, int C, int64_t x_stride, int64_t y_stride,
                                                                     float eps, int act) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5)) * 2;
  if (row0 >= rows) return;
  const bool two = row0 + 1 < rows;
  const int np = C >> 1;
  float2 v[2][NPP];
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const Tin* xr = x + (row0 + (two ? r : 0)) * x_stride;
#pragma unroll
    for (int i = 0; i < NPP; ++i) {
      const int pi = lane + 32 * i;
      v[r][i] = (pi < np) ? Ln2<Tin>::ld(xr + 2 * pi) : make_float2(0.f, 0.f);
      s[r] += v[r][i].x + v[r][i].y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
    s[1] += __shfl_xor_sync(0xffffffffu, s[1], o);
  }
  const float invC = 1.f / (float)C;
  float mu[2] = {s[0] * invC, s[1] * invC}, q[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < NPP; ++i) {
      const int pi = lane + 32 * i;
      if (pi < np) {
        const float a = v[r][i].x - mu[r], b = v[r][i].y - mu[r];
        q[r] = fmaf(a, a, fmaf(b, b, q[r]));
      }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    q[0] += __shfl_xor_sync(0xffffffffu, q[0], o);
    q[1] += __shfl_xor_sync(0xffffffffu, q[1], o);
  }
  const float rs[2] = {rsqrtf(q[0] * invC + eps), rsqrtf(q[1] * invC + eps)};
#pragma unroll
  for (int i = 0; i < NPP; ++i) {
    const int pi = lane + 32 * i;
    if (pi < np) {
      const float2 g = gamma ? __ldg(reinterpret_cast<const float2*>(gamma + 2 * pi)) : make_float2(1.f, 1.f);
      const float2 bb = beta ? __ldg(reinterpret_cast<const float2*>(beta + 2 * pi)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 1 && !two) break;
        float2 o = make_float2(fmaf((v[r][i].x - mu[r]) * rs[r], g.x, bb.x), fmaf((v[r][i].y - mu[r]) * rs[r], g.y, bb.y));
        if (act) o = make_float2(gelu_f(o.x), gelu_f(o.y));          // CM_LN_ACT_GELU epilogue (exact erf GELU)
        Ln2<Tout>::st(y + (row0 + r) * y_stride + 2 * pi, o);
      }
    }
  }
  if (lane == 0) {
    mean[row0] = mu[0]; rstd[row0] = rs[0];
    if (two) { mean[row0 + 1] = mu[1]; rstd[row0 + 1] = rs[1]; }
  }
}

template <typename Tin, typename Tdy, int NPP>
__global__ void __launch_bounds__(32 * kLnWarps) layernorm_bwd2_kernel(const Tin* __restrict__ x, const Tdy* __restrict__ dy,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                     Tin* __restrict__ dx, float* __restrict__ dgamma_part,
                                                                     float* __restrict__ dbeta_part, int64_t rows, int C,
                                                                     int64_t x_stride, int64_t dy_stride, int64_t dx_stride,
                                                                     const float* __restrict__ beta, int act) {
  __shared__ float2 red[kLnWarps][32 * NPP + 1];   // reused for dgamma, then dbeta
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int np = C >> 1;
  float2 g[NPP], dg[NPP], db[NPP];
#pragma unroll
  for (int i = 0; i < NPP; ++i) {
    const int pi = lane + 32 * i;
    g[i] = (gamma && pi < np) ? __ldg(reinterpret_cast<const float2*>(gamma + 2 * pi)) : make_float2(1.f, 1.f);
    dg[i] = make_float2(0.f, 0.f); db[i] = make_float2(0.f, 0.f);
  }
  const float invC = 1.f / (float)C;
  const int64_t rstep = (int64_t)gridDim.x * kLnWarps * 2;
  for (int64_t row0 = ((int64_t)blockIdx.x * kLnWarps + warp) * 2; row0 < rows; row0 += rstep) {
    const bool two = row0 + 1 < rows;
    float2 xh[2][NPP], gy[2][NPP];
    float mu[2], rs[2], s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int64_t row = row0 + (two ? r : 0);
      mu[r] = __ldg(mean + row); rs[r] = __ldg(rstd + row);
      const Tin* xr = x + row * x_stride;
      const Tdy* dr = dy + row * dy_stride;
#pragma unroll
      for (int i = 0; i < NPP; ++i) {
        const int pi = lane + 32 * i;
        xh[r][i] = (pi < np) ? Ln2<Tin>::ld(xr + 2 * pi) : make_float2(0.f, 0.f);
        gy[r][i] = (pi < np && (r == 0 || two)) ? Ln2<Tdy>::ld(dr + 2 * pi) : make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < NPP; ++i) {
        const int pi = lane + 32 * i;
        float2 dv = gy[r][i];
        const float2 h = (pi < np) ? make_float2((xh[r][i].x - mu[r]) * rs[r], (xh[r][i].y - mu[r]) * rs[r]) : make_float2(0.f, 0.f);
        if (act && pi < np) {
          // GELU epilogue of the forward: dy arrives for gelu(n), n = h * gamma + beta is recomputed (beta re-read: L1)
          const float2 bt = beta ? __ldg(reinterpret_cast<const float2*>(beta + 2 * pi)) : make_float2(0.f, 0.f);
          dv.x *= gelu_grad_f(fmaf(h.x, g[i].x, bt.x));
          dv.y *= gelu_grad_f(fmaf(h.y, g[i].y, bt.y));
        }
        xh[r][i] = h;
        gy[r][i] = make_float2(dv.x * g[i].x, dv.y * g[i].y);
        dg[i].x = fmaf(dv.x, h.x, dg[i].x); dg[i].y = fmaf(dv.y, h.y, dg[i].y);
        db[i].x += dv.x; db[i].y += dv.y;
        s1[r] += gy[r][i].x + gy[r][i].y;
        s2[r] = fmaf(gy[r][i].x, h.x, fmaf(gy[r][i].y, h.y, s2[r]));
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], o); s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], o);
      s1[1] += __shfl_xor_sync(0xffffffffu, s1[1], o); s2[1] += __shfl_xor_sync(0xffffffffu, s2[1], o);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r == 1 && !two) break;
      const float m1 = s1[r] * invC, m2 = s2[r] * invC;
      Tin* dxr = dx + (row0 + r) * dx_stride;
#pragma unroll
      for (int i = 0; i < NPP; ++i) {
        const int pi = lane + 32 * i;
        if (pi < np)
          Ln2<Tin>::st(dxr + 2 * pi, make_float2(rs[r] * (gy[r][i].x - m1 - xh[r][i].x * m2), rs[r] * (gy[r][i].y - m1 - xh[r][i].y * m2)));
      }
    }
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (pass) __syncthreads();
#pragma unroll
    for (int i = 0; i < NPP; ++i) red[warp][lane + 32 * i] = pass ? db[i] : dg[i];
    __syncthreads();
    float* dst = pass ? dbeta_part : dgamma_part;
    for (int pi = threadIdx.x; pi < np; pi += blockDim.x) {
      float2 a = make_float2(0.f, 0.f);
#pragma unroll
      for (int w = 0; w < kLnWarps; ++w) { a.x += red[w][pi].x; a.y += red[w][pi].y; }
      *reinterpret_cast<float2*>(dst + (int64_t)blockIdx.x * C + 2 * pi) = a;
    }
  }
}

template <typename Tin, typename Tout>
static int ln_fwd_launch(const void* x, const float* g, const float* b, void* y, float* mean, float* rstd, int64_t rows,
                         int C, int64_t xs, int64_t ys, float eps, int act, cudaStream_t st) {
  if ((C & 1) == 0 && (xs & 1) == 0 && (ys & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(y) & 7) == 0 && (!g || (reinterpret_cast<uintptr_t>(g) & 7) == 0) &&
      (!b || (reinterpret_cast<uintptr_t>(b) & 7) == 0)) {
    const unsigned grid2 = (unsigned)((rows + 2 * kLnWarps - 1) / (2 * kLnWarps));
#define LN_FWD2(N) layernorm_fwd2_kernel<Tin, Tout, N><<<grid2, 32 * kLnWarps, 0, st>>>( \
      static_cast<const Tin*>(x), g, b, static_cast<Tout*>(y), mean, rstd, rows, C, xs, ys, eps, act)
    if (C <= 192) LN_FWD2(3); else if (C <= 256) LN_FWD2(4); else if (C <= 512) LN_FWD2(8); else LN_FWD2(16);
#undef LN_FWD2
    CM_LAUNCH_CHECK();
    return 0;
  }
  if (act) return CM_ERR_UNSUPPORTED;          // the activation epilogue exists on the pair-vectorised kernels only
  const unsigned grid = (unsigned)((rows + kLnWarps - 1) / kLnWarps);
#define LN_FWD(N) layernorm_fwd_kernel<Tin, Tout, N><<<grid, 32 * kLnWarps, 0, st>>>( \
      static_cast<const Tin*>(x), g, b, static_cast<Tout*>(y), mean, rstd, rows, C, xs, ys, eps)
  if (C <= 160) LN_FWD(5); else if (C <= 256) LN_FWD(8); else if (C <= 512) LN_FWD(16); else LN_FWD(32);
#undef LN_FWD
  CM_LAUNCH_CHECK();
  return 0;
}

template <typename Tin, typename Tdy>
static int ln_bwd_launch(const void* x, const void* dy, const float* g, const float* mean, const float* rstd, void* dx,
                         float* dgp, float* dbp, int64_t rows, int C, int64_t xs, int64_t dys, int64_t dxs, int nblk,
                         const float* beta, int act, cudaStream_t st) {
  if ((C & 1) == 0 && (xs & 1) == 0 && (dys & 1) == 0 && (dxs & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & 7) == 0 && (reinterpret_cast<uintptr_t>(dx) & 7) == 0 &&
      (!g || (reinterpret_cast<uintptr_t>(g) & 7) == 0) && (reinterpret_cast<uintptr_t>(dgp) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(dbp) & 7) == 0) {
#define LN_BWD2(N) layernorm_bwd2_kernel<Tin, Tdy, N><<<nblk, 32 * kLnWarps, 0, st>>>( \
      static_cast<const Tin*>(x), static_cast<const Tdy*>(dy), g, mean, rstd, static_cast<Tin*>(dx), dgp, dbp, rows, C, xs, dys, dxs, beta, act)
    if (C <= 192) LN_BWD2(3); else if (C <= 256) LN_BWD2(4); else if (C <= 512) LN_BWD2(8); else LN_BWD2(16);
#undef LN_BWD2
    CM_LAUNCH_CHECK();
    return 0;
  }
  if (act) return CM_ERR_UNSUPPORTED;
#define LN_BWD(N) layernorm_bwd_kernel<Tin, Tdy, N><<<nblk, 32 * kLnWarps, 0, st>>>( \
      static_cast<const Tin*>(x), static_cast<const Tdy*>(dy), g, mean, rstd, static_cast<Tin*>(dx), dgp, dbp, rows, C, xs, dys, dxs)
  if (C <= 160) LN_BWD(5); else if (C <= 256) LN_BWD(8); else if (C <= 512) LN_BWD(16); else LN_BWD(32);
#undef LN_BWD
  CM_LAUNCH_CHECK();
  return 0;
}

}  // namespace cm

extern "C" int cm_layernorm_num_part(int64_t rows) {
  const int64_t need = (rows + cm::kLnWarps - 1) / cm::kLnWarps;
#ifndef CM_LN_CTAS_PER_SM
#define CM_LN_CTAS_PER_SM 4
#endif
  const int64_t cap = 148 * CM_LN_CTAS_PER_SM;      // 4 CTAs of 8 warps per SM
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

extern "C" int cm_layernorm_fwd(const cm_layernorm_args* a, void* stream) {
  if (!a || !a->x || !a->y || !a->mean || !a->rstd || a->rows <= 0 || a->cols <= 0) return CM_ERR_BAD_ARG;
  if (a->cols > 1024) return CM_ERR_UNSUPPORTED;
  if (a->act != 0 && a->act != CM_LN_OUT_GELU) return CM_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FWD(TI, TO) return cm::ln_fwd_launch<TI, TO>(a->x, a->gamma, a->beta, a->y, a->mean, a->rstd, a->rows, a->cols, \
                                                     a->x_stride, a->y_stride, a->eps, a->act, st)
  if (a->x_dtype == CM_F32 && a->y_dtype == CM_F32) FWD(float, float);
  if (a->x_dtype == CM_BF16 && a->y_dtype == CM_F32) FWD(__nv_bfloat16, float);
  if (a->x_dtype == CM_BF16 && a->y_dtype == CM_BF16) FWD(__nv_bfloat16, __nv_bfloat16);
  if (a->x_dtype == CM_F16 && a->y_dtype == CM_F32) FWD(__half, float);
  if (a->x_dtype == CM_F16 && a->y_dtype == CM_F16) FWD(__half, __half);
  if (a->x_dtype == CM_F32 && a->y_dtype == CM_BF16) FWD(float, __nv_bfloat16);
#undef FWD
  return CM_ERR_UNSUPPORTED;
}

namespace cm { int ln_bwd_routed(const cm_add_ln_args& a, int nblk, bool act, cudaStream_t st); }   // fused_ln.cu

static bool ln_route_on() {
  static const bool off = getenv("CM_LN_NO_ROUTE") != nullptr;      // A/B switch: the pair kernels of this file
  return !off;
}

// v17: the partial rows cm_layernorm_bwd writes when the caller passes them as args.n_part - the grid of the kernel that
// will run for this row width (the quad / staged kernels of fused_ln.cu for cols % 4 == 0)
extern "C" int cm_layernorm_num_part2(int64_t rows, int32_t cols) {
  if (ln_route_on() && cols > 0 && (cols & 3) == 0) return cm_add_ln_num_part(rows, cols);
  return cm_layernorm_num_part(rows);
}

extern "C" int cm_layernorm_bwd(const cm_layernorm_args* a, void* stream) {
  if (!a || !a->x || !a->dy || !a->dx || !a->mean || !a->rstd || !a->dgamma_part || !a->dbeta_part) return CM_ERR_BAD_ARG;
  if (a->rows <= 0 || a->cols <= 0 || a->n_part < 0) return CM_ERR_BAD_ARG;
  if (a->cols > 1024) return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = a->n_part > 0 ? a->n_part : cm_layernorm_num_part(a->rows);
  if (a->act != 0 && a->act != CM_LN_OUT_GELU) return CM_ERR_BAD_ARG;
  if (a->act != 0 && a->beta != nullptr && (reinterpret_cast<uintptr_t>(a->beta) & 7) != 0) return CM_ERR_UNSUPPORTED;
  if (a->n_part > 0 && ln_route_on() && (a->cols & 3) == 0) {
    // LayerNorm backward = the residual-add LayerNorm backward without ds / db: 16-byte accesses, rows staged by bulk copies
    cm_add_ln_args q{};
    q.rows = a->rows; q.cols = a->cols;
    q.a_dtype = a->x_dtype; q.b_dtype = CM_BF16; q.y_dtype = a->y_dtype;
    q.alpha = 1.0f; q.p_drop = 0.0f;
    q.s = const_cast<void*>(a->x); q.s_stride = a->x_stride;
    q.dy = a->dy; q.dy_stride = a->dy_stride;
    q.da = a->dx; q.da_stride = a->dx_stride;
    q.gamma = a->gamma; q.beta = a->beta; q.mean = a->mean; q.rstd = a->rstd;
    q.dgamma_part = a->dgamma_part; q.dbeta_part = a->dbeta_part;
    const int rc = cm::ln_bwd_routed(q, nblk, a->act == CM_LN_OUT_GELU, st);
    if (rc != CM_ERR_UNSUPPORTED) return rc;
  }
#define BWD(TI, TD) return cm::ln_bwd_launch<TI, TD>(a->x, a->dy, a->gamma, a->mean, a->rstd, a->dx, a->dgamma_part, \
                                                     a->dbeta_part, a->rows, a->cols, a->x_stride, a->dy_stride, a->dx_stride, nblk, \
                                                     a->beta, a->act, st)
  if (a->x_dtype == CM_F32 && a->y_dtype == CM_F32) BWD(float, float);
  if (a->x_dtype == CM_BF16 && a->y_dtype == CM_F32) BWD(__nv_bfloat16, float);
  if (a->x_dtype == CM_BF16 && a->y_dtype == CM_BF16) BWD(__nv_bfloat16, __nv_bfloat16);
  if (a->x_dtype == CM_F16 && a->y_dtype == CM_F32) BWD(__half, float);
  if (a->x_dtype == CM_F16 && a->y_dtype == CM_F16) BWD(__half, __half);
  if (a->x_dtype == CM_F32 && a->y_dtype == CM_BF16) BWD(float, __nv_bfloat16);
#undef BWD
  return CM_ERR_UNSUPPORTED;
}
