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

template <typename Tin, typename Tout>
static int ln_fwd_launch(const void* x, const float* g, const float* b, void* y, float* mean, float* rstd, int64_t rows,
                         int C, int64_t xs, int64_t ys, float eps, cudaStream_t st) {
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
                         cudaStream_t st) {
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
  const int64_t cap = 148 * 4;                      // 4 CTAs of 8 warps per SM
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

extern "C" int cm_layernorm_fwd(const cm_layernorm_args* a, void* stream) {
  if (!a || !a->x || !a->y || !a->mean || !a->rstd || a->rows <= 0 || a->cols <= 0) return CM_ERR_BAD_ARG;
  if (a->cols > 1024) return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FWD(TI, TO) return cm::ln_fwd_launch<TI, TO>(a->x, a->gamma, a->beta, a->y, a->mean, a->rstd, a->rows, a->cols, \
                                                     a->x_stride, a->y_stride, a->eps, st)
  if (a->x_dtype == CM_F32 && a->y_dtype == CM_F32) FWD(float, float);
  if (a->x_dtype == CM_BF16 && a->y_dtype == CM_F32) FWD(__nv_bfloat16, float);
  if (a->x_dtype == CM_BF16 && a->y_dtype == CM_BF16) FWD(__nv_bfloat16, __nv_bfloat16);
  if (a->x_dtype == CM_F16 && a->y_dtype == CM_F32) FWD(__half, float);
  if (a->x_dtype == CM_F16 && a->y_dtype == CM_F16) FWD(__half, __half);
  if (a->x_dtype == CM_F32 && a->y_dtype == CM_BF16) FWD(float, __nv_bfloat16);
#undef FWD
  return CM_ERR_UNSUPPORTED;
}

extern "C" int cm_layernorm_bwd(const cm_layernorm_args* a, void* stream) {
  if (!a || !a->x || !a->dy || !a->dx || !a->mean || !a->rstd || !a->dgamma_part || !a->dbeta_part) return CM_ERR_BAD_ARG;
  if (a->rows <= 0 || a->cols <= 0) return CM_ERR_BAD_ARG;
  if (a->cols > 1024) return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = cm_layernorm_num_part(a->rows);
#define BWD(TI, TD) return cm::ln_bwd_launch<TI, TD>(a->x, a->dy, a->gamma, a->mean, a->rstd, a->dx, a->dgamma_part, \
                                                     a->dbeta_part, a->rows, a->cols, a->x_stride, a->dy_stride, a->dx_stride, nblk, st)
  if (a->x_dtype == CM_F32 && a->y_dtype == CM_F32) BWD(float, float);
  if (a->x_dtype == CM_BF16 && a->y_dtype == CM_F32) BWD(__nv_bfloat16, float);
  if (a->x_dtype == CM_BF16 && a->y_dtype == CM_BF16) BWD(__nv_bfloat16, __nv_bfloat16);
  if (a->x_dtype == CM_F16 && a->y_dtype == CM_F32) BWD(__half, float);
  if (a->x_dtype == CM_F16 && a->y_dtype == CM_F16) BWD(__half, __half);
  if (a->x_dtype == CM_F32 && a->y_dtype == CM_BF16) BWD(float, __nv_bfloat16);
#undef BWD
  return CM_ERR_UNSUPPORTED;
}
