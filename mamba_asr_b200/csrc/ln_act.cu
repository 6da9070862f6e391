// LayerNorm + activation in one pass (cm_ln_act_fwd / cm_ln_act_bwd), sm_100a: the wide rows of the CNN front-end
// (LeakyReLU, conv bias folded in) and the LayerNorm -> GELU of the ConMamba convolution module.
//
// SURVEY.md section 8(f) rank 2 ("SpeechBrain-free layer shell"): the reference's ConvolutionFrontEnd normalises every
// conv block's output over (freq, channel) - LayerNorm([F', C]) = rows of 2560 and 640 elements at the BASELINE shapes -
// and applies LeakyReLU (hparams/CTC/conmamba_large.yaml:187-199).  Under bf16 autocast torch runs this as
// cast -> fp32 LayerNorm -> fp32 LeakyReLU -> cast, plus three kernels in backward, over the largest activations of the
// whole step (64 x 1001 x 2560): 3.5 ms of the 54.5 ms ConMamba-large step.  Here: one pass forward (read x, write y in
// the same dtype, statistics in fp32), one pass backward (read x and dy, recompute the pre-activation sign, write dx,
// per-CTA dgamma / dbeta partial rows summed in fixed order by cm_reduce_multi - no atomics).
//
// Mapping: a group of G = 32 / 64 / 128 threads owns a row, a thread kNQ = 5 quads (4 consecutive elements) at a
// stride of G quads, so a warp's accesses are 256-byte (bf16) / 512-byte (fp32) contiguous runs and the row stays in
// registers between the statistics and the output.  CTAs are persistent over row blocks (grid-stride).  HBM-bound:
// algorithmic bytes per element 2s forward, 3s backward (s = bytes per element).
#include "ln_act_common.cuh"

namespace cm {
namespace lna {

// bias added to x before the statistics (the conv bias of the front-end blocks): period pbn columns, pbn % 4 == 0
__device__ __forceinline__ float4 pre_bias4(const float* pb, int pbn, int q, bool ok) {
  return (pb != nullptr && ok) ? __ldg(reinterpret_cast<const float4*>(pb + (4 * q) % pbn)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <typename T, int G, int ACT>
__global__ void __launch_bounds__(kThreads) ln_act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ pre_bias, int pbn,
                                                              float* __restrict__ mean, float* __restrict__ rstd,
                                                              int64_t rows, int cols, float eps, float slope) {
  constexpr int RPC = kThreads / G;
  constexpr int ES = (int)sizeof(T);
  __shared__ float2 red[2][kThreads / 32];
  const int grp = threadIdx.x / G, gl = threadIdx.x % G;
  const int nq = cols >> 2;
  const float inv_n = 1.0f / (float)cols;
  float4 gm[kNQ], bt[kNQ];
#pragma unroll
  for (int i = 0; i < kNQ; ++i) {
    const int q = gl + i * G;
    gm[i] = q < nq ? __ldg(reinterpret_cast<const float4*>(gamma) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[i] = q < nq ? __ldg(reinterpret_cast<const float4*>(beta) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 pbv[kNQ];
#pragma unroll
  for (int i = 0; i < kNQ; ++i) pbv[i] = pre_bias4(pre_bias, pbn, gl + i * G, gl + i * G < nq);
  const int64_t nblk = (rows + RPC - 1) / RPC;
  int par = 0;
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t row = blk * RPC + grp;
    const bool rv = row < rows;
    const char* px = reinterpret_cast<const char*>(x) + row * (int64_t)cols * ES;
    typename Quad<T>::Raw raw[kNQ];
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const int q = gl + i * G;
      raw[i] = (rv && q < nq) ? Quad<T>::ld_nc(px + (int64_t)q * 4 * ES) : Quad<T>::zero();
    }
    float v[kNQ][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      Quad<T>::cvt(raw[i], v[i]);
      const float4 pb = pbv[i];
      v[i][0] += pb.x; v[i][1] += pb.y; v[i][2] += pb.z; v[i][3] += pb.w;
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    const float mu = group_sum<G>(make_float2(s, 0.f), red, par).x * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const bool qv = gl + i * G < nq;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[i][e] = qv ? v[i][e] - mu : 0.f;
        sq = fmaf(v[i][e], v[i][e], sq);
      }
    }
    const float var = group_sum<G>(make_float2(sq, 0.f), red, par).x * inv_n;
    const float rs = rsqrtf(var + eps);
    if (rv) {
      char* py = reinterpret_cast<char*>(y) + row * (int64_t)cols * ES;
#pragma unroll
      for (int i = 0; i < kNQ; ++i) {
        const int q = gl + i * G;
        if (q < nq) {
          const float g4[4] = {gm[i].x, gm[i].y, gm[i].z, gm[i].w}, b4[4] = {bt[i].x, bt[i].y, bt[i].z, bt[i].w};
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            o[e] = act_fwd<ACT>(fmaf(v[i][e] * rs, g4[e], b4[e]), slope);
          }
          st4<T>(py + (int64_t)q * 4 * ES, o);
        }
      }
      if (gl == 0) {
        mean[row] = mu;
        rstd[row] = rs;
      }
    }
  }
}

template <typename T, int G, int ACT>
__global__ void __launch_bounds__(kThreads, 4) ln_act_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                 T* __restrict__ dx, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 const float* __restrict__ pre_bias, int pbn,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, float* __restrict__ dg_part,
                                                                 float* __restrict__ db_part, int64_t rows, int cols,
                                                                 float slope) {
  constexpr int RPC = kThreads / G;
  constexpr int ES = (int)sizeof(T);
  __shared__ float2 red[2][kThreads / 32];
  __shared__ float4 comb[RPC > 1 ? kThreads * kNQ : 1];
  const int grp = threadIdx.x / G, gl = threadIdx.x % G;
  const int nq = cols >> 2;
  const float inv_n = 1.0f / (float)cols;
  float dg[kNQ][4], db[kNQ][4];
#pragma unroll
  for (int i = 0; i < kNQ; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dg[i][e] = db[i][e] = 0.f;
  const int64_t nblk = (rows + RPC - 1) / RPC;
  int par = 0;
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t row = blk * RPC + grp;
    const bool rv = row < rows;
    const int64_t off = row * (int64_t)cols * ES;
    const char* px = reinterpret_cast<const char*>(x) + off;
    const char* pdy = reinterpret_cast<const char*>(dy) + off;
    typename Quad<T>::Raw rx[kNQ], rdy[kNQ];
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const int q = gl + i * G;
      const bool ok = rv && q < nq;
      rx[i] = ok ? Quad<T>::ld_nc(px + (int64_t)q * 4 * ES) : Quad<T>::zero();
      rdy[i] = ok ? Quad<T>::ld_nc(pdy + (int64_t)q * 4 * ES) : Quad<T>::zero();
    }
    const float mu = rv ? __ldg(mean + row) : 0.f;
    const float rs = rv ? __ldg(rstd + row) : 0.f;
    float xh[kNQ][4], dxh[kNQ][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const int q = gl + i * G;
      const bool qv = q < nq;
      const float4 g4v = qv ? __ldg(reinterpret_cast<const float4*>(gamma) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 b4v = qv ? __ldg(reinterpret_cast<const float4*>(beta) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float g4[4] = {g4v.x, g4v.y, g4v.z, g4v.w}, b4[4] = {b4v.x, b4v.y, b4v.z, b4v.w};
      float xv[4], dv[4];
      Quad<T>::cvt(rx[i], xv);
      Quad<T>::cvt(rdy[i], dv);
      const float4 pb = pre_bias4(pre_bias, pbn, q, rv && qv);
      xv[0] += pb.x; xv[1] += pb.y; xv[2] += pb.z; xv[3] += pb.w;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h = (xv[e] - mu) * rs;
        const float pre = fmaf(h, g4[e], b4[e]);
        const float gg = act_bwd<ACT>(pre, dv[e], slope);     // zero for padding quads / rows (dy loaded as 0)
        dg[i][e] = fmaf(gg, h, dg[i][e]);
        db[i][e] += gg;
        const float d = gg * g4[e];
        s1 += d;
        s2 = fmaf(d, h, s2);
        xh[i][e] = h;
        dxh[i][e] = d;
      }
    }
    const float2 ss = group_sum<G>(make_float2(s1, s2), red, par);
    const float m1 = ss.x * inv_n, m2 = ss.y * inv_n;
    if (rv) {
      char* pdx = reinterpret_cast<char*>(dx) + off;
#pragma unroll
      for (int i = 0; i < kNQ; ++i) {
        const int q = gl + i * G;
        if (q < nq) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = rs * (dxh[i][e] - m1 - xh[i][e] * m2);
          st4<T>(pdx + (int64_t)q * 4 * ES, o);
        }
      }
    }
  }
  // one partial row per CTA: the RPC row groups of a CTA own the same columns and are summed here in fixed order
  float* const outs[2] = {dg_part + (int64_t)blockIdx.x * cols, db_part + (int64_t)blockIdx.x * cols};
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    float(*acc)[4] = which == 0 ? dg : db;
    if (RPC > 1) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kNQ; ++i) comb[(grp * kNQ + i) * G + gl] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      __syncthreads();
      if (grp == 0) {
#pragma unroll
        for (int i = 0; i < kNQ; ++i) {
          const int q = gl + i * G;
          float4 s = comb[i * G + gl];
#pragma unroll
          for (int r = 1; r < RPC; ++r) {
            const float4 t = comb[(r * kNQ + i) * G + gl];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
          }
          if (q < nq) reinterpret_cast<float4*>(outs[which])[q] = s;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kNQ; ++i) {
        const int q = gl + i * G;
        if (q < nq) reinterpret_cast<float4*>(outs[which])[q] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    }
  }
}

static int group_size(int cols) { return cols <= 32 * kNQ * 4 ? 32 : cols <= 64 * kNQ * 4 ? 64 : 128; }

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      v = 148;
    n = v;
  }
  return n;
}

static int64_t row_blocks(int64_t rows, int cols) { const int rpc = kThreads / group_size(cols); return (rows + rpc - 1) / rpc; }

static int bwd_grid(int64_t rows, int cols) {
  const int64_t nblk = row_blocks(rows, cols), cap = (int64_t)sm_count() * 4;   // one wave at 4 CTAs / SM
  return (int)(nblk < cap ? (nblk < 1 ? 1 : nblk) : cap);
}

static bool args_ok(const cm_ln_act_args* a) {
  return a && a->x && a->gamma && a->beta && a->mean && a->rstd && a->rows > 0 && a->cols > 0 &&
         (a->act == CM_LN_ACT_LEAKY_RELU || a->act == CM_LN_ACT_GELU);
}
// pre-norm bias: absent, or a period that is a multiple of 4 dividing cols, 16-byte aligned
static bool pre_bias_ok(const cm_ln_act_args* a) {
  if (a->pre_bias == nullptr) return true;
  return a->pre_bias_n > 0 && (a->pre_bias_n & 3) == 0 && a->cols % a->pre_bias_n == 0 &&
         (reinterpret_cast<uintptr_t>(a->pre_bias) & 15) == 0;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// persistent grid = one wave: SMs x resident CTAs of this instantiation (ncu on the first version: 1184 CTAs at 5 resident
// per SM = 1.6 waves, profiles/r01_ln_act_fwd_cfg3_ncu.txt)
template <typename KernelT>
static int one_wave(KernelT kern, int* cached) {
  if (*cached == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0) != cudaSuccess || per_sm <= 0) {
      (void)cudaGetLastError();
      per_sm = 4;
    }
    *cached = per_sm * sm_count();
  }
  return *cached;
}

template <typename T, int G, int ACT>
static int fwd_launch(const cm_ln_act_args* a, cudaStream_t st) {
  static int wave = 0;   // idempotent; a benign race computes it twice
  auto kern = ln_act_fwd_kernel<T, G, ACT>;
  const int64_t nblk = row_blocks(a->rows, a->cols), cap = one_wave(kern, &wave);
  const int grid = (int)(nblk < cap ? nblk : cap);
  kern<<<grid, kThreads, 0, st>>>(static_cast<const T*>(a->x), static_cast<T*>(a->y), a->gamma, a->beta, a->pre_bias,
                                  a->pre_bias_n, a->mean, a->rstd, a->rows, a->cols, a->eps, a->slope);
  CM_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int fwd_t(const cm_ln_act_args* a, cudaStream_t st) {
#define CM_LNA_FWD_G(G)                                                                        \
  return a->act == CM_LN_ACT_GELU ? fwd_launch<T, G, CM_LN_ACT_GELU>(a, st)                    \
                                  : fwd_launch<T, G, CM_LN_ACT_LEAKY_RELU>(a, st)
  switch (group_size(a->cols)) {
    case 32: CM_LNA_FWD_G(32);
    case 64: CM_LNA_FWD_G(64);
    default: CM_LNA_FWD_G(128);
  }
#undef CM_LNA_FWD_G
}

template <typename T>
static int bwd_t(const cm_ln_act_args* a, cudaStream_t st) {
  const int grid = bwd_grid(a->rows, a->cols);
#define CM_LNA_BWD(G, ACT)                                                                                                  \
  ln_act_bwd_kernel<T, G, ACT><<<grid, kThreads, 0, st>>>(static_cast<const T*>(a->x), static_cast<const T*>(a->dy),          \
                                                          static_cast<T*>(a->dx), a->gamma, a->beta, a->pre_bias,            \
                                                          a->pre_bias_n, a->mean, a->rstd, a->dgamma_part, a->dbeta_part,   \
                                                          a->rows, a->cols, a->slope)
#define CM_LNA_BWD_G(G)                                                         \
  do {                                                                          \
    if (a->act == CM_LN_ACT_GELU) CM_LNA_BWD(G, CM_LN_ACT_GELU);                \
    else CM_LNA_BWD(G, CM_LN_ACT_LEAKY_RELU);                                   \
  } while (0)
  switch (group_size(a->cols)) {
    case 32: CM_LNA_BWD_G(32); break;
    case 64: CM_LNA_BWD_G(64); break;
    default: CM_LNA_BWD_G(128); break;
  }
#undef CM_LNA_BWD_G
#undef CM_LNA_BWD
  CM_LAUNCH_CHECK();
  return 0;
}

}  // namespace lna
}  // namespace cm

extern "C" int cm_ln_act_num_part(int64_t rows, int32_t cols) {
  if (rows <= 0 || cols <= 0) return 1;
  return cm::lna::bwd_grid(rows, cols);
}

extern "C" int cm_ln_act_fwd(const cm_ln_act_args* a, void* stream) {
  if (!cm::lna::args_ok(a) || !a->y) return CM_ERR_BAD_ARG;
  if (a->cols > cm::lna::kMaxCols || (a->cols & 3) != 0 || !cm::dtype_ok(a->dtype) || !cm::lna::pre_bias_ok(a)) return CM_ERR_UNSUPPORTED;
  if (!cm::lna::aligned16(a->x) || !cm::lna::aligned16(a->y) || !cm::lna::aligned16(a->gamma) || !cm::lna::aligned16(a->beta))
    return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case CM_F32: return cm::lna::fwd_t<float>(a, st);
    case CM_BF16: return cm::lna::fwd_t<__nv_bfloat16>(a, st);
    default: return cm::lna::fwd_t<__half>(a, st);
  }
}

extern "C" int cm_ln_act_bwd(const cm_ln_act_args* a, void* stream) {
  if (!cm::lna::args_ok(a) || !a->dy || !a->dx || !a->dgamma_part || !a->dbeta_part) return CM_ERR_BAD_ARG;
  if (a->cols > cm::lna::kMaxCols || (a->cols & 3) != 0 || !cm::dtype_ok(a->dtype) || !cm::lna::pre_bias_ok(a)) return CM_ERR_UNSUPPORTED;
  if (!cm::lna::aligned16(a->x) || !cm::lna::aligned16(a->dy) || !cm::lna::aligned16(a->dx) || !cm::lna::aligned16(a->gamma) ||
      !cm::lna::aligned16(a->beta) || !cm::lna::aligned16(a->dgamma_part) || !cm::lna::aligned16(a->dbeta_part))
    return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->dtype) {
    case CM_F32: return cm::lna::bwd_t<float>(a, st);
    case CM_BF16: return cm::lna::bwd_t<__nv_bfloat16>(a, st);
    default: return cm::lna::bwd_t<__half>(a, st);
  }
}
