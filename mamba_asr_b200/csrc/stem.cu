// First block of the CNN front-end in one kernel each way (cm_stem_fwd / cm_stem_bwd), sm_100a:
//     y = LeakyReLU( LayerNorm_[F', C]( Conv2d(1 -> C, 3 x 3, stride 2, zero padding 1)(feats) + bias ) )
//
// SURVEY.md section 8(f) rank 2 ("SpeechBrain-free layer shell"): ConvolutionFrontEnd block 1 of
// hparams/CTC/conmamba_large.yaml:187-199 (reached from train_CTC.py:288).  With ONE input channel the conv is 9 FMAs per
// output and its output - (64, 1001, 40, 64) at the ConMamba-large shape, 328 MB in bf16 - is the largest activation of the
// step.  cuDNN pads the single channel to 8 in a conversion kernel (164 MB written), runs an implicit GEMM, and the LayerNorm
// kernel then reads the 328 MB back: 0.26 + 0.34 + 0.15 ms forward, 0.39 (LN) + 0.26 (pad) + 0.16 (wgrad) ms backward.
// Here the conv output never exists in memory:
//   forward : a CTA stages the three input rows of an output row (3 x 82 floats) in shared memory, every thread forms its 20
//             conv outputs in registers, the row statistics follow as in ln_act.cu, y is written once (+ mean, rstd);
//   backward: the conv output is recomputed from the features (20 MB, L2 resident), dy is read once, the gradient of the conv
//             output stays in registers and feeds the 3 x 3 weight gradient, the conv-bias gradient and dgamma / dbeta -
//             per-CTA partial rows for cm_reduce_multi (fixed order, no atomics).  The features are an input of the network:
//             no input gradient is formed.
// Mapping as in ln_act.cu with G = 128: thread gl owns the quads q = gl + 128 i (i < 5) of the (F', C) row; 512 % C == 0 makes
// the four channels of a thread the same for all its quads (c0 = 4 gl mod C, f'_i = 4 gl / C + i * 512 / C), so the weight
// gradient accumulates in 36 registers per thread.  HBM-bound by y (forward) / dy (backward): s bytes per element each way.
#include "ln_act_common.cuh"

namespace cm {
namespace stem {

using cm::lna::kNQ;
using cm::lna::kThreads;
using cm::lna::Quad;

constexpr int kMaxC = 128;                 // channels (w_s below)
constexpr int kStage = 4;                  // staged input elements per thread: 3 * (feats + 2) <= kStage * kThreads
constexpr int kMaxFeats = kStage * kThreads / 3 - 2;   // 168
constexpr int kFp = kMaxFeats + 2;         // padded row: index j = f_in + 1
constexpr int kBwdCtasPerSm = 3;

struct Params {
  const void* in;
  const float* weight;
  const float* bias;
  const float* gamma;
  const float* beta;
  void* y;
  float* mean;
  float* rstd;
  const void* dy;
  float* dg_part;
  float* db_part;
  float* dw_part;
  float* dcb_part;
  int batch, frames, feats, C, t_out, f_out;
  float eps, slope;
};

template <typename T> __device__ __forceinline__ float ld1(const void* p, int64_t i);
template <> __device__ __forceinline__ float ld1<float>(const void* p, int64_t i) { return __ldg(static_cast<const float*>(p) + i); }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const void* p, int64_t i) {
  return __bfloat162float(__ldg(static_cast<const __nv_bfloat16*>(p) + i));
}
template <> __device__ __forceinline__ float ld1<__half>(const void* p, int64_t i) {
  return __half2float(__ldg(static_cast<const __half*>(p) + i));
}

// the three input rows of output row (b, to), zero padded, as kStage registers per thread (element idx = tid + s * 128)
template <typename TI>
__device__ __forceinline__ void load_rows(const Params& p, int64_t row, float* r) {
  const int b = (int)(row / p.t_out), to = (int)(row - (int64_t)b * p.t_out);
  const int fp = p.feats + 2;
#pragma unroll
  for (int s = 0; s < kStage; ++s) {
    const int idx = threadIdx.x + s * kThreads;
    const int kh = idx / fp, j = idx - kh * fp;
    const int t = 2 * to + kh - 1, f = j - 1;
    float v = 0.f;
    if (kh < 3 && t >= 0 && t < p.frames && f >= 0 && f < p.feats) v = ld1<TI>(p.in, ((int64_t)b * p.frames + t) * p.feats + f);
    r[s] = v;
  }
}
__device__ __forceinline__ void store_rows(const Params& p, const float* r, float (*in_s)[kFp]) {
  const int fp = p.feats + 2;
#pragma unroll
  for (int s = 0; s < kStage; ++s) {
    const int idx = threadIdx.x + s * kThreads;
    const int kh = idx / fp, j = idx - kh * fp;
    if (kh < 3) in_s[kh][j] = r[s];
  }
}

// conv outputs (+ bias) of this thread's quads from the staged rows
__device__ __forceinline__ void conv_row(const float (*in_s)[kFp], const float4* w_s, const float4 cb, const int c4, const int C4,
                                         const int* fq, float (*v)[4]) {
#pragma unroll
  for (int i = 0; i < kNQ; ++i) { v[i][0] = cb.x; v[i][1] = cb.y; v[i][2] = cb.z; v[i][3] = cb.w; }
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const float4 w = w_s[(kh * 3 + kw) * C4 + c4];
#pragma unroll
      for (int i = 0; i < kNQ; ++i) {
        const float t = in_s[kh][2 * fq[i] + kw];
        v[i][0] = fmaf(w.x, t, v[i][0]); v[i][1] = fmaf(w.y, t, v[i][1]);
        v[i][2] = fmaf(w.z, t, v[i][2]); v[i][3] = fmaf(w.w, t, v[i][3]);
      }
    }
}

__device__ __forceinline__ void load_weights(const Params& p, float4* w_s) {
  // w_s[k][c] = weight[c][k]  (torch (C, 1, 3, 3) -> tap-major so a thread's four channels are one float4)
  float* w = reinterpret_cast<float*>(w_s);
  for (int i = threadIdx.x; i < 9 * p.C; i += kThreads) {
    const int k = i / p.C, c = i - k * p.C;
    w[i] = __ldg(p.weight + c * 9 + k);
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kThreads) stem_fwd_kernel(const Params p) {
  constexpr int ES = (int)sizeof(TO);
  __shared__ float2 red[2][kThreads / 32];
  __shared__ float in_s[3][kFp];
  __shared__ float4 w_s[9 * kMaxC / 4];
  const int gl = threadIdx.x;
  const int cols = p.f_out * p.C, nq = cols >> 2, C4 = p.C >> 2;
  const int c4 = gl % C4, f0 = gl / C4, fstep = kThreads / C4;
  const float inv_n = 1.0f / (float)cols;
  const int64_t rows = (int64_t)p.batch * p.t_out;
  load_weights(p, w_s);
  int fq[kNQ];                                   // f' of each quad; idle quads (f' >= f_out) are clamped into the staged row
#pragma unroll
  for (int i = 0; i < kNQ; ++i) fq[i] = min(f0 + i * fstep, p.f_out - 1);
  float4 gm[kNQ], bt[kNQ];
#pragma unroll
  for (int i = 0; i < kNQ; ++i) {
    const int q = gl + i * kThreads;
    gm[i] = q < nq ? __ldg(reinterpret_cast<const float4*>(p.gamma) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[i] = q < nq ? __ldg(reinterpret_cast<const float4*>(p.beta) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4 cb = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  int par = 0;
  float stage[kStage];
  if ((int64_t)blockIdx.x < rows) load_rows<TI>(p, blockIdx.x, stage);
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    store_rows(p, stage, in_s);
    __syncthreads();
    if (row + gridDim.x < rows) load_rows<TI>(p, row + gridDim.x, stage);
    float v[kNQ][4];
    conv_row(in_s, w_s, cb, c4, C4, fq, v);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const bool qv = gl + i * kThreads < nq;
#pragma unroll
      for (int e = 0; e < 4; ++e) v[i][e] = qv ? v[i][e] : 0.f;
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    const float mu = cm::lna::group_sum<kThreads>(make_float2(s, 0.f), red, par).x * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const bool qv = gl + i * kThreads < nq;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[i][e] = qv ? v[i][e] - mu : 0.f;
        sq = fmaf(v[i][e], v[i][e], sq);
      }
    }
    const float var = cm::lna::group_sum<kThreads>(make_float2(sq, 0.f), red, par).x * inv_n;
    const float rs = rsqrtf(var + p.eps);
    char* py = static_cast<char*>(p.y) + row * (int64_t)cols * ES;
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const int q = gl + i * kThreads;
      if (q < nq) {
        const float g4[4] = {gm[i].x, gm[i].y, gm[i].z, gm[i].w}, b4[4] = {bt[i].x, bt[i].y, bt[i].z, bt[i].w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float t = fmaf(v[i][e] * rs, g4[e], b4[e]);
          o[e] = t > 0.f ? t : t * p.slope;
        }
        cm::lna::st4<TO>(py + (int64_t)q * 4 * ES, o);
      }
    }
    if (gl == 0) {
      p.mean[row] = mu;
      p.rstd[row] = rs;
    }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kThreads, kBwdCtasPerSm) stem_bwd_kernel(const Params p) {
  constexpr int ES = (int)sizeof(TO);
  __shared__ float2 red[2][kThreads / 32];
  __shared__ float in_s2[2][3][kFp];     // by row parity: the weight-gradient pass reads the rows after the last barrier
  __shared__ float4 w_s[9 * kMaxC / 4];
  __shared__ float4 acc_s[kThreads][10];       // per-thread weight / bias gradient accumulators, for the final sum
  const int gl = threadIdx.x;
  const int cols = p.f_out * p.C, nq = cols >> 2, C4 = p.C >> 2;
  const int c4 = gl % C4, f0 = gl / C4, fstep = kThreads / C4;
  const float inv_n = 1.0f / (float)cols;
  const int64_t rows = (int64_t)p.batch * p.t_out;
  load_weights(p, w_s);
  const float4 cb = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  int fq[kNQ];
#pragma unroll
  for (int i = 0; i < kNQ; ++i) fq[i] = min(f0 + i * fstep, p.f_out - 1);
  float dg[kNQ][4], db[kNQ][4], dw[9][4], dcb[4];
#pragma unroll
  for (int i = 0; i < kNQ; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dg[i][e] = db[i][e] = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) dw[k][e] = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) dcb[e] = 0.f;
  int par = 0;
  float stage[kStage];
  typename Quad<TO>::Raw rdy[kNQ];
  auto load_dy = [&](int64_t row) {
    const char* pdy = static_cast<const char*>(p.dy) + row * (int64_t)cols * ES;
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const int q = gl + i * kThreads;
      rdy[i] = q < nq ? Quad<TO>::ld_nc(pdy + (int64_t)q * 4 * ES) : Quad<TO>::zero();
    }
  };
  if ((int64_t)blockIdx.x < rows) {
    load_rows<TI>(p, blockIdx.x, stage);
    load_dy(blockIdx.x);
  }
  int buf = 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x, buf ^= 1) {
    float (*in_s)[kFp] = in_s2[buf];
    store_rows(p, stage, in_s);
    __syncthreads();
    const bool more = row + gridDim.x < rows;
    if (more) load_rows<TI>(p, row + gridDim.x, stage);
    const float mu = __ldg(p.mean + row), rs = __ldg(p.rstd + row);
    float xh[kNQ][4], dxh[kNQ][4];
    conv_row(in_s, w_s, cb, c4, C4, fq, xh);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const int q = gl + i * kThreads;
      const bool qv = q < nq;
      const float4 g4v = qv ? __ldg(reinterpret_cast<const float4*>(p.gamma) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 b4v = qv ? __ldg(reinterpret_cast<const float4*>(p.beta) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float g4[4] = {g4v.x, g4v.y, g4v.z, g4v.w}, b4[4] = {b4v.x, b4v.y, b4v.z, b4v.w};
      float dv[4];
      Quad<TO>::cvt(rdy[i], dv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h = qv ? (xh[i][e] - mu) * rs : 0.f;
        const float pre = fmaf(h, g4[e], b4[e]);
        const float gg = pre > 0.f ? dv[e] : dv[e] * p.slope;     // zero for idle quads (dy loaded as 0)
        dg[i][e] = fmaf(gg, h, dg[i][e]);
        db[i][e] += gg;
        const float d = gg * g4[e];
        s1 += d;
        s2 = fmaf(d, h, s2);
        xh[i][e] = h;
        dxh[i][e] = d;
      }
    }
    if (more) load_dy(row + gridDim.x);        // in flight during the reduction and the weight-gradient pass
    const float2 ss = cm::lna::group_sum<kThreads>(make_float2(s1, s2), red, par);
    const float m1 = ss.x * inv_n, m2 = ss.y * inv_n;
    // gradient of the conv output (registers only) -> conv bias and 3 x 3 weight gradients
#pragma unroll
    for (int i = 0; i < kNQ; ++i) {
      const bool qv = gl + i * kThreads < nq;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = qv ? rs * (dxh[i][e] - m1 - xh[i][e] * m2) : 0.f;
        dxh[i][e] = d;
        dcb[e] += d;
      }
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int i = 0; i < kNQ; ++i) {
          const float t = in_s[kh][2 * fq[i] + kw];
#pragma unroll
          for (int e = 0; e < 4; ++e) dw[kh * 3 + kw][e] = fmaf(dxh[i][e], t, dw[kh * 3 + kw][e]);
        }
  }
  // one partial row per CTA, fixed order
  float* og = p.dg_part + (int64_t)blockIdx.x * cols;
  float* ob = p.db_part + (int64_t)blockIdx.x * cols;
#pragma unroll
  for (int i = 0; i < kNQ; ++i) {
    const int q = gl + i * kThreads;
    if (q < nq) {
      reinterpret_cast<float4*>(og)[q] = make_float4(dg[i][0], dg[i][1], dg[i][2], dg[i][3]);
      reinterpret_cast<float4*>(ob)[q] = make_float4(db[i][0], db[i][1], db[i][2], db[i][3]);
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) acc_s[gl][k] = make_float4(dw[k][0], dw[k][1], dw[k][2], dw[k][3]);
  acc_s[gl][9] = make_float4(dcb[0], dcb[1], dcb[2], dcb[3]);
  __syncthreads();
  // (channel c, tap k) <- sum over the 512 / C threads that own channel c, in increasing thread order; k = 9: conv bias
  float* ow = p.dw_part + (int64_t)blockIdx.x * p.C * 9;
  float* oc = p.dcb_part + (int64_t)blockIdx.x * p.C;
  for (int idx = threadIdx.x; idx < p.C * 10; idx += kThreads) {
    const int c = idx / 10, k = idx - c * 10;
    float s = 0.f;
    for (int t = c >> 2; t < kThreads; t += C4) s += reinterpret_cast<const float*>(&acc_s[t][k])[c & 3];
    if (k < 9) ow[c * 9 + k] = s;
    else oc[c] = s;
  }
}

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

static int bwd_grid(int64_t rows) {
  const int64_t cap = (int64_t)sm_count() * kBwdCtasPerSm;
  return (int)(rows < cap ? (rows < 1 ? 1 : rows) : cap);
}

static bool shape_ok(const cm_stem_args* a) {
  if (a->batch <= 0 || a->frames <= 0 || a->feats <= 0 || a->channels <= 0) return false;
  const int C = a->channels;
  if ((C & 3) || C > kMaxC || (4 * kThreads) % C != 0) return false;
  if (a->feats > kMaxFeats) return false;
  const int64_t cols = (int64_t)((a->feats - 1) / 2 + 1) * C;
  return cols <= cm::lna::kMaxCols;
}

static Params make_params(const cm_stem_args* a) {
  Params p;
  p.in = a->in; p.weight = a->weight; p.bias = a->bias; p.gamma = a->gamma; p.beta = a->beta;
  p.y = a->y; p.mean = a->mean; p.rstd = a->rstd; p.dy = a->dy;
  p.dg_part = a->dgamma_part; p.db_part = a->dbeta_part; p.dw_part = a->dweight_part; p.dcb_part = a->dbias_part;
  p.batch = a->batch; p.frames = a->frames; p.feats = a->feats; p.C = a->channels;
  p.t_out = (a->frames - 1) / 2 + 1; p.f_out = (a->feats - 1) / 2 + 1;
  p.eps = a->eps; p.slope = a->slope;
  return p;
}

template <typename TI, typename TO>
static int fwd_launch(const Params& p, cudaStream_t st) {
  static int wave = 0;   // idempotent; a benign race computes it twice
  auto kern = stem_fwd_kernel<TI, TO>;
  if (wave == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0) != cudaSuccess || per_sm <= 0) {
      (void)cudaGetLastError();
      per_sm = 4;
    }
    wave = per_sm * sm_count();
  }
  const int64_t rows = (int64_t)p.batch * p.t_out;
  kern<<<(int)(rows < wave ? rows : wave), kThreads, 0, st>>>(p);
  CM_LAUNCH_CHECK();
  return 0;
}

template <typename TI, typename TO>
static int bwd_launch(const Params& p, cudaStream_t st) {
  const int64_t rows = (int64_t)p.batch * p.t_out;
  stem_bwd_kernel<TI, TO><<<bwd_grid(rows), kThreads, 0, st>>>(p);
  CM_LAUNCH_CHECK();
  return 0;
}

template <typename TI>
static int dispatch_out(const cm_stem_args* a, const Params& p, cudaStream_t st, bool bwd) {
  switch (a->out_dtype) {
    case CM_F32: return bwd ? bwd_launch<TI, float>(p, st) : fwd_launch<TI, float>(p, st);
    case CM_BF16: return bwd ? bwd_launch<TI, __nv_bfloat16>(p, st) : fwd_launch<TI, __nv_bfloat16>(p, st);
    default: return bwd ? bwd_launch<TI, __half>(p, st) : fwd_launch<TI, __half>(p, st);
  }
}

static int dispatch(const cm_stem_args* a, cudaStream_t st, bool bwd) {
  const Params p = make_params(a);
  switch (a->in_dtype) {
    case CM_F32: return dispatch_out<float>(a, p, st, bwd);
    case CM_BF16: return dispatch_out<__nv_bfloat16>(a, p, st, bwd);
    default: return dispatch_out<__half>(a, p, st, bwd);
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace stem
}  // namespace cm

extern "C" int cm_stem_supported(int32_t feats, int32_t channels) {
  cm_stem_args a{};
  a.batch = 1; a.frames = 1; a.feats = feats; a.channels = channels;
  return cm::stem::shape_ok(&a) ? 1 : 0;
}

extern "C" int cm_stem_num_part(int32_t batch, int32_t frames) {
  if (batch <= 0 || frames <= 0) return 1;
  return cm::stem::bwd_grid((int64_t)batch * ((frames - 1) / 2 + 1));
}

extern "C" int cm_stem_fwd(const cm_stem_args* a, void* stream) {
  if (!a || !a->in || !a->weight || !a->gamma || !a->beta || !a->y || !a->mean || !a->rstd) return CM_ERR_BAD_ARG;
  if (!cm::dtype_ok(a->in_dtype) || !cm::dtype_ok(a->out_dtype) || !cm::stem::shape_ok(a)) return CM_ERR_UNSUPPORTED;
  if (!cm::stem::al16(a->y) || !cm::stem::al16(a->gamma) || !cm::stem::al16(a->beta) || (a->bias && !cm::stem::al16(a->bias)))
    return CM_ERR_UNSUPPORTED;
  return cm::stem::dispatch(a, static_cast<cudaStream_t>(stream), false);
}

extern "C" int cm_stem_bwd(const cm_stem_args* a, void* stream) {
  if (!a || !a->in || !a->weight || !a->gamma || !a->beta || !a->mean || !a->rstd || !a->dy || !a->dgamma_part ||
      !a->dbeta_part || !a->dweight_part || !a->dbias_part)
    return CM_ERR_BAD_ARG;
  if (!cm::dtype_ok(a->in_dtype) || !cm::dtype_ok(a->out_dtype) || !cm::stem::shape_ok(a)) return CM_ERR_UNSUPPORTED;
  if (!cm::stem::al16(a->dy) || !cm::stem::al16(a->gamma) || !cm::stem::al16(a->beta) || (a->bias && !cm::stem::al16(a->bias)) ||
      !cm::stem::al16(a->dgamma_part) || !cm::stem::al16(a->dbeta_part))
    return CM_ERR_UNSUPPORTED;
  return cm::stem::dispatch(a, static_cast<cudaStream_t>(stream), true);
}
