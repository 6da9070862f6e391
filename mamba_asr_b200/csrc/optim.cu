// Optimizer step of the training workloads on flat fp32 buffers (the optimizer / scheduler
// stand-ins of train_CTC.py:716-717).  Reference objects: torch.optim.AdamW(lr, betas=(0.9, 0.98), eps=1e-9, weight_decay)
// + speechbrain's gradient clipping to max_grad_norm (hparams/CTC/conmamba_large.yaml:91, 248-252).
//
//   cm_sumsq_partial : per-CTA partial sums of g^2 over a flat gradient buffer (fixed order: deterministic)
//   cm_adamw_step    : every CTA re-reduces the partials in the same fixed order -> global norm -> clip coefficient
//                      c = min(1, max_norm / (norm * gscale + 1e-6)) * gscale   (gscale = 1 / world_size folds the DDP average
//                      in), then one pass over p, g, m, v:  g' = c*g ; m = b1*m + (1-b1)*g' ; v = b2*v + (1-b2)*g'^2 ;
//                      p = p*(1 - lr*wd) - (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)        (torch.optim.AdamW, decoupled decay)
//                      and optionally the bf16 copy of p that the next forward's GEMMs read.
// Two launches per step replace torch's clip_grad_norm_ (foreach norm + stack + norm + clamp + foreach mul) and the fused
// multi-tensor AdamW; HBM traffic = 4 B (norm) + 28 B (+2 B) per parameter.  Roof: HBM.
#include "common.cuh"

namespace cm {

constexpr int kOptThreads = 256;
constexpr int kOptMaxPart = 2048;

__global__ void __launch_bounds__(kOptThreads) sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ part) {
  __shared__ float sm[kOptThreads / 32];
  const int64_t n4 = n / 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOptThreads) {
    const float4 v = __ldg(g4 + i);
    acc[0] = fmaf(v.x, v.x, acc[0]); acc[1] = fmaf(v.y, v.y, acc[1]);
    acc[2] = fmaf(v.z, v.z, acc[2]); acc[3] = fmaf(v.w, v.w, acc[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
    const float v = g[n4 * 4 + threadIdx.x];
    acc[0] = fmaf(v, v, acc[0]);
  }
  float s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) t += sm[w];
    part[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kOptThreads) adamw_kernel(cm_adamw_args a) {
  __shared__ float sm[kOptThreads / 32];
  __shared__ float coef_s;
  // global gradient norm: every CTA sums the same partials in the same order
  float coef = a.grad_scale;
  if (a.sumsq_part != nullptr) {
    float s = 0.f;
    for (int i = threadIdx.x; i < a.n_part; i += kOptThreads) s += a.sumsq_part[i];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kOptThreads / 32; ++w) t += sm[w];
      const float norm = sqrtf(t) * a.grad_scale;                 // norm of the averaged gradient
      if (blockIdx.x == 0 && a.norm_out != nullptr) *a.norm_out = norm;
      float c = a.max_grad_norm / (norm + 1e-6f);                 // torch.nn.utils.clip_grad_norm_
      c = c < 1.f ? c : 1.f;
      coef_s = (a.max_grad_norm > 0.f ? c : 1.f) * a.grad_scale;
    }
    __syncthreads();
    coef = coef_s;
  }
  const float b1 = a.beta1, b2 = a.beta2, ob1 = 1.f - a.beta1, ob2 = 1.f - a.beta2;
  const float decay = 1.f - a.lr * a.weight_decay, step = a.lr / a.bias_corr1, rs2 = rsqrtf(a.bias_corr2), eps = a.eps;
  const int64_t n4 = a.n / 4;
  float4* p4 = reinterpret_cast<float4*>(a.p);
  float4* m4 = reinterpret_cast<float4*>(a.m);
  float4* v4 = reinterpret_cast<float4*>(a.v);
  const float4* g4 = reinterpret_cast<const float4*>(a.g);
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOptThreads) {
    float4 p = p4[i], m = m4[i], v = v4[i];
    const float4 g = __ldg(g4 + i);
    float* pp = &p.x; float* mm = &m.x; float* vv = &v.x;
    const float gg[4] = {g.x * coef, g.y * coef, g.z * coef, g.w * coef};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mm[e] = fmaf(b1, mm[e], ob1 * gg[e]);
      vv[e] = fmaf(b2, vv[e], ob2 * gg[e] * gg[e]);
      pp[e] = fmaf(pp[e], decay, -step * mm[e] / (sqrtf(vv[e]) * rs2 + eps));
    }
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (a.p_bf16 != nullptr) {
      __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(a.p_bf16) + 2 * i;
      q[0] = __floats2bfloat162_rn(p.x, p.y);
      q[1] = __floats2bfloat162_rn(p.z, p.w);
    }
  }
}

}  // namespace cm

extern "C" int cm_optim_num_part(int64_t n) {
  if (n <= 0) return 0;
  const int64_t blocks = (n / 4 + cm::kOptThreads - 1) / cm::kOptThreads;
  const int64_t cap = 148 * 8;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

extern "C" int cm_sumsq_partial(const float* g, int64_t n, float* part, void* stream) {
  if (g == nullptr || part == nullptr || n <= 0) return CM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(g) & 15) != 0) return CM_ERR_UNSUPPORTED;
  cm::sumsq_partial_kernel<<<cm_optim_num_part(n), cm::kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(g, n, part);
  CM_LAUNCH_CHECK();
  return 0;
}

extern "C" int cm_adamw_step(const cm_adamw_args* args, void* stream) {
  if (args == nullptr || !args->p || !args->g || !args->m || !args->v || args->n <= 0) return CM_ERR_BAD_ARG;
  if (args->n % 4 != 0) return CM_ERR_UNSUPPORTED;                 // flat buffers are padded by the caller
  if (args->sumsq_part != nullptr && (args->n_part <= 0 || args->n_part > cm::kOptMaxPart)) return CM_ERR_BAD_ARG;
  const uintptr_t al = reinterpret_cast<uintptr_t>(args->p) | reinterpret_cast<uintptr_t>(args->g) |
                       reinterpret_cast<uintptr_t>(args->m) | reinterpret_cast<uintptr_t>(args->v) |
                       reinterpret_cast<uintptr_t>(args->p_bf16);
  if ((al & 15) != 0) return CM_ERR_UNSUPPORTED;
  if (!(args->bias_corr1 > 0.f) || !(args->bias_corr2 > 0.f)) return CM_ERR_BAD_ARG;
  const int64_t blocks = (args->n / 4 + cm::kOptThreads - 1) / cm::kOptThreads;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  cm::adamw_kernel<<<grid, cm::kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(*args);
  CM_LAUNCH_CHECK();
  return 0;
}
