// Single-token selective-state update for incremental decoding (sm_100a).
//
// Replaces `selective_state_update` of mamba-ssm 1.1.3.post1 (a Triton kernel upstream; reference call site
// modules/mamba/bimamba.py:354-356) and the torch fallback the reference runs when that import is missing
// (bimamba.py:345-352):
//     dt    = softplus(dt + dt_bias)
//     state = state * exp(dt * A) + (dt * B) * x           (batch, dim, dstate), in place
//     y     = <state, C> + D * x ;  y *= silu(z)
// One launch per token.  The state is the only tensor of size batch*dim*dstate: 4 lanes own one (batch, channel) row
// (lane q holds states q, q+4, q+8, ...), so a warp reads and writes 8 consecutive rows = one contiguous 512-byte
// segment at dstate 16 / fp32; everything else is per-channel scalars.  fp32 arithmetic, one rounding on the way out.
// Roof: HBM - 2 * batch*dim*dstate*sizeof(state) bytes per token (the state read + write); at decode batch sizes the
// launch itself (a few microseconds) dominates, which is why the whole update is one kernel.
#include "common.cuh"

namespace cm {

constexpr int kStepLanes = 4;      // lanes per (batch, channel) row
constexpr int kStepThreads = 128;  // 32 rows per CTA

template <typename T, typename TS>
__global__ void __launch_bounds__(kStepThreads) ssm_step_kernel(const cm_ssm_step_args a) {
  const int q = threadIdx.x & (kStepLanes - 1);
  const int d = blockIdx.x * (kStepThreads / kStepLanes) + (threadIdx.x >> 2);
  const int b = blockIdx.y;
  const bool live = d < a.dim;
  const int dc = live ? d : a.dim - 1;      // dead lanes shadow the last channel so that the shuffles stay full-warp
  const int N = a.dstate;
  constexpr bool PRECISE = sizeof(T) == 4;

  const float x = Elem<T>::ld(static_cast<const T*>(a.x) + b * a.x_sb + dc);
  float dt = Elem<T>::ld(static_cast<const T*>(a.dt) + b * a.dt_sb + dc);
  if (a.dt_bias) dt += __ldg(a.dt_bias + dc);
  if (a.flags & CM_FLAG_DELTA_SOFTPLUS) dt = softplus_fwd<PRECISE>(dt);
  const float dtx = dt * x;
  const float dtl = dt * kLog2e;
  TS* st = static_cast<TS*>(a.state) + ((int64_t)b * a.dim + dc) * N;
  const float* Ar = a.A + (int64_t)dc * N;
  const T* Br = static_cast<const T*>(a.Bm) + b * a.b_sb;
  const T* Cr = static_cast<const T*>(a.Cm) + b * a.c_sb;
  float y = 0.f;
  for (int n = q; n < N; n += kStepLanes) {
    const float h = fmaf(static_cast<float>(st[n]), ex2(dtl * __ldg(Ar + n)), dtx * Elem<T>::ld(Br + n));
    const float hr = Elem<TS>::round(h);      // the reference reads the stored state back for y (bimamba.py:349-350)
    if (live) Elem<TS>::st(st + n, h);
    y = fmaf(hr, Elem<T>::ld(Cr + n), y);
  }
  y += __shfl_xor_sync(0xffffffffu, y, 1);
  y += __shfl_xor_sync(0xffffffffu, y, 2);
  if (q == 0 && live) {
    if (a.Dskip) y = fmaf(__ldg(a.Dskip + d), x, y);
    if (a.z) {
      const float z = Elem<T>::ld(static_cast<const T*>(a.z) + b * a.z_sb + d);
      y *= z * sigmoid_sel<PRECISE>(z);
    }
    Elem<T>::st(static_cast<T*>(a.out) + b * a.out_sb + d, y);
  }
}

template <typename T>
static int launch_step(const cm_ssm_step_args& a, cudaStream_t st) {
  const dim3 grid((a.dim + kStepThreads / kStepLanes - 1) / (kStepThreads / kStepLanes), a.batch);
  switch (a.state_dtype) {
    case CM_F32: ssm_step_kernel<T, float><<<grid, kStepThreads, 0, st>>>(a); break;
    case CM_BF16: ssm_step_kernel<T, __nv_bfloat16><<<grid, kStepThreads, 0, st>>>(a); break;
    default: ssm_step_kernel<T, __half><<<grid, kStepThreads, 0, st>>>(a); break;
  }
  CM_LAUNCH_CHECK();
  return 0;
}

}  // namespace cm

extern "C" int cm_ssm_step(const cm_ssm_step_args* args, void* stream) {
  if (!args) return CM_ERR_BAD_ARG;
  const cm_ssm_step_args& a = *args;
  if (!a.state || !a.x || !a.dt || !a.Bm || !a.Cm || !a.A || !a.out) return CM_ERR_BAD_ARG;
  if (a.batch <= 0 || a.dim <= 0 || a.dstate <= 0) return CM_ERR_BAD_ARG;
  if (!cm::dtype_ok(a.dtype) || !cm::dtype_ok(a.state_dtype)) return CM_ERR_BAD_ARG;
  if (a.batch > 65535 || a.dstate > 256) return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a.dtype) {
    case CM_F32: return cm::launch_step<float>(a, st);
    case CM_BF16: return cm::launch_step<__nv_bfloat16>(a, st);
    default: return cm::launch_step<__half>(a, st);
  }
}
