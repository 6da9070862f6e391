// Fused STFT-power -> mel filterbank -> dB (+ per-utterance max) and the top_db floor, for sm_100a.
//
// Replaces the torch op chain behind speechbrain.lobes.features.Fbank after torch.stft
// (reference call sites train_CTC.py:285, train_S2S.py:349):  view_as_real/transpose -> pow(2).sum(-1) ->
// matmul(fbank) -> clamp/log10 -> amax -> max.  That chain writes and re-reads (B,T,F,2), (B,T,F) and (B,T,M)
// intermediates; here the complex STFT is read once (coalesced along time, the layout torch.stft produces),
// the power tile lives in shared memory, the triangular filterbank is applied in its sparse form (each mel
// touches only its own band of bins) and only the (B,T,M) result is written.
#include <math_constants.h>

#include "common.cuh"

namespace cm {

constexpr int kFT = 32;          // frames per CTA
constexpr int kFbThreads = 256;  // 8 warps

// smem: power tile [nbins][kFT+1] + output tile [kFT][nmels+1]
__global__ void __launch_bounds__(kFbThreads) fbank_logmel_kernel(const cm_fbank_args p) {
  extern __shared__ float smem[];
  float* pw = smem;                                   // [nbins][kFT + 1]
  float* ot = smem + (size_t)p.nbins * (kFT + 1);     // [kFT][nmels + 1]
  __shared__ float wmax[kFbThreads / 32];
  __shared__ int flo[128], fhi[128];   // support [flo, fhi) of each mel filter (nmels <= 128)

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFT;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  // 1. power spectrum tile; lanes along time (s_t == 1 for torch.stft output) -> coalesced float2 loads
  const float2* st = reinterpret_cast<const float2*>(p.stft) + (int64_t)b * p.s_b;
  for (int f = warp; f < p.nbins; f += kFbThreads / 32) {
    const int t = t0 + lane;
    float v = 0.f;
    if (t < p.frames) {
      const float2 c = __ldg(st + (int64_t)f * p.s_f + (int64_t)t * p.s_t);
      v = fmaf(c.x, c.x, c.y * c.y);
    }
    pw[f * (kFT + 1) + lane] = v;
  }
  // support of each filter (first / last non-zero bin of its column); triangular filters are ~2-7 bins wide
  if (tid < p.nmels) {
    int lo = p.nbins, hi = 0;
    for (int f = 0; f < p.nbins; ++f) {
      if (__ldg(p.fbank + (int64_t)f * p.nmels + tid) != 0.f) {
        lo = min(lo, f);
        hi = f + 1;
      }
    }
    flo[tid] = lo;
    fhi[tid] = hi;
  }
  __syncthreads();

  // 2. mel projection over each filter's support, dB, running max.  thread -> (frame = lane, mel = warp + 8*i)
  float tmax = -CUDART_INF_F;
  const float k10 = p.multiplier * 0.30102999566398120f;   // multiplier * log10(2)
  for (int m = warp; m < p.nmels; m += kFbThreads / 32) {
    float acc = 0.f;
    const int hi = fhi[m];
    for (int f = flo[m]; f < hi; ++f)
      acc = fmaf(pw[f * (kFT + 1) + lane], __ldg(p.fbank + (int64_t)f * p.nmels + m), acc);
    const float db = k10 * lg2(fmaxf(acc, p.amin)) - p.db_offset;
    ot[lane * (p.nmels + 1) + m] = db;
    if (t0 + lane < p.frames) tmax = fmaxf(tmax, db);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  if (lane == 0) wmax[warp] = tmax;
  __syncthreads();

  // 3. coalesced store of the (frames x nmels) tile
  const int nfr = min(kFT, p.frames - t0);
  float* outp = p.out + ((int64_t)b * p.frames + t0) * p.nmels;
  for (int i = tid; i < nfr * p.nmels; i += kFbThreads) {
    const int fr = i / p.nmels, m = i - fr * p.nmels;
    outp[i] = ot[fr * (p.nmels + 1) + m];
  }
  if (tid == 0) {
    float mx = wmax[0];
#pragma unroll
    for (int w = 1; w < kFbThreads / 32; ++w) mx = fmaxf(mx, wmax[w]);
    // float max via ordered-int atomics (valid for mixed signs)
    int* addr = reinterpret_cast<int*>(p.utt_max + b);
    if (mx >= 0.f) atomicMax(addr, __float_as_int(mx));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(mx));
  }
}

__global__ void fbank_floor_kernel(const cm_fbank_args p) {
  const int64_t per = (int64_t)p.frames * p.nmels;
  const int b = blockIdx.y;
  const float floor_db = __ldg(p.utt_max + b) - p.top_db;
  float* o = p.out + (int64_t)b * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x)
    o[i] = fmaxf(o[i], floor_db);
}

}  // namespace cm

static int check_fbank(const cm_fbank_args* a) {
  if (a == nullptr) return CM_ERR_BAD_ARG;
  if (a->batch <= 0 || a->frames <= 0 || a->nbins <= 0 || a->nmels <= 0) return CM_ERR_BAD_ARG;
  if (!a->out || !a->utt_max) return CM_ERR_BAD_ARG;
  if (a->batch > 65535 || a->nmels > 128) return CM_ERR_UNSUPPORTED;
  return 0;
}

extern "C" int cm_fbank_logmel(const cm_fbank_args* a, void* stream) {
  const int e = check_fbank(a);
  if (e) return e;
  if (!a->stft || !a->fbank) return CM_ERR_BAD_ARG;
  const size_t smem = ((size_t)a->nbins * (cm::kFT + 1) + (size_t)cm::kFT * (a->nmels + 1)) * sizeof(float);
  if (smem > 200 * 1024) return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (smem > 48 * 1024) {
    cudaError_t ce = cudaFuncSetAttribute(cm::fbank_logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return (int)ce;
  }
  const dim3 grid(cm_ceil_div(a->frames, cm::kFT), a->batch);
  cm::fbank_logmel_kernel<<<grid, cm::kFbThreads, smem, st>>>(*a);
  CM_LAUNCH_CHECK();
  return 0;
}

extern "C" int cm_fbank_floor(const cm_fbank_args* a, void* stream) {
  const int e = check_fbank(a);
  if (e) return e;
  const int64_t per = (int64_t)a->frames * a->nmels;
  const int threads = 256;
  int bx = (int)((per + threads * 4 - 1) / (threads * 4));
  if (bx < 1) bx = 1;
  if (bx > 1184) bx = 1184;   // 8 x 148 SMs
  cm::fbank_floor_kernel<<<dim3(bx, a->batch), threads, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  CM_LAUNCH_CHECK();
  return 0;
}

extern "C" int cm_abi_sizeof(int32_t which) {
  switch (which) {
    case 0: return (int)sizeof(cm_tensor3);
    case 1: return (int)sizeof(cm_scan_dir);
    case 2: return (int)sizeof(cm_scan_fwd_args);
    case 3: return (int)sizeof(cm_scan_bwd_dir);
    case 4: return (int)sizeof(cm_scan_bwd_args);
    case 5: return (int)sizeof(cm_conv_dir);
    case 6: return (int)sizeof(cm_conv_args);
    case 7: return (int)sizeof(cm_fbank_args);
    case 8: return (int)sizeof(cm_reduce_job);
    case 9: return (int)sizeof(cm_layernorm_args);
    case 10: return (int)sizeof(cm_dwconv_args);
    case 11: return (int)sizeof(cm_ssm_step_args);
    case 12: return (int)sizeof(cm_add_ln_args);
    case 13: return (int)sizeof(cm_ln_act_args);
    case 14: return (int)sizeof(cm_adamw_args);
    case 15: return (int)sizeof(cm_fbank_wav_args);
    case 16: return (int)sizeof(cm_ctc_args);
    case 17: return (int)sizeof(cm_stem_args);
    case 18: return (int)sizeof(cm_act_args);
    case 19: return (int)sizeof(cm_reduce_job2);
    default: return CM_ERR_BAD_ARG;
  }
}

extern "C" int cm_version(int32_t* sm_arch) {
  if (sm_arch) *sm_arch = 100;
  return CM_ABI_VERSION;
}
