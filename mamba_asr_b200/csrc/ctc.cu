// CTC loss, forward and gradient in one launch, for sm_100a (the `ctc_loss` stand-in of the layer
// shell; reference call site train_CTC.py:297-302 -> speechbrain.nnet.losses.ctc_loss -> torch.nn.functional.ctc_loss).
//
// torch evaluates it with three kernels - the alpha recursion (0.35 ms at 64 x 501 x 31 on B200), the beta recursion (0.25 ms)
// and a gather (0.17 ms): two sequential passes of T steps with one small step per global-memory round trip.  Here a CTA
// owns one utterance and runs BOTH recursions at the same time - one group of warps walks alpha forward in time, the other beta
// backward.  Each thread owns one extended-label state: its log-probabilities lp[t, l'_s] are gathered into shared memory by
// cp.async a block of steps ahead (only the 2S+1 columns the recursion reads, so the class count does not matter), the state
// value lives in a register, neighbours are exchanged through a double-buffered shared array with ONE named barrier per step
// and group.  alpha and beta go to a workspace; after a CTA barrier the gradient
//     d nll / d lp[t, c] = - sum_{s : l'_s = c} exp( alpha_t(s) + beta_t(s) + nll - lp[t, c] )
// is formed without atomics: the block is zero-filled, every distinct label of the utterance walks the chain of its
// occurrences (increasing s) with one thread per (t, label), and a warp per frame sums the blank states with a fixed shuffle
// tree - bitwise reproducible.  Each term is a path posterior (<= 1), so no running maximum is needed.
// Definitions as in torch (LossCTC.cu): alpha_t(s) and beta_t(s) both include lp[t, l'_s]; nll = -logsumexp(alpha_{T-1}(S'-1),
// alpha_{T-1}(S'-2)).  Infeasible alignments give nll = +inf and a zero gradient (the Python wrapper applies zero_infinity).
#include <limits>

#include <math_constants.h>

#include "common.cuh"

namespace cm {
namespace ctc {

constexpr int kRows = 16;                 // steps per staging slot
constexpr float kNegInf = -std::numeric_limits<float>::infinity();

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == kNegInf) return kNegInf;
  return m + kLn2 * lg2(ex2((a - m) * kLog2e) + ex2((b - m) * kLog2e));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == kNegInf) return kNegInf;
  return m + kLn2 * lg2(ex2((a - m) * kLog2e) + ex2((b - m) * kLog2e) + ex2((c - m) * kLog2e));
}
__device__ __forceinline__ void group_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// grid = batch; block = 2 * NT threads (NT = states rounded up to a warp): threads [0, NT) alpha, [NT, 2 NT) beta
__global__ void __launch_bounds__(1024) ctc_kernel(const cm_ctc_args p) {
  extern __shared__ __align__(16) unsigned char ctc_smem[];
  const int b = blockIdx.x;
  const int NT = blockDim.x >> 1;
  const int grp = threadIdx.x >= NT ? 1 : 0;            // 0: alpha (forward in time), 1: beta (backward)
  const int s = threadIdx.x - grp * NT;                 // extended-label state of this thread
  const int C = p.classes;
  const int T = p.input_lengths ? (int)max((int64_t)0, min((int64_t)p.max_time, p.input_lengths[b])) : p.max_time;
  const int S = p.target_lengths ? (int)max((int64_t)0, min((int64_t)p.max_target, p.target_lengths[b])) : p.max_target;
  const int SP = 2 * S + 1;
  // shared memory: labels[lab_words] | next[lab_words] | heads[lab_words] | exchange[2 groups][2 parities][NT]
  //                | staged[2 groups][2 slots][kRows][NT]
  const int lab_words = (p.max_target + 4) & ~3;
  int* labels = reinterpret_cast<int*>(ctc_smem);
  int* nxt_same = labels + lab_words;                   // next occurrence of the same label, or -1
  int* heads = nxt_same + lab_words;                    // first occurrences, compacted
  float* xch = reinterpret_cast<float*>(heads + lab_words);
  float* staged = xch + 4 * NT;
  __shared__ float s_nll;
  __shared__ int s_nheads;
  const float* lp = p.log_probs + (int64_t)b * p.lp_sb;
  const int64_t* tg = p.targets + (int64_t)b * p.tg_sb;
  for (int i = threadIdx.x; i < p.max_target; i += blockDim.x) {
    int v = i < S ? (int)tg[i] : p.blank;
    labels[i] = min(max(v, 0), C - 1);                  // out-of-range labels are clamped (torch leaves them undefined)
  }
  float* wa = p.workspace + (int64_t)b * 2 * p.max_time * p.ws_states;          // alpha [T][ws_states]
  float* wb = wa + (int64_t)p.max_time * p.ws_states;                           // beta
  __syncthreads();

  const bool live = s < SP;
  const int my = (live && (s & 1)) ? labels[s >> 1] : p.blank;                       // l'_s
  // may state s take the skip transition from s-2 (alpha) / to s+2 (beta)?
  const bool skip_a = live && (s & 1) && s >= 2 && labels[s >> 1] != labels[(s >> 1) - 1];
  const bool skip_b = live && (s & 1) && s + 2 < SP && labels[s >> 1] != labels[(s >> 1) + 1];
  float* myx = xch + grp * 2 * NT;
  float* mystage = staged + grp * 2 * kRows * NT;
  const int bar_id = 1 + grp;

  // gather lp[t, l'_s] for the steps [t0, t0 + kRows) (clipped to [0, T)) of this group's walk: a thread stages and reads
  // only its own column, so cp.async.wait_group is all the synchronisation the staging needs
  auto stage = [&](int slot, int t0) {
    if (live) {
      const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(mystage + slot * kRows * NT + s));
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int t = t0 + r;
        if (t >= 0 && t < T)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + r * NT * 4), "l"(lp + (int64_t)t * p.lp_st + my)
                       : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float cur = kNegInf;                                  // alpha_t(s) / beta_t(s)
  if (T > 0) {
    // the walk visits blocks of kRows steps; block j of alpha covers t = j*kRows .., of beta t = T-1 - j*kRows .. downwards
    const int nblk = (T + kRows - 1) / kRows;
    auto blk_t0 = [&](int j) { return grp == 0 ? j * kRows : T - (j + 1) * kRows; };   // lowest time of block j
    float* wsp = grp == 0 ? wa : wb;
    stage(0, blk_t0(0));
    for (int j = 0; j < nblk; ++j) {
      if (j + 1 < nblk) {
        stage((j + 1) & 1, blk_t0(j + 1));
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      const float* slot = mystage + (j & 1) * kRows * NT + s;
      const int t0 = blk_t0(j);
#pragma unroll 4
      for (int q = 0; q < kRows; ++q) {
        const int r = grp == 0 ? q : kRows - 1 - q;
        const int t = t0 + r;
        if (t < 0 || t >= T) continue;                  // (uniform over the group)
        const float lpt = live ? slot[r * NT] : kNegInf;
        const int step = grp == 0 ? t : T - 1 - t;      // 0 at the first step of the walk
        float nxt;
        if (step == 0) {
          if (grp == 0) nxt = (s == 0 || s == 1) && live ? lpt : kNegInf;
          else nxt = (s == SP - 1 || s == SP - 2) && live ? lpt : kNegInf;
        } else {
          const float* prev = myx + ((step - 1) & 1) * NT;
          float a1, a2;
          if (grp == 0) {
            a1 = s >= 1 ? prev[s - 1] : kNegInf;
            a2 = skip_a ? prev[s - 2] : kNegInf;
          } else {
            a1 = (live && s + 1 < SP) ? prev[s + 1] : kNegInf;
            a2 = skip_b ? prev[s + 2] : kNegInf;
          }
          nxt = live ? lse3(cur, a1, a2) + lpt : kNegInf;
        }
        cur = nxt;
        myx[(step & 1) * NT + s] = cur;
        if (live) wsp[(int64_t)t * p.ws_states + s] = cur;
        group_bar(bar_id, NT);                          // the step's values are in the exchange buffer
      }
    }
  }
  // chains of equal labels for the gradient (thread k < S): next occurrence of labels[k]; first occurrences are heads
  for (int k = threadIdx.x; k < S; k += blockDim.x) {
    const int c = labels[k];
    int nx = -1;
    for (int j = k + 1; j < S; ++j) if (labels[j] == c) { nx = j; break; }
    nxt_same[k] = nx;
    bool first = (c != p.blank);                        // a label equal to the blank index never matches (as in torch)
    for (int j = 0; j < k && first; ++j) first = labels[j] != c;
    heads[k] = first ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float nll = CUDART_INF_F;
    if (T > 0) {
      const float* last = xch + ((T - 1) & 1) * NT;     // alpha group's exchange buffer of the last step
      const float a = last[SP - 1], a2 = SP >= 2 ? last[SP - 2] : kNegInf;
      nll = -lse2(a, a2);
    } else if (S == 0) {
      nll = 0.f;
    }
    s_nll = nll;
    p.nll[b] = nll;
    int nh = 0;
    for (int k = 0; k < S; ++k) if (heads[k]) heads[nh++] = k;     // in-place compaction (nh <= k)
    s_nheads = nh;
  }
  if (p.grad == nullptr) return;
  // zero fill of this utterance's gradient block (rows are g_st apart; C floats each)
  float* g = p.grad + (int64_t)b * p.g_sb;
  if (p.g_st == C) {
    const int64_t n = (int64_t)p.max_time * C;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) g[i] = 0.f;
  } else {
    for (int t = 0; t < p.max_time; ++t)
      for (int c = threadIdx.x; c < C; c += blockDim.x) g[(int64_t)t * p.g_st + c] = 0.f;
  }
  __syncthreads();
  const float nll = s_nll;
  if (!(nll < CUDART_INF_F) || T <= 0) return;
  const int nh = s_nheads;
  // labels: one thread per (t, distinct label), walking the label's occurrences in increasing s
  for (int i = threadIdx.x; i < T * nh; i += blockDim.x) {
    const int t = i / nh, u = i - t * nh;
    int k = heads[u];
    const int c = labels[k];
    const float* at = wa + (int64_t)t * p.ws_states;
    const float* bt = wb + (int64_t)t * p.ws_states;
    const float off = nll - lp[(int64_t)t * p.lp_st + c];
    float acc = 0.f;
    while (k >= 0) {
      acc += ex2((at[2 * k + 1] + bt[2 * k + 1] + off) * kLog2e);
      k = nxt_same[k];
    }
    g[(int64_t)t * p.g_st + c] = -acc;
  }
  // blank: one warp per frame over the even states
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int t = warp; t < T; t += nwarps) {
    const float* at = wa + (int64_t)t * p.ws_states;
    const float* bt = wb + (int64_t)t * p.ws_states;
    const float off = nll - lp[(int64_t)t * p.lp_st + p.blank];
    float acc = 0.f;
    for (int k = lane; k <= S; k += 32) acc += ex2((at[2 * k] + bt[2 * k] + off) * kLog2e);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) g[(int64_t)t * p.g_st + p.blank] = -acc;
  }
}

}  // namespace ctc
}  // namespace cm

extern "C" int64_t cm_ctc_workspace_floats(int32_t batch, int32_t max_time, int32_t max_target) {
  if (batch <= 0 || max_time <= 0 || max_target < 0) return 0;
  return (int64_t)batch * 2 * max_time * (2 * (int64_t)max_target + 1);
}

extern "C" int cm_ctc_loss(const cm_ctc_args* a, void* stream) {
  if (a == nullptr || a->log_probs == nullptr || a->targets == nullptr || a->nll == nullptr || a->workspace == nullptr)
    return CM_ERR_BAD_ARG;
  if (a->batch <= 0 || a->max_time <= 0 || a->classes <= 0 || a->max_target < 0) return CM_ERR_BAD_ARG;
  if (a->blank < 0 || a->blank >= a->classes) return CM_ERR_BAD_ARG;
  if (a->ws_states != 2 * a->max_target + 1) return CM_ERR_BAD_ARG;
  const int nt = ((2 * a->max_target + 1) + 31) / 32 * 32;
  if (2 * nt > 1024) return CM_ERR_UNSUPPORTED;          // up to 255 labels per utterance
  const size_t smem = (size_t)(3 * ((a->max_target + 4) & ~3) + 4 * nt + 2 * 2 * cm::ctc::kRows * nt) * sizeof(float);
  if (smem > 200 * 1024) return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(cm::ctc::ctc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cm::ctc::ctc_kernel<<<a->batch, 2 * nt, smem, st>>>(*a);
  CM_LAUNCH_CHECK();
  return 0;
}
