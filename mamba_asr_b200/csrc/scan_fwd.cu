// Selective-scan forward for sm_100a: channel-sequential recurrence, bidirectional fusion in one launch.
//
// Replaces selective_scan_cuda.fwd (reference call sites modules/mamba/selective_scan_interface.py:42,218) and
// the flip / second call / 0.5*(a+b) of modules/mamba/bimamba.py:223-253.
//
// Mapping.  One warp owns 32/LPC channels of one batch row in ONE time direction; LPC (1, 2 or 4) lanes share
// a channel and split its 16 states.  Each lane walks time sequentially with its states in registers:
//     a = ex2(Delta * A*log2e);  h = a*h + (Delta*u)*B_n;  y += C_n*h        (1 MUFU + 4 FMA-pipe ops per state)
// so there is no scan tree, no shuffle in the LPC=1 case and - with channel-last tensors (sd == 1) - every global
// access is one coalesced row segment.  B and C (shared by all channels of a batch row) are staged per warp in
// shared memory as fp32, 32 time steps at a time, and read back as broadcast LDS.128.
//
// Bidirectional launches put the ascending warp and the descending warp of the same channels in one CTA.
// Ascending first covers [0, M), descending [M, L); both stash their pre-gate y in `out`.  After one
// __syncthreads() each continues into the half its partner already covered, reads the stash (ld.global.cg),
// adds its own y, applies out_scale * silu(z) once and writes the final value: the flip, the second output
// tensor and the add kernel of the reference disappear, and the gate is evaluated once instead of twice.
//
// Every CM_SCAN_CKPT_STEPS processed steps the warp saves its fp32 state (64 B per lane, contiguous) so the
// backward kernel can recompute states tile by tile without dividing by the decay.
#include <cstdlib>

#include "common.cuh"

namespace cm {

constexpr int kGroup = 4;                       // steps per register-prefetched, fully unrolled group
constexpr int kCkptGroups = CM_SCAN_CKPT_STEPS / kGroup;   // a checkpoint every 2 groups
constexpr int kBcPitch = 36;                    // floats per staged step (16 B + 16 C + pad, rows stay 16B aligned)

enum { MODE_UNI = 0, MODE_STASH = 1, MODE_COMBINE = 2 };

// raw (unconverted) per-lane inputs of one group, converted at the point of use.  bc[k] is this lane's share of
// the group's 4 x 32 B/C values (shared by the warp through shared memory).
template <typename T>
struct FwdGroup {
  typename Elem<T>::Raw u[kGroup], dl[kGroup], z[kGroup], st[kGroup], bc[kGroup];
};

template <typename T, int LPC, bool BC_CONST>
struct FwdCtx {
  static constexpr int NS = 16 / LPC;
  static constexpr int CPW = 32 / LPC;
  using Raw = typename Elem<T>::Raw;

  // per-lane state
  int dstate, sg, lane, mode, seqlen;
  bool dvalid, softplus, has_z, bc_time_contig;
  float scale, Dsk, bias;
  float2 kA2[NS / 2], h2[NS / 2];   // state pairs for FFMA2 / FMUL2
  float Bc[NS], Cc[NS];
  // byte pointers positioned at processed step 0; the s* members are signed BYTE strides per processed step, so an
  // address is one IMAD.WIDE (stride * step + base)
  const char *up, *dlp, *zp;
  const char* bcpA;                 // this lane's B/C source (see load_group)
  bool bc_okA;                      // ... and whether it is a real state
  const cm_scan_dir* dirp;          // kernel parameter block of this direction (slow paths re-derive from it)
  int bidx;
  bool stash_mode, pre_ok, st_ok;   // per-range output predicates (see run_range)
  int su, sdl, sz, so, sop, sbc;
  char *outp, *outprep;
  float* ckp;  // row base of the checkpoints, or nullptr
  float* bc;   // this warp's staging buffer [kGroup][kBcPitch]

  static __device__ __forceinline__ const T* at(const char* base, int stride_bytes, int s) {
    return reinterpret_cast<const T*>(base + (int64_t)stride_bytes * (int64_t)s);
  }
  static __device__ __forceinline__ T* at(char* base, int stride_bytes, int s) {
    return reinterpret_cast<T*>(base + (int64_t)stride_bytes * (int64_t)s);
  }

  // Prefetch of one group: every global load of the group is issued here, back to back, into raw registers.
  // FULL: all kGroup steps exist (no predicates in the instruction stream of the steady state).
  template <bool FULL>
  __device__ __forceinline__ void load_group(FwdGroup<T>& g, int s0, int nvalid) const {
    const bool need_z = has_z && mode != MODE_STASH;
    const bool need_st = mode == MODE_COMBINE;
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {
      const int s = s0 + k;
      g.u[k] = Raw(0); g.dl[k] = Raw(0); g.z[k] = Raw(0); g.st[k] = Raw(0); g.bc[k] = Raw(0);
      if (FULL || k < nvalid) {
        g.u[k] = Elem<T>::ld_raw(at(up, su, s));
        g.dl[k] = Elem<T>::ld_raw(at(dlp, sdl, s));
        if (need_z) g.z[k] = Elem<T>::ld_raw(at(zp, sz, s));
        if (need_st) g.st[k] = Elem<T>::ld_raw_cg(at(outp, so, s));
      }
    }
    if (!BC_CONST) {
      if (bc_time_contig) {
        // time-contiguous B/C (the reference's (B, 1, N, L) layout; compatibility path): lane -> step k = lane & 3,
        // state q = lane >> 2; the four loads are B[q], B[q + 8], C[q], C[q + 8] at that step.  Addresses are
        // re-derived from the parameter block to keep the hot path's register footprint small.
        static_assert(kGroup == 4, "lane mapping below assumes 4-step groups");
        if (FULL || (lane & 3) < nvalid) {
          const cm_scan_dir& dp = *dirp;
          const int q = lane >> 2;
          const int64_t t = dp.reverse ? (int64_t)(seqlen - 1 - (s0 + (lane & 3))) : (int64_t)(s0 + (lane & 3));
          const T* Bq = static_cast<const T*>(dp.Bm.ptr) + bidx * dp.Bm.sb + q * dp.Bm.sd + t;
          const T* Cq = static_cast<const T*>(dp.Cm.ptr) + bidx * dp.Cm.sb + q * dp.Cm.sd + t;
          if (q < dstate) { g.bc[0] = Elem<T>::ld_raw(Bq); g.bc[2] = Elem<T>::ld_raw(Cq); }
          if (q + 8 < dstate) { g.bc[1] = Elem<T>::ld_raw(Bq + 8 * dp.Bm.sd); g.bc[3] = Elem<T>::ld_raw(Cq + 8 * dp.Cm.sd); }
        }
      } else if (bc_okA) {
        // state-contiguous rows (slices of the time-major x_dbl): lane -> value v = lane, k-th load = step k
#pragma unroll
        for (int k = 0; k < kGroup; ++k)
          if (FULL || k < nvalid) g.bc[k] = Elem<T>::ld_raw(at(bcpA, sbc, s0 + k));
      }
    }
  }

  // publish the group's B/C values to the warp (fp32)
  __device__ __forceinline__ void publish_bc(const FwdGroup<T>& g) {
    if (BC_CONST) return;
    __syncwarp();
    if (bc_time_contig) {
#pragma unroll
      for (int j = 0; j < kGroup; ++j) bc[(lane & 3) * kBcPitch + 8 * j + (lane >> 2)] = Elem<T>::cvt(g.bc[j]);
    } else {
#pragma unroll
      for (int k = 0; k < kGroup; ++k) bc[k * kBcPitch + lane] = Elem<T>::cvt(g.bc[k]);
    }
    __syncwarp();
  }

  // one recurrence step; `row` = this step's staged B/C row.  State update on fp32 pairs:
  //   x2 = dt*kA ; a2 = ex2(x2) ; h2 = a2*h2 + (dt*u)*B2 ; y2 += C2*h2      -> 2 FMUL2 + 2 FFMA2 + 2 MUFU per pair
  __device__ __forceinline__ void step(Raw ur, Raw dlr, Raw zr, Raw str, int s, const float* row) {
    const float uu = Elem<T>::cvt(ur);
    const float x = Elem<T>::cvt(dlr) + bias;
    const float dt = softplus ? softplus_fwd<sizeof(T) == 4>(x) : x;
    const float du = dt * uu;
    const float2 dt2 = make_float2(dt, dt), du2 = make_float2(du, du);
    // pass 1: state update (needs B); pass 2: output contraction (needs C).  Two passes keep only one of the two
    // 16-float operand rows live at a time.
    {
      float2 B2[NS / 2];
      if (BC_CONST) {
#pragma unroll
        for (int i = 0; i < NS / 2; ++i) B2[i] = make_float2(Bc[2 * i], Bc[2 * i + 1]);
      } else {
        const float4* rb = reinterpret_cast<const float4*>(row + sg * NS);
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
          const float4 b4 = rb[i];
          B2[2 * i] = make_float2(b4.x, b4.y); B2[2 * i + 1] = make_float2(b4.z, b4.w);
        }
      }
#pragma unroll
      for (int i = 0; i < NS / 2; ++i) {
        const float2 x2 = fmul2(dt2, kA2[i]);
        const float2 a2 = make_float2(ex2(x2.x), ex2(x2.y));
        h2[i] = ffma2(a2, h2[i], fmul2(du2, B2[i]));
      }
    }
    float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
    {
      float2 C2[NS / 2];
      if (BC_CONST) {
#pragma unroll
        for (int i = 0; i < NS / 2; ++i) C2[i] = make_float2(Cc[2 * i], Cc[2 * i + 1]);
      } else {
        const float4* rc = reinterpret_cast<const float4*>(row + 16 + sg * NS);
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
          const float4 c4 = rc[i];
          C2[2 * i] = make_float2(c4.x, c4.y); C2[2 * i + 1] = make_float2(c4.z, c4.w);
        }
      }
#pragma unroll
      for (int i = 0; i < NS / 2; ++i) {
        if (i & 1) yb = ffma2(C2[i], h2[i], yb); else ya = ffma2(C2[i], h2[i], ya);
      }
    }
    const float2 ys = fadd2(ya, yb);
    float y = ys.x + ys.y;
    if (LPC >= 2) y += __shfl_xor_sync(0xffffffffu, y, 1);
    if (LPC >= 4) y += __shfl_xor_sync(0xffffffffu, y, 2);
    y = fmaf(Dsk, uu, y);
    // branch-free output: first-half steps stash the raw y, the others write the gated sum (stash is 0 and the
    // gate input is 0 whenever they do not apply)
    const float tot = y + Elem<T>::cvt(str);
    const float zz = Elem<T>::cvt(zr);
    const float gate = has_z ? zz * sigmoid_sel<sizeof(T) == 4>(zz) : 1.f;
    const float val = stash_mode ? y : tot * gate * scale;
    if (pre_ok) Elem<T>::st(at(outprep, sop, s), tot);
    if (st_ok) Elem<T>::st(at(outp, so, s), val);
  }

  __device__ __forceinline__ void save_ckpt(int j) const {
    if (ckp != nullptr && dvalid) {
      float4* dst = reinterpret_cast<float4*>(ckp + (int64_t)j * 16 + sg * NS);
#pragma unroll
      for (int i = 0; i < NS / 4; ++i) dst[i] = make_float4(h2[2 * i].x, h2[2 * i].y, h2[2 * i + 1].x, h2[2 * i + 1].y);
    }
  }

  // Processes steps [s_begin, s_end) in groups of kGroup.  While group g computes (fully unrolled, branch-free),
  // every load of group g+1 - u, delta, z, stash and the B/C rows - is already in flight.
  __device__ __forceinline__ void do_group(const FwdGroup<T>& cur, FwdGroup<T>& nxt, int g, int s_begin, int s_end,
                                           int nfull, int j0) {
    const int s0 = s_begin + g * kGroup;
    publish_bc(cur);
    if (g + 1 < nfull) load_group<true>(nxt, s0 + kGroup, kGroup);
    else load_group<false>(nxt, s0 + kGroup, max(0, s_end - s0 - kGroup));
    if ((g % kCkptGroups) == 0) save_ckpt(j0 + g / kCkptGroups);
    if (g < nfull) {
#pragma unroll
      for (int k = 0; k < kGroup; ++k) {
        step(cur.u[k], cur.dl[k], cur.z[k], cur.st[k], s0 + k, bc + k * kBcPitch);
#ifdef CM_FWD_STEP_FENCE
        asm volatile("" ::: "memory");   // keep the B/C LDS of later steps from being hoisted (register pressure)
#endif
      }
    } else {
      const int nvalid = s_end - s0;
#pragma unroll
      for (int k = 0; k < kGroup - 1; ++k)
        if (k < nvalid) step(cur.u[k], cur.dl[k], cur.z[k], cur.st[k], s0 + k, bc + k * kBcPitch);
    }
  }

  __device__ __forceinline__ void run_range(int s_begin, int s_end, int j0) {
    if (s_begin >= s_end) return;
    stash_mode = mode == MODE_STASH;
    st_ok = (sg == 0) && dvalid;
    pre_ok = st_ok && !stash_mode && outprep != nullptr;
    const int n = s_end - s_begin;
    const int ngroup = (n + kGroup - 1) / kGroup;
    const int nfull = n / kGroup;
    FwdGroup<T> ga, gb;     // ping-pong prefetch buffers (no register copies between groups)
    if (nfull > 0) load_group<true>(ga, s_begin, kGroup); else load_group<false>(ga, s_begin, n);
#pragma unroll 1
    for (int g = 0; g < ngroup; g += 2) {
      do_group(ga, gb, g, s_begin, s_end, nfull, j0);
      if (g + 1 < ngroup) do_group(gb, ga, g + 1, s_begin, s_end, nfull, j0);
    }
  }
};

#ifndef CM_FWD_MINB
#define CM_FWD_MINB 8   // <= 128 registers: 8 two-warp CTAs per SM, so e.g. 64 x 512 channels x 2 directions is one wave
#endif
template <typename T, int LPC, bool BC_CONST>
__global__ void __launch_bounds__(64, CM_FWD_MINB) scan_fwd_kernel(const __grid_constant__ cm_scan_fwd_args p) {
  using Ctx = FwdCtx<T, LPC, BC_CONST>;
  constexpr int NS = Ctx::NS, CPW = Ctx::CPW;
  __shared__ __align__(16) float bc_smem[2][kGroup * kBcPitch];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const cm_scan_dir& dp = (warp == 0) ? p.dir[0] : p.dir[1];
  const int b = blockIdx.y;
  const int cl = lane / LPC;
  int d = blockIdx.x * CPW + cl;
  const int L = p.seqlen;

  Ctx c;
  c.dstate = p.dstate;
  c.sg = lane % LPC;
  c.lane = lane;
  c.dvalid = d < p.dim;
  if (!c.dvalid) d = p.dim - 1;
  c.softplus = (p.flags & CM_FLAG_DELTA_SOFTPLUS) != 0;
  c.has_z = p.z.ptr != nullptr;
  c.scale = p.out_scale;
  c.Dsk = dp.Dskip ? __ldg(dp.Dskip + d) : 0.f;
  c.bias = dp.delta_bias ? __ldg(dp.delta_bias + d) : 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int n = c.sg * NS + i;
    const float ka = (n < p.dstate) ? __ldg(dp.A + d * dp.A_sd + n * dp.A_sn) * kLog2e : 0.f;
    if (i & 1) c.kA2[i / 2].y = ka; else c.kA2[i / 2].x = ka;
    if (i & 1) c.h2[i / 2].y = 0.f; else c.h2[i / 2].x = 0.f;
    c.Bc[i] = 0.f;
    c.Cc[i] = 0.f;
    if (BC_CONST && n < p.dstate) {
      c.Bc[i] = __ldg(static_cast<const float*>(dp.Bm.ptr) + d * dp.Bm.sb + n * dp.Bm.sd);   // constants are fp32
      c.Cc[i] = __ldg(static_cast<const float*>(dp.Cm.ptr) + d * dp.Cm.sb + n * dp.Cm.sd);
    }
  }
  // step-0 positions and signed per-step BYTE strides (descending time = negative stride)
  const int64_t l0 = (dp.reverse != 0) ? (L - 1) : 0;
  const int sgn = (dp.reverse != 0) ? -1 : 1;
  constexpr int ES = (int)sizeof(T);
  auto cptr = [](const void* base, int64_t elems) { return static_cast<const char*>(base) + elems * ES; };
  c.up = cptr(dp.u.ptr, b * dp.u.sb + d * dp.u.sd + l0 * dp.u.sl);
  c.su = sgn * ES * (int)dp.u.sl;
  c.dlp = cptr(dp.delta.ptr, b * dp.delta.sb + d * dp.delta.sd + l0 * dp.delta.sl);
  c.sdl = sgn * ES * (int)dp.delta.sl;
  c.zp = c.has_z ? cptr(p.z.ptr, b * p.z.sb + d * p.z.sd + l0 * p.z.sl) : nullptr;
  c.sz = sgn * ES * (int)p.z.sl;
  c.bcpA = nullptr;
  c.bc_okA = false;
  c.sbc = 0;
  c.bc_time_contig = false;
  c.dirp = &dp;
  c.bidx = b;
  c.seqlen = L;
  if (!BC_CONST) {
    const int64_t Boff = b * dp.Bm.sb + l0 * dp.Bm.sl, Coff = b * dp.Cm.sb + l0 * dp.Cm.sl;
    c.bc_time_contig = (dp.Bm.sl == 1 && dp.Cm.sl == 1);
    if (!c.bc_time_contig) {
      const int n = lane & 15;
      c.bcpA = (lane < 16) ? cptr(dp.Bm.ptr, Boff + n * dp.Bm.sd) : cptr(dp.Cm.ptr, Coff + n * dp.Cm.sd);
      c.bc_okA = n < p.dstate;
      c.sbc = sgn * ES * (int)((lane < 16) ? dp.Bm.sl : dp.Cm.sl);
    }
  }
  c.outp = const_cast<char*>(cptr(p.out.ptr, b * p.out.sb + d * p.out.sd + l0 * p.out.sl));
  c.so = sgn * ES * (int)p.out.sl;
  c.outprep = p.out_pre.ptr
                  ? const_cast<char*>(cptr(p.out_pre.ptr, b * p.out_pre.sb + d * p.out_pre.sd + l0 * p.out_pre.sl))
                  : nullptr;
  c.sop = sgn * ES * (int)p.out_pre.sl;
  c.ckp = dp.ckpt ? dp.ckpt + b * dp.ckpt_sb + d * dp.ckpt_sd : nullptr;
  c.bc = bc_smem[warp];

  // One copy of the range body in the instruction stream (it must stay within the instruction cache): the
  // bidirectional launch runs it twice with a CTA barrier in between.
  const int s1 = cm_first_range(L, p.ndir, dp.reverse);
  const int nrange = (p.ndir == 2) ? 2 : 1;
#pragma unroll 1
  for (int range = 0; range < nrange; ++range) {
    if (range == 1) __syncthreads();  // partner's stash for the other half is now visible (same CTA, ld.global.cg)
    c.mode = (p.ndir == 1) ? MODE_UNI : (range == 0 ? MODE_STASH : MODE_COMBINE);
    c.run_range(range == 0 ? 0 : s1, range == 0 ? s1 : L, range == 0 ? 0 : cm_ceil_div(s1, CM_SCAN_CKPT_STEPS));
  }

  if (dp.last_state != nullptr && c.dvalid) {
    float* ls = dp.last_state + b * dp.ls_sb + d * dp.ls_sd;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int n = c.sg * NS + i;
      if (n < p.dstate) ls[n * dp.ls_sn] = (i & 1) ? c.h2[i / 2].y : c.h2[i / 2].x;
    }
  }
}

template <typename T>
static int launch_fwd_t(const cm_scan_fwd_args& a, int lpc, bool bc_const, cudaStream_t st) {
  const dim3 block(32 * a.ndir);
#define CM_FWD_CASE(LPC_, BCC_)                                                  \
  {                                                                              \
    const dim3 grid(cm_ceil_div(a.dim, 32 / LPC_), a.batch);                     \
    scan_fwd_kernel<T, LPC_, BCC_><<<grid, block, 0, st>>>(a);                   \
  }
  if (!bc_const) {
    if (lpc == 1) CM_FWD_CASE(1, false) else if (lpc == 2) CM_FWD_CASE(2, false) else CM_FWD_CASE(4, false)
  } else {
    if (lpc == 1) CM_FWD_CASE(1, true) else if (lpc == 2) CM_FWD_CASE(2, true) else CM_FWD_CASE(4, true)
  }
#undef CM_FWD_CASE
  CM_LAUNCH_CHECK();
  return 0;
}

}  // namespace cm

extern "C" int cm_scan_num_ckpt(int32_t seqlen, int32_t ndir) {
  if (seqlen <= 0) return 0;
  if (ndir == 2) {
    const int m = cm_mid(seqlen);
    return cm_ceil_div(m, CM_SCAN_CKPT_STEPS) + cm_ceil_div(seqlen - m, CM_SCAN_CKPT_STEPS);
  }
  return cm_ceil_div(seqlen, CM_SCAN_CKPT_STEPS);
}

extern "C" int cm_scan_slab_channels(int32_t lanes_per_channel) {
  if (lanes_per_channel != 1 && lanes_per_channel != 2 && lanes_per_channel != 4) return CM_ERR_BAD_ARG;
  return 32 / lanes_per_channel;
}

extern "C" int cm_scan_pick_lanes(int32_t batch, int32_t dim, int32_t ndir) {
  // One warp per SM sub-partition (148 SMs x 4) saturates the MUFU pipes; below that, split each channel's
  // states over more lanes to put more warps in flight.
  const int64_t target = 148 * 4;
  for (int lpc = 1; lpc <= 2; lpc *= 2) {
    const int64_t warps = (int64_t)ndir * batch * cm_ceil_div(dim, 32 / lpc);
    if (warps >= target) return lpc;
  }
  return 4;
}

static int check_dir(const cm_scan_dir& d) {
  if (!d.u.ptr || !d.delta.ptr || !d.Bm.ptr || !d.Cm.ptr || !d.A) return CM_ERR_BAD_ARG;
  return 0;
}

namespace cm {
int scan_fwd_try_channel_last(const cm_scan_fwd_args& a, int lpc, cudaStream_t st, int* rc);   // scan_fwd_cl.cu
int scan_fwd_try_state_parallel(const cm_scan_fwd_args& a, cudaStream_t st, int* rc);            // scan_fwd_sp.cu
int64_t scan_fwd_sp_workspace_bytes(const cm_scan_fwd_args& a);                                  // scan_fwd_sp.cu
int scan_fwd_try_lane_channel(const cm_scan_fwd_args& a, cudaStream_t st, int* rc);              // scan_fwd_lc.cu
int scan_fwd_try_warpgroup(const cm_scan_fwd_args& a, cudaStream_t st, int* rc);                 // scan_fwd_wg.cu
}

extern "C" int64_t cm_scan_fwd_workspace_bytes(const cm_scan_fwd_args* args) {
  if (args == nullptr || args->batch <= 0 || args->dim <= 0 || args->seqlen <= 0) return 0;
  if (args->ndir != 1 && args->ndir != 2) return 0;
  if (args->lanes_per_channel != 0 || getenv("CM_SCAN_GENERIC") != nullptr || getenv("CM_SCAN_NO_SP") != nullptr) return 0;
  return cm::scan_fwd_sp_workspace_bytes(*args);
}

extern "C" int cm_scan_fwd(const cm_scan_fwd_args* args, void* stream) {
  if (args == nullptr) return CM_ERR_BAD_ARG;
  const cm_scan_fwd_args& a = *args;
  if (a.batch <= 0 || a.dim <= 0 || a.seqlen <= 0 || a.dstate <= 0) return CM_ERR_BAD_ARG;
  if (a.ndir != 1 && a.ndir != 2) return CM_ERR_BAD_ARG;
  if (!cm::dtype_ok(a.dtype) || a.out.ptr == nullptr) return CM_ERR_BAD_ARG;
  if (a.dstate > CM_MAX_DSTATE) return CM_ERR_UNSUPPORTED;
  if (a.batch > 65535) return CM_ERR_UNSUPPORTED;
  for (int r = 0; r < a.ndir; ++r)
    if (check_dir(a.dir[r])) return CM_ERR_BAD_ARG;
  if (a.ndir == 2) {
    if ((a.dir[0].reverse != 0) == (a.dir[1].reverse != 0)) return CM_ERR_BAD_ARG;
    if (a.dir[0].bc_const != a.dir[1].bc_const) return CM_ERR_UNSUPPORTED;
  }
  int lpc = a.lanes_per_channel;
  if (lpc != 0 && lpc != 1 && lpc != 2 && lpc != 4) return CM_ERR_BAD_ARG;
  const bool bcc = a.dir[0].bc_const != 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (getenv("CM_SCAN_GENERIC") == nullptr && getenv("CM_SCAN_NO_SP") == nullptr && a.lanes_per_channel == 0) {
    int rc = 0;
    // lane-per-channel TMA kernel (scan_fwd_lc.cu): the default for channel-last operands; launches that the caller lets
    // run as parallel time windows (few long sequences, workspace provided) stay on the state-parallel kernel
    const bool windowed = a.workspace != nullptr && cm::scan_fwd_sp_workspace_bytes(a) > 0 &&
                          a.workspace_bytes >= cm::scan_fwd_sp_workspace_bytes(a);
    // measured on B200 (ms, lc / sp): fp32 I/O 0.283 / 0.393 at 64 x 512 x 2 and 0.493 / 0.681 at 64 x 1024 x 2
    // (broadcast fp32 B|C rows read in place: no conversion pass); bf16 I/O 0.250 / 0.226 and 0.421 / 0.421, 0.085 / 0.059
    // at 32 x 288 x 2 -> the lane-per-channel kernel takes fp32 launches with enough rows, the state-parallel one the rest
    const int64_t rows = (int64_t)a.batch * a.dim * a.ndir / 32;
    const char* force = getenv("CM_SCAN_LC");
    const bool use_lc = force != nullptr ? (force[0] != '0') : (a.dtype == CM_F32 && rows >= 1024);
    if (!windowed && use_lc && getenv("CM_SCAN_NO_LC") == nullptr && cm::scan_fwd_try_lane_channel(a, st, &rc)) return rc;
    // warpgroup kernel (scan_fwd_wg.cu: TMA operands, tensor-core state sums, 2/3 of the instructions): measured EQUAL to the
    // state-parallel kernel at the large shapes (0.234 / 0.229 ms at 64 x 512 x 2, 0.448 / 0.45 at 64 x 1024 x 2 - both sit at
    // 65 % of the MUFU pipe with the issue slots at 55-70 %) and slower below a wave (0.069 / 0.058 ms at
    // 32 x 288 x 2), so it runs only on request
    const char* wg = getenv("CM_SCAN_WG_FWD");
    if (!windowed && force == nullptr && wg != nullptr && wg[0] != '0' && cm::scan_fwd_try_warpgroup(a, st, &rc)) return rc;
    // state-parallel kernel (lane = 4 states x 2 channels)
    if (cm::scan_fwd_try_state_parallel(a, st, &rc)) return rc;
  }
  if (getenv("CM_SCAN_GENERIC") == nullptr) {   // env switch only for A/B measurements of the two kernels
    // The cp.async-staged kernel is fastest with ONE lane per channel at every measured shape (32 x 288 ... 64 x 1024
    // channels, L 376 ... 30 k): splitting a channel over lanes duplicates the per-step scalar work.
    int rc = 0;
    if (cm::scan_fwd_try_channel_last(a, lpc == 0 ? 1 : lpc, st, &rc)) return rc;
  }
  if (lpc == 0) lpc = cm_scan_pick_lanes(a.batch, a.dim, a.ndir);
  switch (a.dtype) {
    case CM_F32: return cm::launch_fwd_t<float>(a, lpc, bcc, st);
    case CM_BF16: return cm::launch_fwd_t<__nv_bfloat16>(a, lpc, bcc, st);
    default: return cm::launch_fwd_t<__half>(a, lpc, bcc, st);
  }
}
