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
#include "common.cuh"

namespace cm {

constexpr int kGroup = CM_SCAN_CKPT_STEPS;  // steps per register-prefetched group == checkpoint period
constexpr int kBcTile = 32;                 // steps per shared-memory B/C tile
constexpr int kBcPitch = 36;                // floats per staged step (32 values + pad, keeps rows 16B aligned)

enum { MODE_UNI = 0, MODE_STASH = 1, MODE_COMBINE = 2 };

template <typename T>
struct FwdGroup {
  float u[kGroup], dl[kGroup], z[kGroup], st[kGroup];
};

template <typename T, int LPC, bool BC_CONST>
struct FwdCtx {
  static constexpr int NS = 16 / LPC;
  static constexpr int CPW = 32 / LPC;

  // per-lane constants
  int L, dstate, sg, lane;
  bool rev, dvalid, softplus, has_z;
  float scale, Dsk, bias;
  float kA[NS], h[NS], Bc[NS], Cc[NS];
  const T *up, *dlp, *zp, *Bp, *Cp;
  int64_t u_sl, dl_sl, z_sl, B_sd, B_sl, C_sd, C_sl, o_sl, op_sl;
  T *outp, *outprep;
  float* ckp;  // row base of the checkpoints, or nullptr
  float* bc;   // this warp's staging buffer

  __device__ __forceinline__ int time_of(int s) const { return rev ? (L - 1 - s) : s; }

  __device__ __forceinline__ void stage_bc(int tile_start, int s_end) {
    if (BC_CONST) return;
    __syncwarp();
    if (B_sl == 1 && C_sl == 1) {
      // time-contiguous B/C (the reference's (B, 1, N, L) layout): lanes over time
      const int s = tile_start + lane;
      const bool ok = s < s_end;
      const int l = ok ? time_of(s) : 0;
#pragma unroll 8
      for (int v = 0; v < 32; ++v) {
        const int n = v & 15;
        float val = 0.f;
        if (ok && n < dstate) val = (v < 16) ? Elem<T>::ld(Bp + n * B_sd + l) : Elem<T>::ld(Cp + n * C_sd + l);
        bc[lane * kBcPitch + v] = val;
      }
    } else {
      // state-contiguous rows (slices of the time-major x_dbl): lanes over the 32 values of one step
      const int n = lane & 15;
      const T* src = (lane < 16) ? (Bp + n * B_sd) : (Cp + n * C_sd);
      const int64_t sl = (lane < 16) ? B_sl : C_sl;
#pragma unroll 8
      for (int t = 0; t < kBcTile; ++t) {
        const int s = tile_start + t;
        float val = 0.f;
        if (s < s_end && n < dstate) val = Elem<T>::ld(src + (int64_t)time_of(s) * sl);
        bc[t * kBcPitch + lane] = val;
      }
    }
    __syncwarp();
  }

  template <int MODE>
  __device__ __forceinline__ void load_group(FwdGroup<T>& g, int s0, int s_end) const {
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {
      const int s = s0 + k;
      g.u[k] = 0.f; g.dl[k] = 0.f; g.z[k] = 0.f; g.st[k] = 0.f;
      if (s < s_end) {
        const int64_t l = time_of(s);
        g.u[k] = Elem<T>::ld(up + l * u_sl);
        g.dl[k] = Elem<T>::ld(dlp + l * dl_sl);
        if (MODE != MODE_STASH && has_z) g.z[k] = Elem<T>::ld(zp + l * z_sl);
        if (MODE == MODE_COMBINE) g.st[k] = Elem<T>::ld_cg(outp + l * o_sl);
      }
    }
  }

  template <int MODE>
  __device__ __forceinline__ void compute_group(const FwdGroup<T>& g, int s0, int s_end, int tile_start) {
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {
      const int s = s0 + k;
      if (s < s_end) {
        float Bv[NS], Cv[NS];
        if (BC_CONST) {
#pragma unroll
          for (int i = 0; i < NS; ++i) { Bv[i] = Bc[i]; Cv[i] = Cc[i]; }
        } else {
          const float4* row = reinterpret_cast<const float4*>(bc + (s - tile_start) * kBcPitch + sg * NS);
          const float4* rowc = reinterpret_cast<const float4*>(bc + (s - tile_start) * kBcPitch + 16 + sg * NS);
#pragma unroll
          for (int i = 0; i < NS / 4; ++i) {
            const float4 b4 = row[i], c4 = rowc[i];
            Bv[4 * i + 0] = b4.x; Bv[4 * i + 1] = b4.y; Bv[4 * i + 2] = b4.z; Bv[4 * i + 3] = b4.w;
            Cv[4 * i + 0] = c4.x; Cv[4 * i + 1] = c4.y; Cv[4 * i + 2] = c4.z; Cv[4 * i + 3] = c4.w;
          }
        }
        const float uu = g.u[k];
        const float x = g.dl[k] + bias;
        const float dt = softplus ? softplus_fwd<sizeof(T) == 4>(x) : x;
        const float du = dt * uu;
        float y0 = 0.f, y1 = 0.f;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          const float a = ex2(dt * kA[i]);
          h[i] = fmaf(a, h[i], du * Bv[i]);
          if (i & 1) y1 = fmaf(Cv[i], h[i], y1); else y0 = fmaf(Cv[i], h[i], y0);
        }
        float y = y0 + y1;
        if (LPC >= 2) y += __shfl_xor_sync(0xffffffffu, y, 1);
        if (LPC >= 4) y += __shfl_xor_sync(0xffffffffu, y, 2);
        y = fmaf(Dsk, uu, y);
        const int64_t l = time_of(s);
        if (MODE == MODE_STASH) {
          if (sg == 0 && dvalid) Elem<T>::st(outp + l * o_sl, y);
        } else {
          float tot = (MODE == MODE_COMBINE) ? (y + g.st[k]) : y;
          if (sg == 0 && dvalid) {
            if (outprep) Elem<T>::st(outprep + l * op_sl, tot);
            if (has_z) { const float zz = g.z[k]; tot *= zz * sigmoidf_fast(zz); }
            Elem<T>::st(outp + l * o_sl, tot * scale);
          }
        }
      }
    }
  }

  __device__ __forceinline__ void save_ckpt(int j) const {
    if (ckp != nullptr && dvalid) {
      float4* dst = reinterpret_cast<float4*>(ckp + (int64_t)j * 16 + sg * NS);
#pragma unroll
      for (int i = 0; i < NS / 4; ++i) dst[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
    }
  }

  // Processes steps [s_begin, s_end) in groups of kGroup; checkpoint index of the first group is j0.
  template <int MODE>
  __device__ __forceinline__ void run_range(int s_begin, int s_end, int j0) {
    if (s_begin >= s_end) return;
    FwdGroup<T> ga, gb;
    load_group<MODE>(ga, s_begin, s_end);
    int j = j0;
    for (int s0 = s_begin; s0 < s_end; s0 += 2 * kGroup) {
      // ---- group A
      int tile_start = s_begin + ((s0 - s_begin) / kBcTile) * kBcTile;
      if (s0 == tile_start) stage_bc(tile_start, s_end);
      load_group<MODE>(gb, s0 + kGroup, s_end);
      save_ckpt(j++);
      compute_group<MODE>(ga, s0, s_end, tile_start);
      // ---- group B
      const int s1 = s0 + kGroup;
      if (s1 < s_end) {
        tile_start = s_begin + ((s1 - s_begin) / kBcTile) * kBcTile;
        if (s1 == tile_start) stage_bc(tile_start, s_end);
        load_group<MODE>(ga, s1 + kGroup, s_end);
        save_ckpt(j++);
        compute_group<MODE>(gb, s1, s_end, tile_start);
      }
    }
  }
};

template <typename T, int LPC, bool BC_CONST>
__global__ void __launch_bounds__(64) scan_fwd_kernel(const cm_scan_fwd_args p) {
  using Ctx = FwdCtx<T, LPC, BC_CONST>;
  constexpr int NS = Ctx::NS, CPW = Ctx::CPW;
  __shared__ __align__(16) float bc_smem[2][kBcTile * kBcPitch];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const cm_scan_dir& dp = (warp == 0) ? p.dir[0] : p.dir[1];
  const int b = blockIdx.y;
  const int cl = lane / LPC;
  int d = blockIdx.x * CPW + cl;

  Ctx c;
  c.L = p.seqlen;
  c.dstate = p.dstate;
  c.sg = lane % LPC;
  c.lane = lane;
  c.rev = dp.reverse != 0;
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
    c.kA[i] = (n < p.dstate) ? __ldg(dp.A + d * dp.A_sd + n * dp.A_sn) * kLog2e : 0.f;
    c.h[i] = 0.f;
    c.Bc[i] = 0.f;
    c.Cc[i] = 0.f;
    if (BC_CONST && n < p.dstate) {
      c.Bc[i] = __ldg(static_cast<const float*>(dp.Bm.ptr) + d * dp.Bm.sb + n * dp.Bm.sd);   // constants are fp32
      c.Cc[i] = __ldg(static_cast<const float*>(dp.Cm.ptr) + d * dp.Cm.sb + n * dp.Cm.sd);
    }
  }
  c.up = static_cast<const T*>(dp.u.ptr) + b * dp.u.sb + d * dp.u.sd;
  c.u_sl = dp.u.sl;
  c.dlp = static_cast<const T*>(dp.delta.ptr) + b * dp.delta.sb + d * dp.delta.sd;
  c.dl_sl = dp.delta.sl;
  c.zp = c.has_z ? static_cast<const T*>(p.z.ptr) + b * p.z.sb + d * p.z.sd : nullptr;
  c.z_sl = p.z.sl;
  if (!BC_CONST) {
    c.Bp = static_cast<const T*>(dp.Bm.ptr) + b * dp.Bm.sb;
    c.Cp = static_cast<const T*>(dp.Cm.ptr) + b * dp.Cm.sb;
  } else {
    c.Bp = nullptr;
    c.Cp = nullptr;
  }
  c.B_sd = dp.Bm.sd; c.B_sl = dp.Bm.sl; c.C_sd = dp.Cm.sd; c.C_sl = dp.Cm.sl;
  c.outp = static_cast<T*>(p.out.ptr) + b * p.out.sb + d * p.out.sd;
  c.o_sl = p.out.sl;
  c.outprep = p.out_pre.ptr ? static_cast<T*>(p.out_pre.ptr) + b * p.out_pre.sb + d * p.out_pre.sd : nullptr;
  c.op_sl = p.out_pre.sl;
  c.ckp = dp.ckpt ? dp.ckpt + b * dp.ckpt_sb + d * dp.ckpt_sd : nullptr;
  c.bc = bc_smem[warp];

  const int L = p.seqlen;
  if (p.ndir == 1) {
    c.template run_range<MODE_UNI>(0, L, 0);
  } else {
    const int s1 = cm_first_range(L, 2, dp.reverse);
    c.template run_range<MODE_STASH>(0, s1, 0);
    __syncthreads();  // partner's stash for the other half is now visible (same CTA, ld.global.cg)
    c.template run_range<MODE_COMBINE>(s1, L, cm_ceil_div(s1, kGroup));
  }

  if (dp.last_state != nullptr && c.dvalid) {
    float* ls = dp.last_state + b * dp.ls_sb + d * dp.ls_sd;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int n = c.sg * NS + i;
      if (n < p.dstate) ls[n * dp.ls_sn] = c.h[i];
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
  if (lpc == 0) lpc = cm_scan_pick_lanes(a.batch, a.dim, a.ndir);
  if (lpc != 1 && lpc != 2 && lpc != 4) return CM_ERR_BAD_ARG;
  const bool bcc = a.dir[0].bc_const != 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a.dtype) {
    case CM_F32: return cm::launch_fwd_t<float>(a, lpc, bcc, st);
    case CM_BF16: return cm::launch_fwd_t<__nv_bfloat16>(a, lpc, bcc, st);
    default: return cm::launch_fwd_t<__half>(a, lpc, bcc, st);
  }
}
