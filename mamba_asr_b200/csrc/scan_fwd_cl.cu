// Selective-scan forward, channel-last fast path for sm_100a.
//
// Same mathematics, lane mapping, bidirectional stash/combine protocol and checkpoint contract as scan_fwd.cu (the
// generic-stride kernel; see its header).  The difference is how operands reach the lanes: every input of a 4-step
// group - u, delta, z, the partner's stash and the B|C rows - is copied global -> shared with 16-byte cp.async
// (LDGSTS, L2-only) THREE groups ahead of the math, through a 4-stage ring per warp.  No prefetch registers, no
// unpack instruction behind a load, ~12 steps of latency cover; the freed registers let 8+ two-warp CTAs share an SM.
//
// Requirements (checked by the launcher, which otherwise falls back to scan_fwd.cu): unit channel stride and
// 16-byte aligned rows for u, delta, z, out, B, C; dim a multiple of the warp's channel count; dstate == 16;
// input-dependent B/C.
#include <type_traits>

#include "common.cuh"

namespace cm {

constexpr int kGrp = 4;                                     // steps per staged group
constexpr int kStages = 4;                                  // ring depth (prefetch distance = 3 groups)
constexpr int kCkGroups = CM_SCAN_CKPT_STEPS / kGrp;
constexpr int kBcP = 36;

__device__ __forceinline__ void cpa16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cpa_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int ROWS, int ROW_BYTES>
__device__ __forceinline__ void cpa_tile(void* dst, const char* src0, int64_t stride, int nvalid, int lane) {
  constexpr int CPR = ROW_BYTES / 16, N = ROWS * CPR;
  static_assert(ROW_BYTES % 16 == 0, "rows must be whole 16-byte chunks");
  const uint32_t d0 = sm_u32(dst);
#pragma unroll
  for (int c0 = 0; c0 < N; c0 += 32) {
    const int c = c0 + lane;
    const int row = c / CPR, col = c % CPR;
    if (c < N && row < nvalid) cpa16(d0 + row * ROW_BYTES + col * 16, src0 + row * stride + col * 16);
  }
}

template <typename T> __device__ __forceinline__ float sld(const T& v);
template <> __device__ __forceinline__ float sld<float>(const float& v) { return v; }
template <> __device__ __forceinline__ float sld<__nv_bfloat16>(const __nv_bfloat16& v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float sld<__half>(const __half& v) { return __half2float(v); }

template <typename T, int LPC>
struct FwdClSmem {
  static constexpr int CPW = 32 / LPC;
  struct Stage {
    T u[kGrp][CPW], dl[kGrp][CPW], z[kGrp][CPW], st[kGrp][CPW];
    T bc[kGrp][32];            // B (0..15) | C (16..31)
  };
  Stage ring[2][kStages];      // per warp (direction)
  float bcf[2][kGrp][kBcP];    // fp32 B/C rows of the group being computed
};

enum { FM_UNI = 0, FM_STASH = 1, FM_COMBINE = 2 };

#ifndef CM_FWDCL_MINB
#define CM_FWDCL_MINB 7
#endif
template <typename T, int LPC>
__global__ void __launch_bounds__(64, CM_FWDCL_MINB) scan_fwd_cl_kernel(const __grid_constant__ cm_scan_fwd_args p) {
  constexpr int NS = 16 / LPC, CPW = 32 / LPC, NP = NS / 2;
  constexpr int ES = (int)sizeof(T);
  constexpr int ROWB = CPW * ES;
  using Smem = FwdClSmem<T, LPC>;
  __shared__ __align__(16) Smem sm;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const cm_scan_dir& dp = p.dir[warp];
  const int b = blockIdx.y;
  const int cl = lane / LPC, sg = lane % LPC;
  const int d0 = blockIdx.x * CPW, d = d0 + cl;
  const int L = p.seqlen;
  const bool rev = dp.reverse != 0;
  const bool softplus = (p.flags & CM_FLAG_DELTA_SOFTPLUS) != 0;
  const bool has_z = p.z.ptr != nullptr;
  const float scale = p.out_scale;
  const float Dsk = dp.Dskip ? __ldg(dp.Dskip + d) : 0.f;
  const float bias = dp.delta_bias ? __ldg(dp.delta_bias + d) : 0.f;

  float2 kA2[NP], h2[NP];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const float ka = __ldg(dp.A + d * dp.A_sd + (sg * NS + i) * dp.A_sn) * kLog2e;
    if (i & 1) { kA2[i / 2].y = ka; h2[i / 2].y = 0.f; } else { kA2[i / 2].x = ka; h2[i / 2].x = 0.f; }
  }

  // tile sources (byte pointers at processed step 0 of the warp's first channel) and signed byte strides per step
  const int64_t l0 = rev ? (L - 1) : 0, sgn = rev ? -1 : 1;
  auto base = [&](const cm_tensor3& t) { return static_cast<const char*>(t.ptr) + (b * t.sb + d0 * t.sd + l0 * t.sl) * ES; };
  const char* u0 = base(dp.u);
  const char* dl0 = base(dp.delta);
  const char* z0 = has_z ? base(p.z) : nullptr;
  const char* o0 = base(p.out);
  const int64_t su = sgn * dp.u.sl * ES, sdl = sgn * dp.delta.sl * ES, sz = sgn * p.z.sl * ES, so = sgn * p.out.sl * ES;
  const char* B0 = static_cast<const char*>(dp.Bm.ptr) + (b * dp.Bm.sb + l0 * dp.Bm.sl) * ES;
  const char* C0 = static_cast<const char*>(dp.Cm.ptr) + (b * dp.Cm.sb + l0 * dp.Cm.sl) * ES;
  const int64_t sB = sgn * dp.Bm.sl * ES, sC = sgn * dp.Cm.sl * ES;
  // this lane's output rows
  T* outp = static_cast<T*>(p.out.ptr) + b * p.out.sb + d * p.out.sd + l0 * p.out.sl;
  T* prep = p.out_pre.ptr ? static_cast<T*>(p.out_pre.ptr) + b * p.out_pre.sb + d * p.out_pre.sd + l0 * p.out_pre.sl : nullptr;
  const int so_e = (int)(sgn * p.out.sl), sp_e = (int)(sgn * p.out_pre.sl);   // 32-bit: one IMAD.WIDE per address
  float* ckp = dp.ckpt ? dp.ckpt + b * dp.ckpt_sb + d * dp.ckpt_sd : nullptr;
  // Pin the per-step scalars in registers: without this ptxas re-derives them from the parameter block inside the
  // step loop (S2R tid -> warp -> LDC p.dir[warp]...), ~25 extra integer instructions per step.
  float scale_r = scale, Dsk_r = Dsk, bias_r = bias;
  int so_r = so_e, sp_r = sp_e;
  asm volatile("" : "+l"(outp), "+l"(prep), "+r"(so_r), "+r"(sp_r), "+f"(scale_r), "+f"(Dsk_r), "+f"(bias_r));

  const int s1 = cm_first_range(L, p.ndir, dp.reverse);
  const int nrange = (p.ndir == 2) ? 2 : 1;
#pragma unroll 1
  for (int range = 0; range < nrange; ++range) {
    if (range == 1) __syncthreads();   // partner's stash for the other half is complete (read back through L2)
    const int mode = (p.ndir == 1) ? FM_UNI : (range == 0 ? FM_STASH : FM_COMBINE);
    const int s_begin = range == 0 ? 0 : s1, s_end = range == 0 ? s1 : L;
    const int j0 = range == 0 ? 0 : cm_ceil_div(s1, CM_SCAN_CKPT_STEPS);
    const int n = s_end - s_begin;
    if (n <= 0) continue;
    const int ngroup = cm_ceil_div(n, kGrp);
    const bool need_z = has_z && mode != FM_STASH;
    const bool need_st = mode == FM_COMBINE;
    const bool stash_mode = mode == FM_STASH;
    const bool pre_ok = (sg == 0) && !stash_mode && prep != nullptr;

    auto issue = [&](int g) {
      if (g < ngroup) {
        const int s0 = s_begin + g * kGrp;
        const int nvalid = min(kGrp, s_end - s0);
        typename Smem::Stage& S = sm.ring[warp][g % kStages];
        cpa_tile<kGrp, ROWB>(S.u, u0 + s0 * su, su, nvalid, lane);
        cpa_tile<kGrp, ROWB>(S.dl, dl0 + s0 * sdl, sdl, nvalid, lane);
        if (need_z) cpa_tile<kGrp, ROWB>(S.z, z0 + s0 * sz, sz, nvalid, lane);
        if (need_st) cpa_tile<kGrp, ROWB>(S.st, o0 + s0 * so, so, nvalid, lane);
        constexpr int CB = 16 * ES / 16;                 // 16-byte chunks of one 16-element B (or C) row
        {
          const int c = lane % (kGrp * CB), which = lane / (kGrp * CB);
          const int row = c / CB, col = c % CB;
          if (kGrp * CB * 2 <= 32) {
            if (which < 2 && row < nvalid)
              cpa16(sm_u32(&S.bc[row][which * 16]) + col * 16, (which ? C0 + (s0 + row) * sC : B0 + (s0 + row) * sB) + col * 16);
          } else {
            if (lane < kGrp * CB && row < nvalid) {
              cpa16(sm_u32(&S.bc[row][0]) + col * 16, B0 + (s0 + row) * sB + col * 16);
              cpa16(sm_u32(&S.bc[row][16]) + col * 16, C0 + (s0 + row) * sC + col * 16);
            }
          }
        }
      }
      cpa_commit();                                      // always commit: keeps the group count uniform
    };

#pragma unroll
    for (int i = 0; i < kStages - 1; ++i) issue(i);
#pragma unroll 1
    for (int g = 0; g < ngroup; ++g) {
      issue(g + kStages - 1);
      cpa_wait<kStages - 1>();
      __syncwarp();
      const typename Smem::Stage& S = sm.ring[warp][g % kStages];
      const int s0 = s_begin + g * kGrp;
      const int nvalid = min(kGrp, s_end - s0);
      float (*bcf)[kBcP] = sm.bcf[warp];
#pragma unroll
      for (int k = 0; k < kGrp; ++k) bcf[k][lane] = sld<T>(S.bc[k][lane]);
      if (ckp != nullptr && (g % kCkGroups) == 0) {
        float4* dst = reinterpret_cast<float4*>(ckp + (int64_t)(j0 + g / kCkGroups) * 16 + sg * NS);
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) dst[i] = make_float4(h2[2 * i].x, h2[2 * i].y, h2[2 * i + 1].x, h2[2 * i + 1].y);
      }
      __syncwarp();
      auto group_body = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;   // all kGrp steps exist: no per-step predicates
      // ---- pre-phase: per-step scalars of the whole group (four independent softplus / gate chains interleave)
        float uu[kGrp], dtv[kGrp], stv[kGrp], gatev[kGrp];
#pragma unroll
        for (int k = 0; k < kGrp; ++k) {
          uu[k] = sld<T>(S.u[k][cl]);
          const float x = sld<T>(S.dl[k][cl]) + bias_r;
          dtv[k] = softplus ? softplus_fwd<sizeof(T) == 4>(x) : x;
          stv[k] = need_st ? sld<T>(S.st[k][cl]) : 0.f;
          gatev[k] = 1.f;
          if (need_z) { const float zz = sld<T>(S.z[k][cl]); gatev[k] = zz * sigmoid_sel<sizeof(T) == 4>(zz); }
        }
        // ---- recurrence: each step is written in pipeline order - operand rows, all exponent arguments, all 16 MUFU,
        // all input products, then the state updates and the output contraction - so that no instruction sits right
        // behind the MUFU / LDS result it consumes
#pragma unroll
        for (int k = 0; k < kGrp; ++k) {
          if (FULL || k < nvalid) {
            const float dt = dtv[k], du = dt * uu[k];
            const float2 dt2 = make_float2(dt, dt), du2 = make_float2(du, du);
            float4 b4[NS / 4], c4[NS / 4];
            const float4* rb = reinterpret_cast<const float4*>(&bcf[k][sg * NS]);
            const float4* rc = reinterpret_cast<const float4*>(&bcf[k][16 + sg * NS]);
#pragma unroll
            for (int q = 0; q < NS / 4; ++q) b4[q] = rb[q];
            float2 a2[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) a2[i] = fmul2(dt2, kA2[i]);
#pragma unroll
            for (int q = 0; q < NS / 4; ++q) c4[q] = rc[q];
#pragma unroll
            for (int i = 0; i < NP; ++i) a2[i] = make_float2(ex2(a2[i].x), ex2(a2[i].y));
            float2 ub[NP];
#pragma unroll
            for (int q = 0; q < NS / 4; ++q) {
              ub[2 * q] = fmul2(du2, make_float2(b4[q].x, b4[q].y));
              ub[2 * q + 1] = fmul2(du2, make_float2(b4[q].z, b4[q].w));
            }
#pragma unroll
            for (int i = 0; i < NP; ++i) h2[i] = ffma2(a2[i], h2[i], ub[i]);
            float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < NS / 4; ++q) {
              ya = ffma2(make_float2(c4[q].x, c4[q].y), h2[2 * q], ya);
              yb = ffma2(make_float2(c4[q].z, c4[q].w), h2[2 * q + 1], yb);
            }
            const float2 ys = fadd2(ya, yb);
            float y = ys.x + ys.y;
            if (LPC >= 2) y += __shfl_xor_sync(0xffffffffu, y, 1);
            if (LPC >= 4) y += __shfl_xor_sync(0xffffffffu, y, 2);
            y = fmaf(Dsk_r, uu[k], y);
            const float tot = y + stv[k];
            const float val = stash_mode ? y : tot * gatev[k] * scale_r;
            const int s = s0 + k;
            if (pre_ok) Elem<T>::st(prep + (int64_t)s * sp_r, tot);
            if (sg == 0) Elem<T>::st(outp + (int64_t)s * so_r, val);
          }
        }
      };
      if (nvalid == kGrp) group_body(std::true_type{}); else group_body(std::false_type{});
      __syncwarp();   // stage g % kStages is refilled by the next iteration's issue
    }
    cpa_wait<0>();
  }

  if (dp.last_state != nullptr) {
    float* ls = dp.last_state + b * dp.ls_sb + d * dp.ls_sd;
#pragma unroll
    for (int i = 0; i < NS; ++i) ls[(sg * NS + i) * dp.ls_sn] = (i & 1) ? h2[i / 2].y : h2[i / 2].x;
  }
}

static bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
template <typename T>
static bool tok(const cm_tensor3& t) {
  const int64_t es = sizeof(T);
  return t.ptr != nullptr && t.sd == 1 && a16(t.ptr) && (t.sb * es) % 16 == 0 && (t.sl * es) % 16 == 0;
}

template <typename T>
static bool fwd_cl_ok(const cm_scan_fwd_args& a, int lpc) {
  if (a.dstate != 16 || a.dim % (32 / lpc) != 0) return false;
  if (!tok<T>(a.out)) return false;
  if (a.z.ptr != nullptr && !tok<T>(a.z)) return false;
  if (a.out_pre.ptr != nullptr && a.out_pre.sd != 1) return false;
  for (int r = 0; r < a.ndir; ++r) {
    const cm_scan_dir& d = a.dir[r];
    if (d.bc_const) return false;
    if (!tok<T>(d.u) || !tok<T>(d.delta) || !tok<T>(d.Bm) || !tok<T>(d.Cm)) return false;
    if (d.ckpt != nullptr && (!a16(d.ckpt) || (d.ckpt_sb % 4) != 0 || (d.ckpt_sd % 4) != 0)) return false;
  }
  return true;
}

template <typename T>
static int launch_fwd_cl_t(const cm_scan_fwd_args& a, int lpc, cudaStream_t st) {
  const dim3 block(32 * a.ndir);
  if (lpc == 1) scan_fwd_cl_kernel<T, 1><<<dim3(a.dim / 32, a.batch), block, 0, st>>>(a);
  else if (lpc == 2) scan_fwd_cl_kernel<T, 2><<<dim3(a.dim / 16, a.batch), block, 0, st>>>(a);
  else scan_fwd_cl_kernel<T, 4><<<dim3(a.dim / 8, a.batch), block, 0, st>>>(a);
  CM_LAUNCH_CHECK();
  return 0;
}

// returns 1 if launched (result in *rc), 0 if the fast path does not apply
int scan_fwd_try_channel_last(const cm_scan_fwd_args& a, int lpc, cudaStream_t st, int* rc) {
  switch (a.dtype) {
    case CM_F32:
      if (!fwd_cl_ok<float>(a, lpc)) return 0;
      *rc = launch_fwd_cl_t<float>(a, lpc, st);
      return 1;
    case CM_BF16:
      if (!fwd_cl_ok<__nv_bfloat16>(a, lpc)) return 0;
      *rc = launch_fwd_cl_t<__nv_bfloat16>(a, lpc, st);
      return 1;
    default:
      if (!fwd_cl_ok<__half>(a, lpc)) return 0;
      *rc = launch_fwd_cl_t<__half>(a, lpc, st);
      return 1;
  }
}

}  // namespace cm
