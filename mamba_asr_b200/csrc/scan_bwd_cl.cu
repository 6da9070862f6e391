// Selective-scan backward, channel-last fast path for sm_100a: every input of a tile is staged global -> shared
// with 16-byte cp.async (LDGSTS) one tile ahead of the math, so no prefetch registers, no conversion instruction
// behind a load, and the next tile's HBM latency is hidden behind the current tile's ~2 k issue slots.
//
// Same mathematics, lane mapping, checkpoint / partial-sum contracts as scan_bwd.cu (which stays the generic-stride
// kernel); see that file's header.  Requirements checked by the launcher (cm_scan_bwd falls back otherwise):
// unit channel stride and 16-byte aligned rows for u, delta, dout, z, out_pre, B, C; dim a multiple of the warp's
// channel count; dstate == 16; input-dependent B/C.
//
// Shared memory per warp (bf16, LPC = 1): 2 stages x (5 x 512 B activations + 512 B B|C rows + 2 KB checkpoint)
// + 16 KB recomputed states + 1.1 KB fp32 B/C rows + 4.5 KB reduce scratch = 32 KB.
#include <type_traits>

#include "common.cuh"

namespace cm {

constexpr int kTileC = CM_SCAN_CKPT_STEPS;
constexpr int kPitchC = 36;

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ROWS x ROW_BYTES tile, rows `stride` bytes apart in global memory (signed), rows >= nvalid skipped
template <int ROWS, int ROW_BYTES>
__device__ __forceinline__ void cp_tile(void* dst, const char* src0, int64_t stride, int nvalid, int lane) {
  constexpr int CPR = ROW_BYTES / 16, N = ROWS * CPR;
  static_assert(ROW_BYTES % 16 == 0, "rows must be whole 16-byte chunks");
  const uint32_t d0 = smem_u32(dst);
#pragma unroll
  for (int c0 = 0; c0 < N; c0 += 32) {
    const int c = c0 + lane;
    const int row = c / CPR, col = c % CPR;
    if (c < N && row < nvalid) cp_async16(d0 + row * ROW_BYTES + col * 16, src0 + row * stride + col * 16);
  }
}

template <typename T, int LPC>
struct BwdClSmem {
  static constexpr int NS = 16 / LPC;
  static constexpr int CPW = 32 / LPC;
  struct Stage {
    T u[kTileC][CPW], dl[kTileC][CPW], go[kTileC][CPW], z[kTileC][CPW], pre[kTileC][CPW];
    T bc[kTileC][32];          // B (0..15) | C (16..31)
    float ck[CPW][16];         // checkpoint of the tile, per channel
  };
  Stage st[2];
  float4 h[kTileC][NS / 4][32];
  float bcf[kTileC][kPitchC];
  float red[32][2 * NS + 4];
};

template <int NP>
__device__ __forceinline__ void lds_row2(const float* src, float2 (&dst)[NP]) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < NP / 2; ++i) {
    const float4 v = s4[i];
    dst[2 * i] = make_float2(v.x, v.y);
    dst[2 * i + 1] = make_float2(v.z, v.w);
  }
}

template <typename T>
__device__ __forceinline__ float ldsf(const T& v);
template <>
__device__ __forceinline__ float ldsf<float>(const float& v) { return v; }
template <>
__device__ __forceinline__ float ldsf<__nv_bfloat16>(const __nv_bfloat16& v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float ldsf<__half>(const __half& v) { return __half2float(v); }

#ifdef CM_BWD_CAP
#define CM_BWD_LB __launch_bounds__(32, (LPC == 1 ? 7 : (LPC == 2 ? 12 : 16)))
#else
#define CM_BWD_LB __launch_bounds__(32)
#endif
template <typename T, int LPC>
__global__ void CM_BWD_LB scan_bwd_cl_kernel(const __grid_constant__ cm_scan_bwd_args p) {
  constexpr int NS = 16 / LPC, CPW = 32 / LPC, NP = NS / 2;
  constexpr int ES = (int)sizeof(T);
  constexpr int ROWB = CPW * ES;          // bytes of one staged activation row
  using Smem = BwdClSmem<T, LPC>;
  __shared__ __align__(16) Smem sm;

  const int lane = threadIdx.x;
  const int r = blockIdx.z;
  const cm_scan_bwd_dir& bd = p.dir[r];
  const cm_scan_dir& dp = bd.in;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int cl = lane / LPC, sg = lane % LPC;
  const int d0 = slab * CPW;
  const int d = d0 + cl;                                   // dim % CPW == 0 on this path: always valid
  const int L = p.seqlen;
  const bool rev = dp.reverse != 0;
  const bool softplus = (p.flags & CM_FLAG_DELTA_SOFTPLUS) != 0;
  const bool has_z = p.z.ptr != nullptr;
  const bool do_dz = has_z && (r == 0);
  const float scale = p.out_scale;
  const float Dsk = dp.Dskip ? __ldg(dp.Dskip + d) : 0.f;
  const float bias = dp.delta_bias ? __ldg(dp.delta_bias + d) : 0.f;

  float2 kA2[NP], mu2[NP], dA2[NP];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const float ka = __ldg(dp.A + d * dp.A_sd + (sg * NS + i) * dp.A_sn) * kLog2e;
    if (i & 1) { kA2[i / 2].y = ka; mu2[i / 2].y = 0.f; dA2[i / 2].y = 0.f; }
    else { kA2[i / 2].x = ka; mu2[i / 2].x = 0.f; dA2[i / 2].x = 0.f; }
  }
  float dD_acc = 0.f, dbias_acc = 0.f;

  // tile sources: byte pointers at processed step 0 of this warp's first channel; signed byte strides per step
  const int64_t l0 = rev ? (L - 1) : 0;
  const int64_t sgn = rev ? -1 : 1;
  auto base = [&](const cm_tensor3& t) {
    return static_cast<const char*>(t.ptr) + (b * t.sb + d0 * t.sd + l0 * t.sl) * ES;
  };
  const char* u0 = base(dp.u);
  const char* dl0 = base(dp.delta);
  const char* go0 = base(p.dout);
  const char* z0 = has_z ? base(p.z) : nullptr;
  const char* pre0 = do_dz ? base(p.out_pre) : nullptr;
  const int64_t su = sgn * dp.u.sl * ES, sdl = sgn * dp.delta.sl * ES, sgo = sgn * p.dout.sl * ES;
  const int64_t sz = sgn * p.z.sl * ES, spre = sgn * p.out_pre.sl * ES;
  const char* B0 = static_cast<const char*>(dp.Bm.ptr) + (b * dp.Bm.sb + l0 * dp.Bm.sl) * ES;
  const char* C0 = static_cast<const char*>(dp.Cm.ptr) + (b * dp.Cm.sb + l0 * dp.Cm.sl) * ES;
  const int64_t sB = sgn * dp.Bm.sl * ES, sC = sgn * dp.Cm.sl * ES;
  const char* ck0 = reinterpret_cast<const char*>(dp.ckpt + b * dp.ckpt_sb + (int64_t)d0 * dp.ckpt_sd);
  const int64_t sck = dp.ckpt_sd * 4;                      // bytes between consecutive channels' checkpoint rows

  // outputs (direct stores)
  T* dzp = do_dz ? static_cast<T*>(p.dz.ptr) + b * p.dz.sb + d * p.dz.sd + l0 * p.dz.sl : nullptr;
  T* dup = static_cast<T*>(bd.du.ptr) + b * bd.du.sb + d * bd.du.sd + l0 * bd.du.sl;
  T* ddp = static_cast<T*>(bd.ddelta.ptr) + b * bd.ddelta.sb + d * bd.ddelta.sd + l0 * bd.ddelta.sl;
  const int sdz = (int)(sgn * p.dz.sl), sdu = (int)(sgn * bd.du.sl), sdd = (int)(sgn * bd.ddelta.sl);
  float* partp = bd.dBC_part + ((int64_t)b * gridDim.x + slab) * (int64_t)L * 32 + l0 * 32;
  const int spart = (int)sgn * 32;

  // Pin the per-step scalars in registers (ptxas otherwise re-derives them from the parameter block per step).
  float scale_r = scale, Dsk_r = Dsk, bias_r = bias;
  int sdz_r = sdz, sdu_r = sdu, sdd_r = sdd, spart_r = spart;
  asm volatile("" : "+l"(dzp), "+l"(dup), "+l"(ddp), "+l"(partp), "+r"(sdz_r), "+r"(sdu_r), "+r"(sdd_r), "+r"(spart_r),
               "+f"(scale_r), "+f"(Dsk_r), "+f"(bias_r));

  const int s1 = cm_first_range(L, p.ndir, dp.reverse);
  // flat tile index over both ranges, walked from the last tile to the first
  const int nt0 = cm_ceil_div(s1, kTileC);                 // tiles of range [0, s1)
  const int nt1 = (p.ndir == 2) ? cm_ceil_div(L - s1, kTileC) : 0;
  const int ntile = nt0 + nt1;

  auto tile_span = [&](int t, int& s0, int& nvalid) {
    if (t < nt0) { s0 = t * kTileC; nvalid = min(kTileC, s1 - s0); }
    else { s0 = s1 + (t - nt0) * kTileC; nvalid = min(kTileC, L - s0); }
  };
  auto issue = [&](int t, int stage) {
    int s0, nvalid;
    tile_span(t, s0, nvalid);
    typename Smem::Stage& S = sm.st[stage];
    cp_tile<kTileC, ROWB>(S.u, u0 + s0 * su, su, nvalid, lane);
    cp_tile<kTileC, ROWB>(S.dl, dl0 + s0 * sdl, sdl, nvalid, lane);
    cp_tile<kTileC, ROWB>(S.go, go0 + s0 * sgo, sgo, nvalid, lane);
    if (has_z) cp_tile<kTileC, ROWB>(S.z, z0 + s0 * sz, sz, nvalid, lane);
    if (do_dz) cp_tile<kTileC, ROWB>(S.pre, pre0 + s0 * spre, spre, nvalid, lane);
    // B | C rows: 16 elements each -> (16 * ES / 16) chunks; lanes [0, 8*CB) take B, the next 8*CB take C
    constexpr int CB = 16 * ES / 16;
    {
      const int c = lane % (kTileC * CB), which = lane / (kTileC * CB);
      const int row = c / CB, col = c % CB;
      if (kTileC * CB * 2 <= 32) {
        if (which < 2 && row < nvalid)
          cp_async16(smem_u32(&S.bc[row][which * 16]) + col * 16, (which ? C0 + (s0 + row) * sC : B0 + (s0 + row) * sB) + col * 16);
      } else {
        // fp32: 8 rows x 4 chunks per matrix = 32 chunks each -> one pass for B, one for C
        if (row < nvalid) {
          cp_async16(smem_u32(&S.bc[row][0]) + col * 16, B0 + (s0 + row) * sB + col * 16);
          cp_async16(smem_u32(&S.bc[row][16]) + col * 16, C0 + (s0 + row) * sC + col * 16);
        }
      }
    }
    // checkpoint of the tile: 64 B per channel
    {
      const char* src = ck0 + (int64_t)t * 64;
#pragma unroll
      for (int c0 = 0; c0 < CPW * 4; c0 += 32) {
        const int c = c0 + lane;
        if (c < CPW * 4) cp_async16(smem_u32(&S.ck[c / 4][0]) + (c % 4) * 16, src + (c / 4) * sck + (c % 4) * 16);
      }
    }
    cp_async_commit();
  };

  if (ntile > 0) issue(ntile - 1, (ntile - 1) & 1);
#pragma unroll 1
  for (int t = ntile - 1; t >= 0; --t) {
    int s0, nvalid;
    tile_span(t, s0, nvalid);
    if (t > 0) {
      issue(t - 1, (t - 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const typename Smem::Stage& S = sm.st[t & 1];

    auto tile_body = [&](auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;   // all 8 steps exist: no per-step predicates
      // fp32 B/C rows for the tile
      {
#pragma unroll
        for (int k = 0; k < kTileC; ++k) sm.bcf[k][lane] = (FULL || k < nvalid) ? ldsf<T>(S.bc[k][lane]) : 0.f;
      }
      // per-step scalars of this lane's channel
      float uu[kTileC], xx[kTileC], gg[kTileC];
#pragma unroll
      for (int k = 0; k < kTileC; ++k) {
        uu[k] = 0.f; xx[k] = 0.f; gg[k] = 0.f;
        if (FULL || k < nvalid) {
          uu[k] = ldsf<T>(S.u[k][cl]);
          xx[k] = ldsf<T>(S.dl[k][cl]) + bias_r;
          const float dov = ldsf<T>(S.go[k][cl]) * scale_r;
          if (has_z) {
            const float zz = ldsf<T>(S.z[k][cl]);
            const float sig = sigmoidf_fast(zz);
            gg[k] = dov * zz * sig;
            if (do_dz && sg == 0)
              Elem<T>::st(dzp + (int64_t)(s0 + k) * sdz_r, dov * ldsf<T>(S.pre[k][cl]) * sig * fmaf(zz, 1.f - sig, 1.f));
          } else {
            gg[k] = dov;
          }
        }
      }
      float2 h2[NP];
      lds_row2<NP>(&S.ck[cl][sg * NS], h2);
      __syncwarp();   // bcf visible

      // ---- phase 1: recompute the tile's states from its checkpoint ------------------------------------
      float dtv[kTileC];
#pragma unroll
      for (int k = 0; k < kTileC; ++k) {
        dtv[k] = 0.f;
        if (FULL || k < nvalid) {
          const float dt = softplus ? softplus_fwd<sizeof(T) == 4>(xx[k]) : xx[k];
          dtv[k] = dt;
          const float du_ = dt * uu[k];
          const float2 dt2 = make_float2(dt, dt), du2 = make_float2(du_, du_);
          float2 B2[NP];
          lds_row2<NP>(&sm.bcf[k][sg * NS], B2);
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const float2 x2 = fmul2(dt2, kA2[i]);
            const float2 a2 = make_float2(ex2(x2.x), ex2(x2.y));
            h2[i] = ffma2(a2, h2[i], fmul2(du2, B2[i]));
          }
#pragma unroll
          for (int i = 0; i < NS / 4; ++i)
            sm.h[k][i][lane] = make_float4(h2[2 * i].x, h2[2 * i].y, h2[2 * i + 1].x, h2[2 * i + 1].y);
        }
      }

      // ---- phase 2: reverse sweep ---------------------------------------------------------------------
#pragma unroll
      for (int k = kTileC - 1; k >= 0; --k) {
        if (FULL || k < nvalid) {
          const int s = s0 + k;
          const float dt = dtv[k], u_ = uu[k], g = gg[k];
          const float du_ = dt * u_;
          const float2 dt2 = make_float2(dt, dt), du2 = make_float2(du_, du_), g2 = make_float2(g, g);
          const float2 ndu2 = make_float2(-du_, -du_);
          __syncwarp();   // previous step's readers of sm.red are done
          float4* red = reinterpret_cast<float4*>(&sm.red[lane][0]);
          float2 sLB2 = make_float2(0.f, 0.f), sWA2 = make_float2(0.f, 0.f);
          // four states (two fp32 pairs) at a time: operands come in as LDS.128, dB / dC leave as STS.128
#pragma unroll
          for (int q = 0; q < NS / 4; ++q) {
            const float4 hv = sm.h[k][q][lane];
            const float4 bv = *reinterpret_cast<const float4*>(&sm.bcf[k][sg * NS + 4 * q]);
            const float4 cv = *reinterpret_cast<const float4*>(&sm.bcf[k][16 + sg * NS + 4 * q]);
            float2 dBq[2], dCq[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = 2 * q + e;
              const float2 hk = e ? make_float2(hv.z, hv.w) : make_float2(hv.x, hv.y);
              const float2 Bq = e ? make_float2(bv.z, bv.w) : make_float2(bv.x, bv.y);
              const float2 Cq = e ? make_float2(cv.z, cv.w) : make_float2(cv.x, cv.y);
              const float2 lam = ffma2(g2, Cq, mu2[i]);           // lambda = g*C + mu
              dCq[e] = fmul2(g2, hk);
              dBq[e] = fmul2(lam, du2);
              sLB2 = ffma2(lam, Bq, sLB2);
              const float2 hp = ffma2(ndu2, Bq, hk);              // = a * h_{s-1}
              const float2 w = fmul2(lam, hp);
              sWA2 = ffma2(w, kA2[i], sWA2);
              dA2[i] = ffma2(w, dt2, dA2[i]);
              const float2 x2 = fmul2(dt2, kA2[i]);
              mu2[i] = fmul2(make_float2(ex2(x2.x), ex2(x2.y)), lam);
            }
            red[q] = make_float4(dBq[0].x, dBq[0].y, dBq[1].x, dBq[1].y);
            red[NS / 4 + q] = make_float4(dCq[0].x, dCq[0].y, dCq[1].x, dCq[1].y);
          }
          float sLB = sLB2.x + sLB2.y, sWA = sWA2.x + sWA2.y;
          if (LPC >= 2) {
            sLB += __shfl_xor_sync(0xffffffffu, sLB, 1);
            sWA += __shfl_xor_sync(0xffffffffu, sWA, 1);
          }
          if (LPC >= 4) {
            sLB += __shfl_xor_sync(0xffffffffu, sLB, 2);
            sWA += __shfl_xor_sync(0xffffffffu, sWA, 2);
          }
          const float du = fmaf(g, Dsk_r, dt * sLB);
          const float ddt = fmaf(u_, sLB, kLn2 * sWA);
          const float ddl = softplus ? ddt * softplus_grad(xx[k]) : ddt;
          dD_acc = fmaf(g, u_, dD_acc);
          dbias_acc += ddl;
          if (sg == 0) {
            Elem<T>::st(dup + (int64_t)s * sdu_r, du);
            Elem<T>::st(ddp + (int64_t)s * sdd_r, ddl);
          }
          // cross-channel reduce of the 32 per-step values
          __syncwarp();
          if (LPC == 1) {
            // lane -> (quad q = lane & 7 of the 32 values, row group rg = lane >> 3): 8 LDS.128 + packed adds, then two
            // shuffle rounds fold the 4 row groups; lanes 0..7 store one coalesced 128-byte row
            const int q = lane & 7, rg = lane >> 3;
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 v = *reinterpret_cast<const float4*>(&sm.red[rg + 4 * i][4 * q]);
              a0 = fadd2(a0, make_float2(v.x, v.y));
              a1 = fadd2(a1, make_float2(v.z, v.w));
            }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              a0.x += __shfl_xor_sync(0xffffffffu, a0.x, o); a0.y += __shfl_xor_sync(0xffffffffu, a0.y, o);
              a1.x += __shfl_xor_sync(0xffffffffu, a1.x, o); a1.y += __shfl_xor_sync(0xffffffffu, a1.y, o);
            }
            if (lane < 8) *reinterpret_cast<float4*>(partp + (int64_t)s * spart_r + 4 * q) = make_float4(a0.x, a0.y, a1.x, a1.y);
          } else {
            const int n = lane & 15;
            const int col = ((lane < 16) ? 0 : NS) + (n % NS);
            const int row0 = n / NS;
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int c = 0; c < CPW; c += 2) {
              acc0 += sm.red[c * LPC + row0][col];
              acc1 += sm.red[(c + 1) * LPC + row0][col];
            }
            partp[(int64_t)s * spart_r + lane] = acc0 + acc1;
          }
        }
      }
    };
#ifdef CM_BWD_FULLSPLIT
    if (nvalid == kTileC) tile_body(std::true_type{}); else tile_body(std::false_type{});
#else
    tile_body(std::false_type{});
#endif
    __syncwarp();   // all lanes done with this stage before it is refilled two tiles later
  }

  const int64_t row = (int64_t)b * p.dim + d;
#pragma unroll
  for (int i = 0; i < NS; ++i) bd.dA_part[row * 16 + sg * NS + i] = (i & 1) ? dA2[i / 2].y : dA2[i / 2].x;
  if (sg == 0) {
    if (bd.dD_part) bd.dD_part[row] = dD_acc;
    if (bd.dbias_part) bd.dbias_part[row] = dbias_acc;
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T>
static bool t_ok(const cm_tensor3& t) {
  const int64_t es = sizeof(T);
  return t.ptr != nullptr && t.sd == 1 && al16(t.ptr) && (t.sb * es) % 16 == 0 && (t.sl * es) % 16 == 0;
}

// whether the cp.async fast path applies
template <typename T>
static bool bwd_cl_ok(const cm_scan_bwd_args& a, int lpc) {
  if (a.dstate != 16 || a.dim % (32 / lpc) != 0) return false;
  if (!t_ok<T>(a.dout)) return false;
  if (a.z.ptr != nullptr && (!t_ok<T>(a.z) || !t_ok<T>(a.out_pre))) return false;
  for (int r = 0; r < a.ndir; ++r) {
    const cm_scan_dir& d = a.dir[r].in;
    if (d.bc_const) return false;
    if (!t_ok<T>(d.u) || !t_ok<T>(d.delta) || !t_ok<T>(d.Bm) || !t_ok<T>(d.Cm)) return false;
    if (!al16(d.ckpt) || (d.ckpt_sb % 4) != 0 || (d.ckpt_sd % 4) != 0) return false;
  }
  return true;
}

template <typename T>
static int launch_bwd_cl_t(const cm_scan_bwd_args& a, int lpc, cudaStream_t st) {
  if (lpc == 1) {
    const dim3 grid(a.dim / 32, a.batch, a.ndir);
    scan_bwd_cl_kernel<T, 1><<<grid, 32, 0, st>>>(a);
  } else if (lpc == 2) {
    const dim3 grid(a.dim / 16, a.batch, a.ndir);
    scan_bwd_cl_kernel<T, 2><<<grid, 32, 0, st>>>(a);
  } else {
    const dim3 grid(a.dim / 8, a.batch, a.ndir);
    scan_bwd_cl_kernel<T, 4><<<grid, 32, 0, st>>>(a);
  }
  CM_LAUNCH_CHECK();
  return 0;
}

// returns 1 if launched, 0 if the fast path does not apply, < 0 / cudaError on failure (positive errors are offset)
int scan_bwd_try_channel_last(const cm_scan_bwd_args& a, int lpc, cudaStream_t st, int* rc) {
  switch (a.dtype) {
    case CM_F32:
      if (!bwd_cl_ok<float>(a, lpc)) return 0;
      *rc = launch_bwd_cl_t<float>(a, lpc, st);
      return 1;
    case CM_BF16:
      if (!bwd_cl_ok<__nv_bfloat16>(a, lpc)) return 0;
      *rc = launch_bwd_cl_t<__nv_bfloat16>(a, lpc, st);
      return 1;
    default:
      if (!bwd_cl_ok<__half>(a, lpc)) return 0;
      *rc = launch_bwd_cl_t<__half>(a, lpc, st);
      return 1;
  }
}

}  // namespace cm
