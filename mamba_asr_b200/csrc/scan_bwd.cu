// Selective-scan backward for sm_100a (both BiMamba directions in one launch).
//
// Replaces selective_scan_cuda.bwd (reference call sites modules/mamba/selective_scan_interface.py:67,252).
// Same lane mapping as the forward kernel (one lane per channel, or LPC lanes per channel splitting the 16
// states).  One CTA = one warp = 32/LPC channels of one batch row in one direction; grid.z = direction.
//
// The adjoint needs the forward state h_s and the reverse-time adjoint lambda_s at the same step.  Forward
// saved h every CM_SCAN_CKPT_STEPS steps; here each 8-step tile (walked last -> first) is
//   1. recomputed forward from its checkpoint, parking h_s for the 8 steps in shared memory (16 KB / warp),
//   2. swept in reverse with   lambda = g*C + mu ;  mu <- a * lambda   carried in registers,
// which never divides by the decay a = exp(Delta*A) (the reversible form is unstable when a underflows).
// Adjoint formulas: SURVEY.md section 9.2; a*h_{s-1} is obtained as h_s - Delta*u*B.
//
// dB/dC need a sum over channels.  The reference kernel uses fp32 atomics per element; here the 32 channel
// lanes of a warp are reduced through a padded shared-memory transpose (8 STS.128 + 32 LDS + 31 FADD per step)
// and written as one coalesced 128-byte row of a [batch][slab][time][32] fp32 partial tensor that
// cm_reduce_dbc() sums over slabs in a fixed order: deterministic, no atomics.
// dA, dD, d(delta_bias) are per-row register sums written once per (batch, channel) and reduced over batch
// by cm_reduce_rows().
#include <cstdlib>

#include "common.cuh"

namespace cm {

constexpr int kTile = CM_SCAN_CKPT_STEPS;
constexpr int kPitch = 36;

template <int LPC>
struct BwdSmem {
  static constexpr int NS = 16 / LPC;
  static constexpr int kRedPitch = 2 * NS + 4;
  float4 h[kTile][NS / 4][32];    // recomputed states of the tile
  float bc[kTile][kPitch];        // B (0..15) and C (16..31) of the tile's steps
  float red[32][kRedPitch];       // cross-channel reduce scratch
};

template <int NS>
__device__ __forceinline__ void load_row(const float* src, float (&dst)[NS]) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < NS / 4; ++i) {
    const float4 v = s4[i];
    dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
  }
}

template <int NP>
__device__ __forceinline__ void load_row2(const float* src, float2 (&dst)[NP]) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < NP / 2; ++i) {
    const float4 v = s4[i];
    dst[2 * i] = make_float2(v.x, v.y); dst[2 * i + 1] = make_float2(v.z, v.w);
  }
}

template <typename T, int LPC, bool BC_CONST>
__global__ void __launch_bounds__(32) scan_bwd_kernel(const cm_scan_bwd_args p) {
  constexpr int NS = 16 / LPC, CPW = 32 / LPC;
  using Smem = BwdSmem<LPC>;
  __shared__ __align__(16) Smem sm;

  const int lane = threadIdx.x;
  const int r = blockIdx.z;
  const cm_scan_bwd_dir& bd = (r == 0) ? p.dir[0] : p.dir[1];
  const cm_scan_dir& dp = bd.in;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int cl = lane / LPC, sg = lane % LPC;
  int d = slab * CPW + cl;
  const bool dvalid = d < p.dim;
  if (!dvalid) d = p.dim - 1;
  const int L = p.seqlen;
  const bool rev = dp.reverse != 0;
  const bool softplus = (p.flags & CM_FLAG_DELTA_SOFTPLUS) != 0;
  const bool has_z = p.z.ptr != nullptr;
  const bool do_dz = has_z && (r == 0);
  const float scale = p.out_scale;
  const float Dsk = dp.Dskip ? __ldg(dp.Dskip + d) : 0.f;
  const float bias = dp.delta_bias ? __ldg(dp.delta_bias + d) : 0.f;

  constexpr int NP = NS / 2;   // state pairs (FFMA2 / FMUL2)
  float2 kA2[NP], mu2[NP], dA2[NP];
  float Bc[NS], Cc[NS], dBacc[NS], dCacc[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int n = sg * NS + i;
    const float ka = (n < p.dstate) ? __ldg(dp.A + d * dp.A_sd + n * dp.A_sn) * kLog2e : 0.f;
    if (i & 1) { kA2[i / 2].y = ka; mu2[i / 2].y = 0.f; dA2[i / 2].y = 0.f; }
    else { kA2[i / 2].x = ka; mu2[i / 2].x = 0.f; dA2[i / 2].x = 0.f; }
    Bc[i] = 0.f; Cc[i] = 0.f; dBacc[i] = 0.f; dCacc[i] = 0.f;
    if (BC_CONST && n < p.dstate) {
      Bc[i] = __ldg(static_cast<const float*>(dp.Bm.ptr) + d * dp.Bm.sb + n * dp.Bm.sd);   // constants are fp32
      Cc[i] = __ldg(static_cast<const float*>(dp.Cm.ptr) + d * dp.Cm.sb + n * dp.Cm.sd);
    }
  }
  float dD_acc = 0.f, dbias_acc = 0.f;

  // pointers positioned at processed step 0; signed element strides per processed step
  using Raw = typename Elem<T>::Raw;
  const int64_t l0 = rev ? (L - 1) : 0;
  const int sgn = rev ? -1 : 1;
  const T* up = static_cast<const T*>(dp.u.ptr) + b * dp.u.sb + d * dp.u.sd + l0 * dp.u.sl;
  const T* dlp = static_cast<const T*>(dp.delta.ptr) + b * dp.delta.sb + d * dp.delta.sd + l0 * dp.delta.sl;
  const T* zp = has_z ? static_cast<const T*>(p.z.ptr) + b * p.z.sb + d * p.z.sd + l0 * p.z.sl : nullptr;
  const T* prep = do_dz ? static_cast<const T*>(p.out_pre.ptr) + b * p.out_pre.sb + d * p.out_pre.sd + l0 * p.out_pre.sl
                        : nullptr;
  const T* dop = static_cast<const T*>(p.dout.ptr) + b * p.dout.sb + d * p.dout.sd + l0 * p.dout.sl;
  T* dzp = do_dz ? static_cast<T*>(p.dz.ptr) + b * p.dz.sb + d * p.dz.sd + l0 * p.dz.sl : nullptr;
  T* dup = static_cast<T*>(bd.du.ptr) + b * bd.du.sb + d * bd.du.sd + l0 * bd.du.sl;
  T* ddp = static_cast<T*>(bd.ddelta.ptr) + b * bd.ddelta.sb + d * bd.ddelta.sd + l0 * bd.ddelta.sl;
  const int su = sgn * (int)dp.u.sl, sdl = sgn * (int)dp.delta.sl, sz = sgn * (int)p.z.sl;
  const int spre = sgn * (int)p.out_pre.sl, sdo = sgn * (int)p.dout.sl, sdz = sgn * (int)p.dz.sl;
  const int sdu = sgn * (int)bd.du.sl, sdd = sgn * (int)bd.ddelta.sl;
  const float* ckp = dp.ckpt + b * dp.ckpt_sb + d * dp.ckpt_sd;
  float* partp = BC_CONST ? nullptr
                          : bd.dBC_part + ((int64_t)b * gridDim.x + slab) * (int64_t)L * 32 + l0 * 32;
  const int spart = sgn * 32;

  // B/C staging source of this lane (see the forward kernel): 8 loads per lane per 8-step tile
  const T *bcpA = nullptr, *bcpB = nullptr;
  int64_t bc_hiA = 0, bc_hiB = 0;
  bool bc_okA = false, bc_tc = false;
  int sbc = 0;
  if (!BC_CONST) {
    const T* Bp = static_cast<const T*>(dp.Bm.ptr) + b * dp.Bm.sb + l0 * dp.Bm.sl;
    const T* Cp = static_cast<const T*>(dp.Cm.ptr) + b * dp.Cm.sb + l0 * dp.Cm.sl;
    bc_tc = (dp.Bm.sl == 1 && dp.Cm.sl == 1);
    if (bc_tc) {
      // lane -> step k = lane & 7, state q = lane >> 3 (0..3); loads j = 0..3: B[q + 4 j], then C[q + 4 j]
      const int q = lane >> 3, k = lane & 7;
      bcpA = Bp + (int64_t)q * dp.Bm.sd + (int64_t)k * sgn;
      bcpB = Cp + (int64_t)q * dp.Cm.sd + (int64_t)k * sgn;
      bc_hiA = 4 * dp.Bm.sd;
      bc_hiB = 4 * dp.Cm.sd;
      sbc = sgn;
    } else {
      const int n = lane & 15;
      bcpA = (lane < 16) ? (Bp + n * dp.Bm.sd) : (Cp + n * dp.Cm.sd);
      bc_okA = n < p.dstate;
      sbc = sgn * (int)((lane < 16) ? dp.Bm.sl : dp.Cm.sl);
    }
  }

  const int s1 = cm_first_range(L, p.ndir, dp.reverse);

  // ranges are walked last -> first: [s1, L) (absent for ndir == 1) then [0, s1)
#pragma unroll 1
  for (int range = (p.ndir == 2 ? 1 : 0); range >= 0; --range) {
    const int s_begin = (range == 1) ? s1 : 0;
    const int s_end = (range == 1) ? L : s1;
    const int j0 = (range == 1) ? cm_ceil_div(s1, kTile) : 0;
    const int ntile = cm_ceil_div(s_end - s_begin, kTile);
#pragma unroll 1
    for (int t = ntile - 1; t >= 0; --t) {
      const int s0 = s_begin + t * kTile;
      const int nvalid = min(kTile, s_end - s0);

      // ---- phase 0: every global load of the tile, back to back, into raw registers ------------------
      Raw ur[kTile], dr[kTile], gr[kTile], zr[kTile], pr[kTile], bcr[kTile];
#pragma unroll
      for (int k = 0; k < kTile; ++k) {
        ur[k] = Raw(0); dr[k] = Raw(0); gr[k] = Raw(0); zr[k] = Raw(0); pr[k] = Raw(0); bcr[k] = Raw(0);
        if (k < nvalid) {
          const int64_t s = s0 + k;
          ur[k] = Elem<T>::ld_raw(up + s * su);
          dr[k] = Elem<T>::ld_raw(dlp + s * sdl);
          gr[k] = Elem<T>::ld_raw(dop + s * sdo);
          if (has_z) zr[k] = Elem<T>::ld_raw(zp + s * sz);
          if (do_dz) pr[k] = Elem<T>::ld_raw(prep + s * spre);
        }
      }
      if (!BC_CONST) {
        if (bc_tc) {
          if ((lane & 7) < nvalid) {
            const int64_t off = (int64_t)s0 * sbc;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if ((lane >> 3) + 4 * j < p.dstate) {
                bcr[j] = Elem<T>::ld_raw(bcpA + j * bc_hiA + off);
                bcr[4 + j] = Elem<T>::ld_raw(bcpB + j * bc_hiB + off);
              }
            }
          }
        } else if (bc_okA) {
#pragma unroll
          for (int k = 0; k < kTile; ++k)
            if (k < nvalid) bcr[k] = Elem<T>::ld_raw(bcpA + (int64_t)(s0 + k) * sbc);
        }
      }
      float2 h2[NP];
      {
        const float4* src = reinterpret_cast<const float4*>(ckp + (int64_t)(j0 + t) * 16 + sg * NS);
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
          const float4 v = __ldg(src + i);
          h2[2 * i] = make_float2(v.x, v.y); h2[2 * i + 1] = make_float2(v.z, v.w);
        }
      }
      // publish B/C to the warp
      if (!BC_CONST) {
        __syncwarp();
        if (bc_tc) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            sm.bc[lane & 7][(lane >> 3) + 4 * j] = Elem<T>::cvt(bcr[j]);
            sm.bc[lane & 7][16 + (lane >> 3) + 4 * j] = Elem<T>::cvt(bcr[4 + j]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < kTile; ++k) sm.bc[k][lane] = Elem<T>::cvt(bcr[k]);
        }
        __syncwarp();
      }
      // per-step scalars
      float uu[kTile], xx[kTile], gg[kTile];
#pragma unroll
      for (int k = 0; k < kTile; ++k) {
        uu[k] = Elem<T>::cvt(ur[k]);
        xx[k] = Elem<T>::cvt(dr[k]) + bias;
        const float dov = Elem<T>::cvt(gr[k]) * scale;
        if (has_z) {
          const float zz = Elem<T>::cvt(zr[k]);
          const float sig = sigmoidf_fast(zz);
          gg[k] = dov * zz * sig;
          if (do_dz && sg == 0 && dvalid && k < nvalid)
            Elem<T>::st(dzp + (int64_t)(s0 + k) * sdz, dov * Elem<T>::cvt(pr[k]) * sig * fmaf(zz, 1.f - sig, 1.f));
        } else {
          gg[k] = dov;
        }
      }

      // ---- phase 1: recompute the tile's states from its checkpoint (fp32 pairs: FMUL2 / FFMA2) ------
      float dtv[kTile];
#pragma unroll
      for (int k = 0; k < kTile; ++k) {
        dtv[k] = 0.f;
        if (k < nvalid) {
          const float dt = softplus ? softplus_fwd<sizeof(T) == 4>(xx[k]) : xx[k];
          dtv[k] = dt;
          const float du_ = dt * uu[k];
          const float2 dt2 = make_float2(dt, dt), du2 = make_float2(du_, du_);
          float2 B2[NP];
          if (BC_CONST) {
#pragma unroll
            for (int i = 0; i < NP; ++i) B2[i] = make_float2(Bc[2 * i], Bc[2 * i + 1]);
          } else {
            load_row2<NP>(&sm.bc[k][sg * NS], B2);
          }
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

      // ---- phase 2: reverse sweep -----------------------------------------------------------------
#pragma unroll
      for (int k = kTile - 1; k >= 0; --k) {
        if (k < nvalid) {
          const int64_t s = s0 + k;
          const float dt = dtv[k], u_ = uu[k], g = gg[k];
          const float du_ = dt * u_;
          const float2 dt2 = make_float2(dt, dt), du2 = make_float2(du_, du_), g2 = make_float2(g, g);
          const float2 ndu2 = make_float2(-du_, -du_);
          float2 hk2[NP], B2[NP], C2[NP], dB2[NP], dC2[NP];
#pragma unroll
          for (int i = 0; i < NS / 4; ++i) {
            const float4 v = sm.h[k][i][lane];
            hk2[2 * i] = make_float2(v.x, v.y); hk2[2 * i + 1] = make_float2(v.z, v.w);
          }
          if (BC_CONST) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              B2[i] = make_float2(Bc[2 * i], Bc[2 * i + 1]);
              C2[i] = make_float2(Cc[2 * i], Cc[2 * i + 1]);
            }
          } else {
            load_row2<NP>(&sm.bc[k][sg * NS], B2);
            load_row2<NP>(&sm.bc[k][16 + sg * NS], C2);
          }
          float2 sLB2 = make_float2(0.f, 0.f), sWA2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const float2 lam = ffma2(g2, C2[i], mu2[i]);          // lambda = g*C + mu
            dC2[i] = fmul2(g2, hk2[i]);
            dB2[i] = fmul2(lam, du2);
            sLB2 = ffma2(lam, B2[i], sLB2);
            const float2 hp = ffma2(ndu2, B2[i], hk2[i]);         // = a * h_{s-1}
            const float2 w = fmul2(lam, hp);
            sWA2 = ffma2(w, kA2[i], sWA2);
            dA2[i] = ffma2(w, dt2, dA2[i]);
            const float2 x2 = fmul2(dt2, kA2[i]);
            mu2[i] = fmul2(make_float2(ex2(x2.x), ex2(x2.y)), lam);
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
          const float du = fmaf(g, Dsk, dt * sLB);
          const float ddt = fmaf(u_, sLB, kLn2 * sWA);
          const float ddl = softplus ? ddt * softplus_grad(xx[k]) : ddt;
          dD_acc = fmaf(g, u_, dD_acc);
          dbias_acc += ddl;
          if (sg == 0 && dvalid) {
            Elem<T>::st(dup + s * sdu, du);
            Elem<T>::st(ddp + s * sdd, ddl);
          }
          if (BC_CONST) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              dBacc[2 * i] += dB2[i].x; dBacc[2 * i + 1] += dB2[i].y;
              dCacc[2 * i] += dC2[i].x; dCacc[2 * i + 1] += dC2[i].y;
            }
          } else {
            // cross-channel reduce of the 32 per-step values
            __syncwarp();
            float4* dst = reinterpret_cast<float4*>(&sm.red[lane][0]);
            const float keep = dvalid ? 1.f : 0.f;
#pragma unroll
            for (int i = 0; i < NS / 4; ++i) {
              dst[i] = make_float4(keep * dB2[2 * i].x, keep * dB2[2 * i].y, keep * dB2[2 * i + 1].x, keep * dB2[2 * i + 1].y);
              dst[NS / 4 + i] =
                  make_float4(keep * dC2[2 * i].x, keep * dC2[2 * i].y, keep * dC2[2 * i + 1].x, keep * dC2[2 * i + 1].y);
            }
            __syncwarp();
            const int n = lane & 15;
            const int col = ((lane < 16) ? 0 : NS) + (n % NS);
            const int row0 = n / NS;
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int c = 0; c < CPW; c += 2) {
              acc0 += sm.red[c * LPC + row0][col];
              acc1 += sm.red[(c + 1) * LPC + row0][col];
            }
            partp[s * spart + lane] = acc0 + acc1;
          }
        }
      }
    }
  }

  if (dvalid) {
    const int64_t row = (int64_t)b * p.dim + d;
#pragma unroll
    for (int i = 0; i < NS; ++i) bd.dA_part[row * 16 + sg * NS + i] = (i & 1) ? dA2[i / 2].y : dA2[i / 2].x;
    if (sg == 0) {
      if (bd.dD_part) bd.dD_part[row] = dD_acc;
      if (bd.dbias_part) bd.dbias_part[row] = dbias_acc;
    }
    if (BC_CONST) {
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        bd.dBC_part[row * 32 + sg * NS + i] = dBacc[i];
        bd.dBC_part[row * 32 + 16 + sg * NS + i] = dCacc[i];
      }
    }
  }
}

template <typename T>
static int launch_bwd_t(const cm_scan_bwd_args& a, int lpc, bool bc_const, cudaStream_t st) {
#define CM_BWD_CASE(LPC_, BCC_)                                                  \
  {                                                                              \
    const dim3 grid(cm_ceil_div(a.dim, 32 / LPC_), a.batch, a.ndir);             \
    scan_bwd_kernel<T, LPC_, BCC_><<<grid, 32, 0, st>>>(a);                      \
  }
  if (!bc_const) {
    if (lpc == 1) CM_BWD_CASE(1, false) else if (lpc == 2) CM_BWD_CASE(2, false) else CM_BWD_CASE(4, false)
  } else {
    if (lpc == 1) CM_BWD_CASE(1, true) else if (lpc == 2) CM_BWD_CASE(2, true) else CM_BWD_CASE(4, true)
  }
#undef CM_BWD_CASE
  CM_LAUNCH_CHECK();
  return 0;
}

// ---- deterministic reducers -----------------------------------------------------------------------
template <typename T>
__global__ void reduce_dbc_kernel(const float* __restrict__ part, int n_slab, int L, int dstate, cm_tensor3 dB,
                                  cm_tensor3 dC) {
  // one thread per (b, l, v): sums the slabs in order
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = idx & 31;
  const int64_t bl = idx >> 5;
  const int b = blockIdx.y;
  if (bl >= L) return;
  const float* src = part + ((int64_t)b * n_slab * L + bl) * 32 + v;
  float acc = 0.f;
  for (int s = 0; s < n_slab; ++s) acc += __ldg(src + (int64_t)s * L * 32);
  const int n = v & 15;
  if (n < dstate) {
    const cm_tensor3& o = (v < 16) ? dB : dC;
    Elem<T>::st(static_cast<T*>(o.ptr) + b * o.sb + n * o.sd + bl * o.sl, acc);
  }
}

__global__ void reduce_rows_kernel(const float* __restrict__ part, int64_t rows, int64_t cols,
                                   float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int64_t r = 0; r < rows; ++r) acc += __ldg(part + r * cols + c);
  out[c] = acc;
}

// Up to CM_REDUCE_MAX_JOBS column-sum jobs in one launch.  blockIdx.y = job; a CTA of 32 x 8 threads owns 32 columns,
// its 8 warps stride over the rows and combine through shared memory in a fixed order (deterministic).
struct ReduceJobs {
  cm_reduce_job j[CM_REDUCE_MAX_JOBS];
};
constexpr int kRmWarps = 32;
constexpr int kRmFly = 8;     // independent loads in flight per thread
__global__ void __launch_bounds__(32 * kRmWarps) reduce_multi_kernel(const ReduceJobs jobs) {
  __shared__ float sm[kRmWarps][33];
  const cm_reduce_job& job = jobs.j[blockIdx.y];
  const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if ((int64_t)blockIdx.x * 32 >= job.cols) return;      // CTA-uniform
  // kRmFly independent running sums per thread (rows ty, ty + 32, ...: round-robin), combined in a fixed order: the
  // loads of a thread are in flight together instead of one L2 round trip per row.  The kernel is pure latency (a few
  // hundred partial rows of a few hundred columns): 592 rows = 3 round trips per thread.
  float acc[kRmFly];
#pragma unroll
  for (int i = 0; i < kRmFly; ++i) acc[i] = 0.f;
  if (c < job.cols) {
    const float* src = job.part + c;
    const int64_t rows = job.rows, cols = job.cols;
    int64_t r = threadIdx.y;
    for (; r + (kRmFly - 1) * kRmWarps < rows; r += kRmFly * kRmWarps) {
      float v[kRmFly];
#pragma unroll
      for (int i = 0; i < kRmFly; ++i) v[i] = __ldg(src + (r + i * kRmWarps) * cols);
#pragma unroll
      for (int i = 0; i < kRmFly; ++i) acc[i] += v[i];
    }
    float v[kRmFly];
#pragma unroll
    for (int i = 0; i < kRmFly; ++i) v[i] = (r + i * kRmWarps < rows) ? __ldg(src + (r + i * kRmWarps) * cols) : 0.f;
#pragma unroll
    for (int i = 0; i < kRmFly; ++i) acc[i] += v[i];
  }
  sm[threadIdx.y][threadIdx.x] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (threadIdx.y == 0 && c < job.cols) {
    float t = sm[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < kRmWarps; ++y) t += sm[y][threadIdx.x];
    job.out[c] = t;
  }
}

}  // namespace cm

extern "C" int cm_reduce_multi(const cm_reduce_job* jobs, int32_t njobs, void* stream) {
  if (jobs == nullptr || njobs <= 0 || njobs > CM_REDUCE_MAX_JOBS) return CM_ERR_BAD_ARG;
  cm::ReduceJobs rj;
  int64_t maxcols = 0;
  for (int i = 0; i < njobs; ++i) {
    if (!jobs[i].part || !jobs[i].out || jobs[i].rows <= 0 || jobs[i].cols <= 0) return CM_ERR_BAD_ARG;
    rj.j[i] = jobs[i];
    if (jobs[i].cols > maxcols) maxcols = jobs[i].cols;
  }
  for (int i = njobs; i < CM_REDUCE_MAX_JOBS; ++i) rj.j[i] = jobs[0];
  const dim3 grid((unsigned)((maxcols + 31) / 32), njobs);
  cm::reduce_multi_kernel<<<grid, dim3(32, cm::kRmWarps), 0, static_cast<cudaStream_t>(stream)>>>(rj);
  CM_LAUNCH_CHECK();
  return 0;
}

namespace cm {
int scan_bwd_try_channel_last(const cm_scan_bwd_args& a, int lpc, cudaStream_t st, int* rc);   // scan_bwd_cl.cu
int scan_bwd_try_state_parallel(const cm_scan_bwd_args& a, cudaStream_t st, int* rc);            // scan_bwd_sp.cu
int scan_bwd_try_lane_channel(const cm_scan_bwd_args& a, cudaStream_t st, int* rc);              // scan_bwd_lc.cu
int scan_bwd_lane_channel_applies(const cm_scan_bwd_args& a);
int scan_bwd_lane_channel_slab();
int scan_bwd_try_warpgroup(const cm_scan_bwd_args& a, cudaStream_t st, int* rc);                 // scan_bwd_wg.cu
int scan_bwd_warpgroup_applies(const cm_scan_bwd_args& a);
int scan_bwd_warpgroup_slab();
}

extern "C" int cm_scan_bwd_slab_channels(const cm_scan_bwd_args* args) {
  if (args == nullptr || args->dim <= 0) return CM_ERR_BAD_ARG;
  cm_scan_bwd_args a = *args;
  // the partial-sum pointers are what the caller is about to allocate: they must not influence the decision
  static float dummy_part[4] __attribute__((aligned(16)));
  for (int r = 0; r < 2; ++r) { a.dir[r].dBC_part = dummy_part; a.dir[r].dA_part = dummy_part; }
  if (cm::scan_bwd_lane_channel_applies(a)) return cm::scan_bwd_lane_channel_slab();
  if (cm::scan_bwd_warpgroup_applies(a)) return cm::scan_bwd_warpgroup_slab();
  int lpc = a.lanes_per_channel;
  if (lpc == 0) lpc = cm_scan_pick_lanes_bwd(a.batch, a.dim, a.ndir);
  return cm_scan_slab_channels(lpc);
}

extern "C" int cm_scan_pick_lanes_bwd(int32_t batch, int32_t dim, int32_t ndir) {
  // The default backward kernel (scan_bwd_sp.cu) reduces dB/dC over 32-channel slabs = cm_scan_slab_channels(1).
  // Callers size the partial tensor from this value; when the state-parallel kernel does not apply (strided layout,
  // dim % 32 != 0, constant B/C) the lane-per-channel kernels run with the same slab width.  (Measured on B200: for
  // those kernels 2 lanes per channel are 10-15 % faster; pass lanes_per_channel = 2 explicitly to get that.)
  (void)batch; (void)dim; (void)ndir;
  return 1;
}

extern "C" int cm_scan_bwd(const cm_scan_bwd_args* args, void* stream) {
  if (args == nullptr) return CM_ERR_BAD_ARG;
  const cm_scan_bwd_args& a = *args;
  if (a.batch <= 0 || a.dim <= 0 || a.seqlen <= 0 || a.dstate <= 0) return CM_ERR_BAD_ARG;
  if (a.ndir != 1 && a.ndir != 2) return CM_ERR_BAD_ARG;
  if (!cm::dtype_ok(a.dtype) || a.dout.ptr == nullptr) return CM_ERR_BAD_ARG;
  if (a.dstate > CM_MAX_DSTATE || a.batch > 65535) return CM_ERR_UNSUPPORTED;
  if (a.z.ptr != nullptr && (a.out_pre.ptr == nullptr || a.dz.ptr == nullptr)) return CM_ERR_BAD_ARG;
  for (int r = 0; r < a.ndir; ++r) {
    const cm_scan_bwd_dir& d = a.dir[r];
    if (!d.in.u.ptr || !d.in.delta.ptr || !d.in.Bm.ptr || !d.in.Cm.ptr || !d.in.A || !d.in.ckpt) return CM_ERR_BAD_ARG;
    if (!d.du.ptr || !d.ddelta.ptr || !d.dBC_part || !d.dA_part) return CM_ERR_BAD_ARG;
  }
  if (a.ndir == 2) {
    if ((a.dir[0].in.reverse != 0) == (a.dir[1].in.reverse != 0)) return CM_ERR_BAD_ARG;
    if (a.dir[0].in.bc_const != a.dir[1].in.bc_const) return CM_ERR_UNSUPPORTED;
  }
  int lpc = a.lanes_per_channel;
  if (lpc == 0) lpc = cm_scan_pick_lanes_bwd(a.batch, a.dim, a.ndir);
  if (lpc != 1 && lpc != 2 && lpc != 4) return CM_ERR_BAD_ARG;
  const bool bcc = a.dir[0].in.bc_const != 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int slab = cm_scan_bwd_slab_channels(args);
  if (slab == cm::scan_bwd_lane_channel_slab()) {
    // the caller sized dBC_part for 128-channel slabs: no other kernel may run (they write narrower slabs)
    int rc = 0;
    if (cm::scan_bwd_try_lane_channel(a, st, &rc)) return rc;
    return CM_ERR_UNSUPPORTED;                   // misaligned partial buffers, or a view the tensor-map encoder refused
  }
  if (slab == cm::scan_bwd_warpgroup_slab()) {   // 64-channel slabs: the warpgroup kernel (scan_bwd_wg.cu), same contract
    int rc = 0;
    if (cm::scan_bwd_try_warpgroup(a, st, &rc)) return rc;
    return CM_ERR_UNSUPPORTED;
  }
  if (getenv("CM_SCAN_GENERIC") == nullptr && getenv("CM_SCAN_NO_SP") == nullptr && lpc == 1) {
    int rc = 0;
    if (cm::scan_bwd_try_state_parallel(a, st, &rc)) return rc;
  }
  if (getenv("CM_SCAN_GENERIC") == nullptr) {   // env switch only for A/B measurements of the two kernels
    int rc = 0;
    if (cm::scan_bwd_try_channel_last(a, lpc, st, &rc)) return rc;
  }
  switch (a.dtype) {
    case CM_F32: return cm::launch_bwd_t<float>(a, lpc, bcc, st);
    case CM_BF16: return cm::launch_bwd_t<__nv_bfloat16>(a, lpc, bcc, st);
    default: return cm::launch_bwd_t<__half>(a, lpc, bcc, st);
  }
}

extern "C" int cm_reduce_dbc(const float* part, int32_t batch, int32_t n_slab, int32_t seqlen, int32_t dstate,
                             int32_t dtype, cm_tensor3 dB, cm_tensor3 dC, void* stream) {
  if (!part || !dB.ptr || !dC.ptr || batch <= 0 || n_slab <= 0 || seqlen <= 0 || dstate <= 0) return CM_ERR_BAD_ARG;
  if (!cm::dtype_ok(dtype)) return CM_ERR_BAD_ARG;
  if (dstate > CM_MAX_DSTATE || batch > 65535) return CM_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  const dim3 grid(cm_ceil_div(seqlen * 32, threads), batch);
  switch (dtype) {
    case CM_F32: cm::reduce_dbc_kernel<float><<<grid, threads, 0, st>>>(part, n_slab, seqlen, dstate, dB, dC); break;
    case CM_BF16: cm::reduce_dbc_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(part, n_slab, seqlen, dstate, dB, dC); break;
    default: cm::reduce_dbc_kernel<__half><<<grid, threads, 0, st>>>(part, n_slab, seqlen, dstate, dB, dC); break;
  }
  CM_LAUNCH_CHECK();
  return 0;
}

extern "C" int cm_reduce_rows(const float* part, int64_t rows, int64_t cols, float* out, void* stream) {
  if (!part || !out || rows <= 0 || cols <= 0) return CM_ERR_BAD_ARG;
  const int threads = 128;
  cm::reduce_rows_kernel<<<(unsigned)((cols + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      part, rows, cols, out);
  CM_LAUNCH_CHECK();
  return 0;
}
