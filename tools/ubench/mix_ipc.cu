// Microbenchmark: what instruction rate can one SM sub-partition sustain on the scan kernels' instruction mix?
// Each thread owns 16 independent fp32 chains (8 packed pairs).  One iteration = one "state step" of a lane-per-channel
// scan: 16 MUFU.EX2, 8 FMUL2 (dt*A), 8 FMUL2 (du*B), 8 FFMA2 (h), 8 FFMA2/FMUL2 (C*h), 8 LDS.128 (broadcast B|C), and NSC
// scalar FP32 instructions.  No dependency shorter than one full iteration except the h chain (1 FFMA2 per pair).
// Prints cycles per iteration and per-pipe rates for 1..8 warps per sub-partition; variants drop one instruction class at
// a time to expose which pipe binds.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm volatile("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float ex2v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MUFU, int PACKED, int LDS, int NSC>
__global__ void k(float* out, int iters, long long* cyc) {
  __shared__ float4 bc[64][8];
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) bc[i / 8][i % 8] = make_float4(1e-3f * i, 0.5f, 0.25f, 0.125f);
  __syncthreads();
  float2 h[8], kA[8];
  for (int j = 0; j < 8; ++j) { h[j] = make_float2(0.f, 0.f); kA[j] = make_float2(-1.f - j, -1.5f - j); }
  float dt = 1e-3f * (threadIdx.x & 31), du = 0.5f, acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int row = it & 63;
    float4 b4[4], c4[4];
    if (LDS) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { b4[q] = bc[row][q]; c4[q] = bc[row][4 + q]; }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) { b4[q] = make_float4(dt, du, dt, du); c4[q] = b4[q]; }
    }
    const float2 dt2 = make_float2(dt, dt), du2 = make_float2(du, du);
    float2 a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a2[j] = PACKED ? fmul2(dt2, kA[j]) : make_float2(dt, du);
    if (MUFU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a2[j] = make_float2(ex2v(a2[j].x), ex2v(a2[j].y));
    }
    float2 ya = make_float2(0.f, 0.f), yb = ya;
    if (PACKED) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h[2 * q] = ffma2(a2[2 * q], h[2 * q], fmul2(du2, make_float2(b4[q].x, b4[q].y)));
        h[2 * q + 1] = ffma2(a2[2 * q + 1], h[2 * q + 1], fmul2(du2, make_float2(b4[q].z, b4[q].w)));
        ya = ffma2(make_float2(c4[q].x, c4[q].y), h[2 * q], ya);
        yb = ffma2(make_float2(c4[q].z, c4[q].w), h[2 * q + 1], yb);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { h[j].x += a2[j].x * 1e-9f; }
    }
    float s = ya.x + ya.y + yb.x + yb.y;
#pragma unroll
    for (int n = 0; n < NSC; ++n) s = fmaf(s, 0.999f, 1e-3f * n);      // scalar work (dependent chain, off the critical path)
    acc += s;
    dt = dt * 0.9999f + 1e-7f;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  float r = acc;
  for (int j = 0; j < 8; ++j) r += h[j].x + h[j].y;
  if (r == 123.456f) out[0] = r;
}
template <int MUFU, int PACKED, int LDS, int NSC>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 4000;
  for (int wps = 1; wps <= 8; wps *= 2) {          // warps per sub-partition
    const int warps_sm = wps * 4;
    k<MUFU, PACKED, LDS, NSC><<<148, 32 * warps_sm>>>(out, 100, cyc);
    k<MUFU, PACKED, LDS, NSC><<<148, 32 * warps_sm>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_iter_smsp = (double)c / iters;               // cycles for `wps` warp-iterations on one SMSP
    const int instr = (MUFU ? 16 : 0) + (PACKED ? 16 + 16 + 8 : 8) + (LDS ? 8 : 0) + NSC + 8;
    printf("%-28s warps/SMSP %d: %7.1f cyc per round = %6.1f cyc per warp-iteration ; ~%d instr/iter -> IPC %.2f\n", name, wps,
           per_iter_smsp, per_iter_smsp / wps, instr, instr * wps / per_iter_smsp);
  }
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  run<1, 1, 1, 24>("full mix (16 MUFU)", out, cyc);
  run<0, 1, 1, 24>("no MUFU", out, cyc);
  run<1, 0, 1, 24>("no packed FP32x2", out, cyc);
  run<1, 1, 0, 24>("no LDS", out, cyc);
  run<1, 1, 1, 0>("no scalar work", out, cyc);
  run<1, 0, 0, 0>("MUFU only", out, cyc);
  run<0, 1, 0, 0>("packed only", out, cyc);
  return 0;
}
