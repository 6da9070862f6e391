// Microbenchmark (round 2): legacy mma.sync m16n8k8 TF32 on sm_100a as a cross-lane reducer, next to SHFL and MUFU.
//   * throughput of HMMA.1688.F32.TF32 per SM sub-partition (independent accumulators), alone and mixed with MUFU.EX2
//   * correctness of the split-precision column-selector reduction used by the scan kernels (sum over the 4 lanes of a quad)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mma_red tools/ubench/mma_red.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
    : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void bench(float* out, long long* cyc, int iters) {
  float acc[4][4] = {};
  float e[8];
  for (int i = 0; i < 8; ++i) e[i] = -0.001f * (threadIdx.x + i);
  unsigned a = __float_as_uint(1.0f + threadIdx.x), b = 0x3f800000u;
  float sh = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_tf32(acc[j], a, a, a, a, b, b);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) e[j] = ex2(e[j]);
    }
    if (MODE == 3) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sh += __shfl_xor_sync(0xffffffffu, sh, 4 << (j & 1));
    }
  }
  long long t1 = clock64();
  float s = sh;
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += acc[j][i];
  for (int i = 0; i < 8; ++i) s += e[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// correctness: every lane has values x (row g) and y (row g+8); selector column n; expect sums over the 4 lanes of the quad
__global__ void check(const float* x, const float* y, float* outx, float* outy, int n_sel) {
  const int lane = threadIdx.x, g = lane >> 2;
  float acc[4] = {0, 0, 0, 0};
  const float xv = x[lane], yv = y[lane];
  const unsigned hx = __float_as_uint(xv) & 0xffffe000u, hy = __float_as_uint(yv) & 0xffffe000u;
  const float lx = xv - __uint_as_float(hx), ly = yv - __uint_as_float(hy);
  for (int n = 0; n < n_sel; ++n) {
    const unsigned sel = (g == n) ? 0x3f800000u : 0u;
    mma_tf32(acc, hx, hy, __float_as_uint(lx), __float_as_uint(ly), sel, sel);
  }
  // lane (g, t) holds D[g][2t], D[g][2t+1], D[g+8][2t], D[g+8][2t+1]
  outx[lane * 2] = acc[0]; outx[lane * 2 + 1] = acc[1];
  outy[lane * 2] = acc[2]; outy[lane * 2 + 1] = acc[3];
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 16 * 1024 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
  const char* names[4] = {"4 HMMA.1688.TF32 per iter", "8 MUFU.EX2 per iter", "4 HMMA + 8 MUFU per iter", "8 SHFL (dependent) per iter"};
  for (int mode = 0; mode < 4; ++mode)
    for (int wps = 1; wps <= 8; wps *= 2) {   // warps per SM sub-partition
      const int threads = 128 * wps, iters = 4096;
      long long h[148];
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) bench<0><<<148, threads>>>(out, cyc, iters);
        if (mode == 1) bench<1><<<148, threads>>>(out, cyc, iters);
        if (mode == 2) bench<2><<<148, threads>>>(out, cyc, iters);
        if (mode == 3) bench<3><<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("%-32s warps/SMSP %d: %.1f cycles per iteration per SMSP (%.2f per warp-iteration)\n", names[mode], wps, avg / iters, avg / iters / wps);
    }
  // correctness
  float hx[32], hy[32], ox[64], oy[64];
  for (int i = 0; i < 32; ++i) { hx[i] = 1.0f + 0.123456789f * i + 1e-5f * i * i; hy[i] = -3.14159265f * (i + 1) * 1e-3f; }
  float *dx, *dy, *dox, *doy;
  cudaMalloc(&dx, 128); cudaMalloc(&dy, 128); cudaMalloc(&dox, 256); cudaMalloc(&doy, 256);
  cudaMemcpy(dx, hx, 128, cudaMemcpyHostToDevice); cudaMemcpy(dy, hy, 128, cudaMemcpyHostToDevice);
  check<<<1, 32>>>(dx, dy, dox, doy, 8);
  cudaMemcpy(ox, dox, 256, cudaMemcpyDeviceToHost); cudaMemcpy(oy, doy, 256, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int g = 0; g < 8; ++g) {
    double sx = 0, sy = 0;
    for (int t = 0; t < 4; ++t) { sx += hx[g * 4 + t]; sy += hy[g * 4 + t]; }
    // MMA number n adds the quad sums of every row into column n only: after n_sel = 8 calls every column holds the sum once
    for (int t = 0; t < 4; ++t)
      for (int s = 0; s < 2; ++s) {
        const int col = 2 * t + s;
        const double ex = sx, ey = sy; (void)col;
        const double rx = fabs(ox[(g * 4 + t) * 2 + s] - ex) / (fabs(sx) + 1e-30), ry = fabs(oy[(g * 4 + t) * 2 + s] - ey) / (fabs(sy) + 1e-30);
        if (rx > worst) worst = rx;
        if (ry > worst) worst = ry;
      }
  }
  printf("selector reduction: worst relative error %.3e (fp32 eps 6e-8, tf32 eps 4.9e-4)\n", worst);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
