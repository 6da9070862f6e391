// Microbenchmark (round 2): issue cost of packed FP32x2 (FFMA2) against scalar FFMA on sm_100a, alone and with an ALU
// instruction stream next to it, per SM sub-partition.  Answers: does an FFMA2 take one or two issue slots?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ffma2_rate tools/ubench/ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
template <int MODE>
__global__ void bench(float* out, long long* cyc, int iters) {
  float2 acc[8]; float s[16]; int iv[8];
  for (int i = 0; i < 8; ++i) { acc[i] = make_float2(threadIdx.x * 0.001f + i, 1.f); iv[i] = threadIdx.x + i; }
  for (int i = 0; i < 16; ++i) s[i] = threadIdx.x * 0.002f + i;
  const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, 2e-3f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = ffma2(acc[j], m, c);            // 8 FFMA2
    }
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[j]) : "f"(m.x), "f"(c.x));   // 16 FFMA
    }
    if (MODE == 2 || MODE == 3) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(iv[j]) : "r"(it), "r"(j));      // 8 LOP3 (ALU pipe)
    }
  }
  long long t1 = clock64();
  float r = 0;
  for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y + iv[i];
  for (int i = 0; i < 16; ++i) r += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const char* names[4] = {"8 FFMA2", "16 FFMA", "8 FFMA2 + 8 LOP3", "8 LOP3"};
  for (int mode = 0; mode < 4; ++mode)
    for (int wps = 1; wps <= 8; wps *= 2) {
      const int threads = 128 * wps, iters = 8192;
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
      printf("%-20s warps/SMSP %d: %.2f cycles per iteration per warp (%.2f per SMSP-iteration)\n", names[mode], wps, avg / iters / wps, avg / iters);
    }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
