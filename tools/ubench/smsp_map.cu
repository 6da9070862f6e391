// Microbenchmark: how are the warps of small CTAs mapped onto the 4 SM sub-partitions (SMSPs)?
// A MUFU-bound loop (XU pipe: 4 lanes / clk / SMSP) is run with the same total number of warps per SM arranged as
// CTAs of 1, 2, 4 and 8 warps.  If a CTA's warps were mapped by (warp id within the CTA) % 4, 1- and 2-warp CTAs would
// leave 3 resp. 2 of the 4 XU pipes idle and run 4x / 2x slower.  Also prints %warpid of the first warps of SM 0.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, unsigned* wid) {
  float x = threadIdx.x * 1e-3f, acc = 0.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x + j));
      acc += y;
    }
  }
  unsigned sm, w;
  asm("mov.u32 %0, %%smid;" : "=r"(sm));
  asm("mov.u32 %0, %%warpid;" : "=r"(w));
  if (sm == 0 && (threadIdx.x & 31) == 0) wid[blockIdx.x * 8 + (threadIdx.x >> 5)] = w + 1;
  if (acc == 123.f) out[0] = acc;
}
int main() {
  float* out; unsigned* wid;
  cudaMalloc(&out, 4); cudaMalloc(&wid, 1 << 20);
  const int sms = 148, warps_per_sm = 16, iters = 20000;
  for (int wpc = 1; wpc <= 8; wpc *= 2) {
    cudaMemset(wid, 0, 1 << 20);
    const int grid = sms * warps_per_sm / wpc;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<grid, 32 * wpc>>>(out, 100, wid);
    cudaEventRecord(a);
    k<<<grid, 32 * wpc>>>(out, iters, wid);
    cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double mufu = (double)grid * wpc * iters * 8;
    printf("warps/CTA %d grid %d: %.3f ms  -> %.2f cycles per warp-MUFU per SMSP (8 = all four XU pipes busy)\n", wpc, grid, ms,
           ms * 1e-3 * 1.965e9 / (mufu / (sms * 4)));
    static unsigned h[1 << 18];
    cudaMemcpy(h, wid, 1 << 20, cudaMemcpyDeviceToHost);
    printf("  %%warpid of SM 0's warps (CTA order): ");
    int n = 0;
    for (int i = 0; i < (1 << 18) && n < 24; ++i) if (h[i]) { printf("%u ", h[i] - 1); ++n; }
    printf("\n");
  }
  return 0;
}
