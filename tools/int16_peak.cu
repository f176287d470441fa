// Microbenchmark of the packed-int16 instruction mix used by k_map16 (SURVEY.md 8d asks for a
// MEASURED integer-SIMD peak rather than a paper one).  Each test runs N independent dependency
// chains per thread so that the issue rate, not latency, is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/int16_peak tools/int16_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
template <int MODE>
__global__ void __launch_bounds__(256) k(u32* out, u32 seed, int iters) {
  u32 a[8], x = seed + threadIdx.x, y = seed * 3 + 1, z = seed ^ 0x00070003u;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = x + i * 0x00010001u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = __vadd2(a[i], y);                                   // VIADD.16x2
        if (MODE == 1) a[i] = __viaddmax_s16x2(a[i], y, z);                       // VIADDMNMX.S16x2
        if (MODE == 2) a[i] = __vmaxs2(a[i], y ^ a[(i + 1) & 7]);                 // VIMNMX.S16x2 (+LOP3)
        if (MODE == 3) a[i] = a[i] * 0xffffu + y;                                 // IMAD
        if (MODE == 4) { a[i] = __viaddmax_s16x2(a[i], y, z); a[i] = __vadd2(a[i], z); }              // 1:1 mix
        if (MODE == 5) { a[i] = __viaddmax_s16x2(a[i], y, z); a[i] = a[i] * 3u + z; }                 // DPX + IMAD (FMA pipe)
        if (MODE == 6) { a[i] = __viaddmax_s16x2(a[i], y, z); a[i] = __viaddmax_s16x2(a[i], z, y); a[i] = __vadd2(a[i], y); a[i] = a[i] * 3u + z; }  // k_map16-like 2:1:1
        if (MODE == 7) a[i] = __vimax3_s16x2(a[i], y, z ^ a[(i + 1) & 7]);        // VIMNMX3.S16x2 (+LOP3)
      }
    }
  }
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, int ops_per_inner) {
  int dev, sms, khz;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int grid = sms * 8, iters = 2000;
  u32* out;
  cudaMalloc(&out, grid * 256 * 4);
  k<MODE><<<grid, 256>>>(out, 12345, 10);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, 12345, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)grid * 8 /*warps*/ * iters * 64 * ops_per_inner;
  const double per_sm_per_clk = warp_instr / sms / (ms * 1e-3 * khz * 1e3);
  printf("%-44s %8.3f ms  %6.3f warp-instr/clk/SM (at max clock %d MHz)  = %7.1f G packed ops/s = %7.1f G int16 ops/s\n", name, ms,
         per_sm_per_clk, khz / 1000, warp_instr * 32 / (ms * 1e-3) / 1e9, warp_instr * 64 / (ms * 1e-3) / 1e9);
  cudaFree(out);
}
int main() {
  run<0>("VIADD.16x2", 1);
  run<1>("VIADDMNMX.S16x2", 1);
  run<2>("VIMNMX.S16x2 + LOP3", 2);
  run<3>("IMAD", 1);
  run<4>("VIADDMNMX + VIADD", 2);
  run<5>("VIADDMNMX + IMAD", 2);
  run<6>("2 VIADDMNMX + VIADD + IMAD", 4);
  run<7>("VIMNMX3.S16x2 + LOP3", 2);
  return 0;
}
