// probe: are __vaddss2/__vsubss2/__vneg2/__vmaxs2 exact on sm_100a for edge values?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(const int16_t* v, int n, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n) return;
  int a = v[i / n], b = v[i % n];
  unsigned pa = ((unsigned)a & 0xffff) | ((unsigned)b << 16), pb = ((unsigned)b & 0xffff) | ((unsigned)a << 16);
  unsigned add = __vaddss2(pa, pb), sub = __vsubss2(pa, pb);
  auto sat = [](int x) { return x > 32767 ? 32767 : (x < -32768 ? -32768 : x); };
  int e_add_lo = sat(a + b), e_add_hi = sat(b + a), e_sub_lo = sat(a - b), e_sub_hi = sat(b - a);
  if ((int16_t)(add & 0xffff) != e_add_lo || (int16_t)(add >> 16) != e_add_hi) atomicAdd(&bad[0], 1);
  if ((int16_t)(sub & 0xffff) != e_sub_lo || (int16_t)(sub >> 16) != e_sub_hi) { if (atomicAdd(&bad[1], 1) < 5) printf("sub a=%d b=%d got lo=%d hi=%d want %d %d\n", a, b, (int16_t)(sub & 0xffff), (int16_t)(sub >> 16), e_sub_lo, e_sub_hi); }
}
int main() {
  const int n = 2048;
  int16_t h[n];
  for (int i = 0; i < n; ++i) h[i] = (int16_t)(i < 1024 ? -32768 + i * 37 % 2000 : 32767 - (i * 53) % 3000);
  h[0] = -32768; h[1] = 32767; h[2] = 0; h[3] = -1; h[4] = 1; h[5] = -32767; h[6] = 16384; h[7] = -16384;
  for (int i = 8; i < 600; ++i) h[i] = (int16_t)(rand() & 0xffff);
  int16_t* d; int* bad; int hb[2];
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  k<<<(n * n + 255) / 256, 256>>>(d, n, bad);
  cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost);
  printf("vaddss2 mismatches %d, vsubss2 mismatches %d (of %d pairs)\n", hb[0], hb[1], n * n);
  return 0;
}
