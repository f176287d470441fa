// DRAM efficiency vs. contiguity of per-block accesses: 4-thread groups each stream through their own
// region (like one code block's array in k_map16), CH bytes contiguous per group and iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/gran_probe tools/gran_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NLD>
__global__ void __launch_bounds__(256) k(const uint4* base, long region_u4, int nregions, uint4* out) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long grp = gt >> 2; const int t = gt & 3;
  if (grp >= nregions) return;
  const uint4* p = base + grp * region_u4 + t;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (long i = 0; i < region_u4 / (4 * NLD); ++i) {
    uint4 v[NLD];
#pragma unroll
    for (int j = 0; j < NLD; ++j) v[j] = __ldg(p + (i * NLD + j) * 4);
#pragma unroll
    for (int j = 0; j < NLD; ++j) { acc.x ^= v[j].x; acc.y ^= v[j].y; acc.z ^= v[j].z; acc.w ^= v[j].w; }
  }
  if (acc.x == 0x12345678u) out[gt] = acc;
}
template <int NLD>
void run(const uint4* buf, uint4* out, long region_bytes, int nregions, int cta, size_t smem) {
  const long region_u4 = region_bytes / 16;
  const long threads = (long)nregions * 4;
  const int grid = (int)((threads + cta - 1) / cta);
  cudaFuncSetAttribute(k<NLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<NLD><<<grid, cta, smem>>>(buf, region_u4, nregions, out);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) k<NLD><<<grid, cta, smem>>>(buf, region_u4, nregions, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("contiguous %4d B/group/iter  cta %3d smem %6zu: %8.1f GB/s\n", NLD * 64, cta, smem, 3.0 * nregions * region_bytes / (ms * 1e-3) / 1e9);
}
int main() {
  const long region_bytes = 12288; const int nregions = 75776;
  uint4* buf; uint4* out;
  cudaMalloc(&buf, (size_t)region_bytes * nregions); cudaMemset(buf, 1, (size_t)region_bytes * nregions);
  cudaMalloc(&out, (size_t)nregions * 4 * 16);
  for (int pass = 0; pass < 2; ++pass) {
    const int cta = pass ? 64 : 256; const size_t smem = pass ? 45056 : 0;     // pass 1: k_map16-like occupancy (10 warps/SM)
    run<1>(buf, out, region_bytes, nregions, cta, smem);
    run<2>(buf, out, region_bytes, nregions, cta, smem);
    run<4>(buf, out, region_bytes, nregions, cta, smem);
    run<8>(buf, out, region_bytes, nregions, cta, smem);
  }
  return 0;
}
