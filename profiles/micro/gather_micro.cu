// gather_micro.cu -- what does it cost to look again at ~1 M random places of a 268 MB buffer that a streaming kernel has
// just read once (the flagged grams of lookup_kernel)?  Variables: L2 fetch granularity (cudaLimitMaxL2FetchGranularity),
// order of the positions (random / ascending), bytes per place (4 / 32), where they lie (whole buffer / last 64 MB).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_micro gather_micro.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void stream_pass(const uint4* __restrict__ p, size_t n16, uint32_t* out) {
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
template <int WORDS>
__global__ void gather(const uint8_t* __restrict__ buf, const uint32_t* __restrict__ pos, uint32_t n, uint32_t* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(buf + (pos[i] & ~3u));
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < WORDS; k++) { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p + k)); acc ^= v; }
    out[i] = acc;
  }
}
int main(int argc, char** argv) {
  const size_t bytes = 268435456ull + 64;
  for (int gran : {0, 32, 64, 128}) {
    if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("# set granularity %d -> %s, now %zu\n", gran, cudaGetErrorString(e), g); }
    uint8_t* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    uint32_t *dpos, *dout; const uint32_t NMAX = 4u << 20; cudaMalloc(&dpos, NMAX * 4); cudaMalloc(&dout, NMAX * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint32_t n : {850000u, 3400000u}) for (int sorted = 0; sorted < 2; sorted++) for (int tailonly = 0; tailonly < 2; tailonly++) {
      std::vector<uint32_t> pos(n); uint64_t x = 88172645463325252ull;
      for (auto& p : pos) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; p = (uint32_t)(tailonly ? (268435456ull - 67108864ull) + x % 67108864ull : x % 268435456ull) & ~1u; }
      if (sorted) std::sort(pos.begin(), pos.end());
      cudaMemcpy(dpos, pos.data(), n * 4, cudaMemcpyHostToDevice);
      for (int words : {1, 2, 8}) for (int grid : {148 * 6, 148 * 24}) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
          stream_pass<<<148 * 8, 256>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, dout);
          cudaEventRecord(e0);
          if (words == 1) gather<1><<<grid, 256>>>(buf, dpos, n, dout); else if (words == 2) gather<2><<<grid, 256>>>(buf, dpos, n, dout); else gather<8><<<grid, 256>>>(buf, dpos, n, dout);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
        }
        printf("gran %3d  n %7u  %s  %s  words %d  grid %4d : %7.1f us  (%.2f G places/s)\n", gran, n, sorted ? "ascending" : "random   ", tailonly ? "last-64MB" : "whole-buf", words, grid, best * 1e3, n / best / 1e6);
      }
    }
    { float ms; cudaEventRecord(e0); stream_pass<<<148 * 8, 256>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, dout); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); printf("# stream pass %.1f us\n", ms * 1e3); }
    cudaFree(buf); cudaFree(dpos); cudaFree(dout);
  }
  return 0;
}
