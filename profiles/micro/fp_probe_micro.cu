// Microbenchmark (round 1): is a lane-private fingerprint-table probe per byte fast enough?
// Each lane streams its own 256-byte message (16-byte loads), and for EVERY byte position hashes the
// last 4 folded bytes, reads one 32-bit bucket from ITS OWN shared-memory bank (conflict-free by
// construction) and compares two 16-bit fingerprints.  Reports achieved GB/s of message bytes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int T = 1024;
constexpr int BUCKETS = 1024;                 // 1024 x 4 B x 32 lanes-banks = 128 KB
__global__ void __launch_bounds__(T, 1) probe(const uint4* __restrict__ msgs, uint32_t n, uint32_t words_per_msg, uint32_t* __restrict__ out, uint32_t seed) {
  extern __shared__ uint32_t tab[];           // [BUCKETS][32]
  for (uint32_t i = threadIdx.x; i < BUCKETS * 32; i += T) tab[i] = (i * 2654435761u) ^ seed;   // "random" fingerprints
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = T / 32;
  uint32_t tabs = (uint32_t)__cvta_generic_to_shared(tab) + lane * 4;
  uint32_t hits = 0;
  for (uint32_t tile = blockIdx.x * wpb + warp; tile * 32 < n; tile += gridDim.x * wpb) {
    uint32_t m = tile * 32 + lane; if (m >= n) continue;
    const uint4* p = msgs + (size_t)m * (words_per_msg / 4);
    uint32_t prev = 0, acc = 0;
    for (uint32_t c = 0; c < words_per_msg / 4; c++) {
      uint4 v = __ldg(p + c);
      uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t f = (wd[q] & 0x1f1f1f1fu) | ((wd[q] >> 1) & 0x20202020u);     // fold6
#pragma unroll
        for (int k = 0; k < 4; k++) {
          uint32_t win = __funnelshift_r(prev, f, 8 * (k + 1) & 31);           // last 4 folded bytes ending at byte k
          if (k == 3) win = f;
          uint32_t h = win * 0x9E3779B1u;
          uint32_t bucket; asm("ld.shared.u32 %0, [%1];" : "=r"(bucket) : "r"(tabs + ((h >> 22) << 7)));
          uint32_t fp2 = __byte_perm(h, 0, 0x1010);                             // fingerprint in both halves
          uint32_t x = bucket ^ fp2;
          acc |= (x - 0x00010001u) & ~x & 0x80008000u;                          // zero half-word <=> fingerprint match
        }
        prev = f;
      }
    }
    hits += acc != 0;
  }
  if (hits) atomicAdd(out, hits);
}
int main() {
  const uint32_t n = 1 << 20, L = 256;
  uint4* d; uint32_t* o; cudaMalloc(&d, (size_t)n * L); cudaMalloc(&o, 4); cudaMemset(o, 0, 4);
  cudaMemset(d, 0x61, (size_t)n * L);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, BUCKETS * 128);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; it++) probe<<<148, T, BUCKETS * 128>>>(d, n, L / 4, o, it);
  cudaEventRecord(e0);
  for (int it = 0; it < 20; it++) probe<<<148, T, BUCKETS * 128>>>(d, n, L / 4, o, it);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
  printf("fingerprint probe per byte: %.3f ms per 1Mi x 256B  -> %.1f GB/s  (err=%s)\n", ms, n * (double)L / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
