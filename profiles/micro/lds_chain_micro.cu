// lds_chain_micro.cu -- what does one dependent shared-memory table step cost on sm_100a, as a function of the
// bank-conflict pattern?  Every warp walks a chain  v = smem16[v * mul + lane_off]  (PRMT-free: IMAD + LDS.U16),
// 32 warps per SM, one CTA per SM -- the shape of scan_kernel's level-1 walk.
//   pattern 0: every lane reads the same address (broadcast)
//   pattern 1: lane-private banks (bank == lane): conflict-free, 32 distinct addresses
//   pattern 2: random rows, padded stride 132 and a per-lane random column: random banks (what scan_kernel does)
//   pattern 3: like 2 but only every 2nd lane active / 4: every 4th lane active
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_chain_micro lds_chain_micro.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) chain(const uint16_t* __restrict__ init, uint32_t bytes, uint32_t mul, const uint32_t* __restrict__ lane_off,
                                                 uint32_t lane_mask, int steps, uint32_t* out) {
  extern __shared__ __align__(16) uint8_t smem[];
  for (uint32_t i = threadIdx.x; i < bytes / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = init[i];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t off = lane_off[threadIdx.x];
  uint32_t v = threadIdx.x % 7;
  if ((lane_mask >> lane) & 1u) {
#pragma unroll 16
    for (int i = 0; i < steps; i++) v = *reinterpret_cast<const uint16_t*>(smem + v * mul + off);
  }
  if (v == 0xffffffffu) out[0] = v;
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = v;
}

int main(int argc, char** argv) {
  const int steps = 16384, rows = 1500;
  const uint32_t bytes = 200 * 1024;
  cudaFuncSetAttribute(chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  std::vector<uint16_t> t(bytes / 2);
  std::vector<uint32_t> lo(1024);
  uint16_t* d_t; uint32_t *d_lo, *d_out;
  cudaMalloc(&d_t, bytes); cudaMalloc(&d_lo, 4096); cudaMalloc(&d_out, 4 * (1 + 148 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int pat = 0; pat < 5; pat++) for (int threads = 1024; threads >= 512; threads /= 2) {
    srand(1234);
    uint32_t mul = 132, mask = 0xffffffffu;
    if (pat == 0) { for (auto& x : t) x = 0; for (auto& x : lo) x = 0; }
    else if (pat == 1) { mul = 128; for (auto& x : t) x = (uint16_t)(rand() % rows); for (int i = 0; i < 1024; i++) lo[i] = (i & 31) * 4 + 2 * (rand() & 1); }
    else { for (auto& x : t) x = (uint16_t)(rand() % rows); for (int i = 0; i < 1024; i++) lo[i] = 2 * (rand() % 64); mask = pat == 2 ? 0xffffffffu : pat == 3 ? 0x55555555u : 0x11111111u; }
    cudaMemcpy(d_t, t.data(), bytes, cudaMemcpyHostToDevice); cudaMemcpy(d_lo, lo.data(), 4096, cudaMemcpyHostToDevice);
    chain<<<148, threads, bytes>>>(d_t, bytes, mul, d_lo, mask, steps, d_out);
    cudaEventRecord(e0);
    chain<<<148, threads, bytes>>>(d_t, bytes, mul, d_lo, mask, steps, d_out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * clk_khz * 1e3;
    printf("pattern %d warps/SM %2d  %.3f ms  %.2f cycles per warp-LDS per SM (%s)\n", pat, threads / 32, ms, cyc / ((double)steps * (threads / 32)),
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
