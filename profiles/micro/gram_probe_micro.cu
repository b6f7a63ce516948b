// gram_probe_micro.cu -- prototype of the stateless level-1 filter that replaces the DFA walk of scan_kernel:
// position-parallel (a lane owns 16 consecutive bytes of the batch buffer, LDG.128 or TMA-staged), every aligned
// 4-byte gram (stride 4, NPROBE = 1) or every even-offset gram (stride 2, NPROBE = 2) is folded (case bit and
// bit 7 dropped, digits collapsed), hashed multiplicatively and tested against a 64 KB bitmap in shared memory.
// Flagged probes are compacted into a per-warp ring (ballot + popc) and drained 32 at a time by a stand-in for
// the level-1b check (re-read 32 bytes of text from L2, one random LDS.64, eight dependent LDS + tests).
// What it answers: instructions and shared-memory wavefronts per byte are guesses until measured -- how many GB/s
// does this loop sustain on one B200 as a function of NPROBE, the fold, the flag rate and the staging path?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gram_probe_micro gram_probe_micro.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define FULL 0xffffffffu
constexpr int kThreads = 1024, kWarps = kThreads / 32;
constexpr uint32_t kBitmapBytes = 64 * 1024, kL1bBytes = 32 * 1024, kRing = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
  asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(bar), "r"(phase) : "memory");
}

template <int COLLAPSE> __device__ __forceinline__ uint32_t fold_word(uint32_t w) {
  if (!COLLAPSE) return w & 0x5f5f5f5fu;
  // z = (w & 0x5f) ^ 0x10 per byte: digits (and : ; < = > ?) become 0x00..0x0f; z + 0x70 has bit 7 clear exactly for those bytes;
  // PRMT in sign-replicate mode turns bit 7 into a byte mask; the low nibble of the digit bytes is dropped.
  const uint32_t z = (w & 0x5f5f5f5fu) ^ 0x10101010u;
  const uint32_t nd = __byte_perm(z + 0x70707070u, 0, 0xba98);    // 0xff for every byte that is NOT a digit
  return z & (nd | 0xf0f0f0f0u);
}

// one probe: bit (31 - (h & 31)) of bitmap word (h >> 18); the result bit is shifted into `flags`
__device__ __forceinline__ void probe(uint32_t bm, uint32_t g, uint32_t mult, uint32_t& flags) {
  const uint32_t h = g * mult;
  const uint32_t wv = lds32(bm + ((h >> 16) & 0xfffcu));
  const uint32_t t = wv << (h & 31u);
  flags = __funnelshift_l(t, flags, 1);
}

struct Drain {
  const uint8_t* text; uint32_t l1b; uint32_t* counters; uint32_t mult;
  // stand-in for level 1b: one event per lane
  __device__ __forceinline__ void run(uint32_t ring, uint32_t head, uint32_t n, uint32_t lane, int shift) {
    if (lane < n) {
      const uint32_t ev = lds32(ring + (((head + lane) & (kRing - 1)) << 2));
      const size_t p = (size_t)ev << shift;                 // byte position of the gram
      const uint4 a = *reinterpret_cast<const uint4*>(text + (p & ~(size_t)15));
      const uint4 b = *reinterpret_cast<const uint4*>(text + (p & ~(size_t)15) + 16);
      uint32_t h = (a.x ^ b.y) * mult;
      uint32_t e = lds32(l1b + ((h >> 17) & 0x7ffcu));
      uint32_t ok = 1;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t by = ((j & 1 ? b.x : a.w) >> (8 * (j & 3))) & 0x3fu;
        const uint32_t m = lds32(l1b + ((e + 8u * j + by) & 0x7ffcu));
        ok &= (m >> by) | (e >> j);
      }
      if (ok & 1u) atomicAdd(&counters[1], 1u);
    }
  }
};

// MODE 0: LDG.128 straight into registers, DEPTH warp-tiles (512 B each) in flight per warp
// MODE 1: CTA-wide TMA pipeline: 16 KB tiles (one 512-byte slice per warp), kStages deep, full barrier per stage, the last
//         warp to finish a stage refills it
constexpr uint32_t kStagesC = 4, kTileC = 16384, kTileStride = 16512;
template <int NPROBE, int COLLAPSE, int MODE, int DEPTH>
__global__ void __launch_bounds__(kThreads, 1)
l1a_kernel(const uint4* __restrict__ text, uint32_t n_iter /* 512-byte warp tiles */, const uint32_t* __restrict__ bitmap_g,
           const uint32_t* __restrict__ l1b_g, uint32_t mult, uint32_t* __restrict__ counters) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(smem);
  uint32_t* l1b = reinterpret_cast<uint32_t*>(smem + kBitmapBytes);
  for (uint32_t i = threadIdx.x; i < kBitmapBytes / 4; i += kThreads) bitmap[i] = bitmap_g[i];
  for (uint32_t i = threadIdx.x; i < kL1bBytes / 4; i += kThreads) l1b[i] = l1b_g[i];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ring = smem_u32(smem + kBitmapBytes + kL1bBytes) + warp * kRing * 4;
  uint8_t* stage_base = smem + kBitmapBytes + kL1bBytes + kWarps * kRing * 4;
  __shared__ __align__(8) uint64_t full_bar[kStagesC];
  __shared__ uint32_t done_cnt[kStagesC];
  if (MODE == 1 && threadIdx.x == 0) {
    for (uint32_t s = 0; s < kStagesC; s++) { mbar_init(smem_u32(&full_bar[s]), 1); done_cnt[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t bm = smem_u32(bitmap);
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t gw = blockIdx.x * kWarps + warp, nw = gridDim.x * kWarps;
  uint32_t ring_n = 0, ring_head = 0, flagged_total = 0;
  Drain drain{reinterpret_cast<const uint8_t*>(text), smem_u32(l1b), counters, mult};
  constexpr int kShift = NPROBE == 2 ? 1 : 2;

  auto process = [&](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t chunk) {
    const uint32_t f0 = fold_word<COLLAPSE>(w0), f1 = fold_word<COLLAPSE>(w1), f2 = fold_word<COLLAPSE>(w2), f3 = fold_word<COLLAPSE>(w3);
    uint32_t flags = 0;
    if (NPROBE == 2) {
      const uint32_t f4 = fold_word<COLLAPSE>(w4);
      probe(bm, f0, mult, flags); probe(bm, __funnelshift_r(f0, f1, 16), mult, flags);
      probe(bm, f1, mult, flags); probe(bm, __funnelshift_r(f1, f2, 16), mult, flags);
      probe(bm, f2, mult, flags); probe(bm, __funnelshift_r(f2, f3, 16), mult, flags);
      probe(bm, f3, mult, flags); probe(bm, __funnelshift_r(f3, f4, 16), mult, flags);
    } else {
      probe(bm, f0, mult, flags); probe(bm, f1, mult, flags); probe(bm, f2, mult, flags); probe(bm, f3, mult, flags);
    }
    // flags: probe j of this chunk is bit (NP - 1 - j)
    if (__any_sync(FULL, flags != 0)) {
      const uint32_t base = chunk * (4u * NPROBE);      // event = gram index (position >> kShift)
      do {
        const bool has = flags != 0;
        const uint32_t m = __ballot_sync(FULL, has);
        if (has) {
          const uint32_t b = 31u - __clz(flags);
          flags &= ~(1u << b);
          const uint32_t slot = (ring_head + ring_n + __popc(m & lt)) & (kRing - 1);
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(ring + slot * 4), "r"(base + (4u * NPROBE - 1u - b)) : "memory");
        }
        ring_n += __popc(m);
        flagged_total += __popc(m);
        if (ring_n >= 32) { __syncwarp(); drain.run(ring, ring_head, 32, lane, kShift); ring_head += 32; ring_n -= 32; }
      } while (__any_sync(FULL, flags != 0));
    }
  };

  if (MODE == 0) {
    uint4 buf[DEPTH]; uint32_t tail[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
      const uint32_t it = gw + d * nw;
      tail[d] = 0;
      if (it < n_iter) { buf[d] = ldg_stream(text + (size_t)it * 32 + lane); if (NPROBE == 2 && lane == 31) tail[d] = text[(size_t)it * 32 + 32].x; }
    }
    for (uint32_t it0 = gw; it0 < n_iter; it0 += DEPTH * nw) {
#pragma unroll
      for (int d = 0; d < DEPTH; d++) {
        const uint32_t it = it0 + d * nw;
        if (it >= n_iter) break;
        const uint4 cur = buf[d]; const uint32_t ct = tail[d];
        const uint32_t nit = it + DEPTH * nw;
        if (nit < n_iter) { buf[d] = ldg_stream(text + (size_t)nit * 32 + lane); if (NPROBE == 2 && lane == 31) tail[d] = text[(size_t)nit * 32 + 32].x; }
        uint32_t w4 = 0;
        if (NPROBE == 2) { w4 = __shfl_down_sync(FULL, cur.x, 1); if (lane == 31) w4 = ct; }
        process(cur.x, cur.y, cur.z, cur.w, w4, it * 32u + lane);
      }
    }
  } else {
    const uint32_t ntiles = n_iter / 32;
    if (threadIdx.x == 0) for (uint32_t s = 0; s < kStagesC; s++) {
      const uint32_t t = blockIdx.x + s * gridDim.x;
      if (t < ntiles) { mbar_expect_tx(smem_u32(&full_bar[s]), kTileC + 16); tma_g2s(smem_u32(stage_base + s * kTileStride), text + (size_t)t * 1024, kTileC + 16, smem_u32(&full_bar[s])); }
    }
    uint32_t j = 0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, j++) {
      const uint32_t s = j % kStagesC, ph = (j / kStagesC) & 1u;
      mbar_wait(smem_u32(&full_bar[s]), ph);
      const uint32_t a = smem_u32(stage_base + s * kTileStride) + warp * 512 + lane * 16;
      const uint4 v = lds128(a);
      const uint32_t w4 = NPROBE == 2 ? lds32(a + 16) : 0u;
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&done_cnt[s], 1u) == kWarps - 1) {
          done_cnt[s] = 0;
          const uint32_t tn = t + kStagesC * gridDim.x;
          if (tn < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(smem_u32(&full_bar[s]), kTileC + 16);
            tma_g2s(smem_u32(stage_base + s * kTileStride), text + (size_t)tn * 1024, kTileC + 16, smem_u32(&full_bar[s]));
          }
        }
      }
      process(v.x, v.y, v.z, v.w, w4, t * 1024u + warp * 32u + lane);
    }
  }
  __syncwarp();
  if (ring_n) drain.run(ring, ring_head, ring_n, lane, kShift);
  if (lane == 0 && flagged_total) atomicAdd(&counters[0], flagged_total);
}

__global__ void gen_text(uint8_t* t, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t r = (uint32_t)i * 2654435761u; r ^= r >> 15; r *= 2246822519u; r ^= r >> 13; r *= 3266489917u; r ^= r >> 16;
    t[i] = (r % 7u == 0) ? ' ' : (uint8_t)('a' + (r >> 8) % 26u);
  }
}
__global__ void copy_read(const uint4* __restrict__ t, size_t n16, uint32_t* out) {
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) { uint4 v = ldg_stream(t + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) out[0] = acc;
}

template <int NP, int C, int T, int D>
static void run(const char* name, const uint4* d_text, size_t bytes, const uint32_t* d_bm, const uint32_t* d_l1b, uint32_t* d_cnt) {
  const size_t smem = kBitmapBytes + kL1bBytes + kWarps * kRing * 4 + (T ? kStagesC * kTileStride : 0) + 128;
  cudaFuncSetAttribute(l1a_kernel<NP, C, T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const uint32_t n_iter = (uint32_t)(bytes / 512);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  uint32_t cnt[2] = {0, 0};
  for (int r = 0; r < 6; r++) {
    cudaMemset(d_cnt, 0, 8);
    cudaEventRecord(e0);
    l1a_kernel<NP, C, T, D><<<148, kThreads, smem>>>(d_text, n_iter, d_bm, d_l1b, 0x9E3779B1u, d_cnt);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 1 && ms < best) best = ms;
  }
  cudaMemcpy(cnt, d_cnt, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %8.1f us  %7.1f GB/s  flagged %u (%.3f per 256 B)  survivors %u  (%s)\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9, cnt[0],
         cnt[0] / (bytes / 256.0), cnt[1], cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const size_t bytes = (size_t)1 << 28;     // 1 Mi x 256 B
  uint8_t* d_text; uint32_t *d_bm, *d_l1b, *d_cnt;
  cudaMalloc(&d_text, bytes + 4096); cudaMalloc(&d_bm, kBitmapBytes); cudaMalloc(&d_l1b, kL1bBytes); cudaMalloc(&d_cnt, 64);
  gen_text<<<148 * 8, 256>>>(d_text, bytes + 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); copy_read<<<148 * 16, 256>>>(reinterpret_cast<const uint4*>(d_text), bytes / 16, d_cnt); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r == 2) printf("read-only stream             %8.1f us  %7.1f GB/s\n", ms * 1e3, bytes / (ms * 1e-3) / 1e9);
  }
  std::vector<uint32_t> l1b(kL1bBytes / 4);
  srand(7); for (auto& x : l1b) x = (uint32_t)rand() * 2654435761u;
  cudaMemcpy(d_l1b, l1b.data(), kL1bBytes, cudaMemcpyHostToDevice);
  const double dens[4] = {0.0, 0.003, 0.013, 0.05};
  for (int di = 0; di < 4; di++) {
    std::vector<uint32_t> bm(kBitmapBytes / 4, 0);
    srand(11);
    const size_t nbits = (size_t)(dens[di] * kBitmapBytes * 8);
    for (size_t i = 0; i < nbits; i++) { uint32_t b = ((uint32_t)rand() * 32768u + (uint32_t)rand()) % (kBitmapBytes * 8); bm[b >> 5] |= 1u << (b & 31); }
    cudaMemcpy(d_bm, bm.data(), kBitmapBytes, cudaMemcpyHostToDevice);
    printf("--- bitmap density %.3f\n", dens[di]);
    const uint4* t = reinterpret_cast<const uint4*>(d_text);
    run<1, 1, 0, 1>("stride4 ldg d1", t, bytes, d_bm, d_l1b, d_cnt);
    run<1, 1, 0, 2>("stride4 ldg d2", t, bytes, d_bm, d_l1b, d_cnt);
    run<1, 1, 0, 3>("stride4 ldg d3", t, bytes, d_bm, d_l1b, d_cnt);
    run<1, 1, 0, 4>("stride4 ldg d4", t, bytes, d_bm, d_l1b, d_cnt);
    run<1, 1, 1, 1>("stride4 tma cta", t, bytes, d_bm, d_l1b, d_cnt);
    run<2, 1, 0, 1>("stride2 ldg d1", t, bytes, d_bm, d_l1b, d_cnt);
    run<2, 1, 0, 3>("stride2 ldg d3", t, bytes, d_bm, d_l1b, d_cnt);
    run<2, 0, 0, 3>("stride2 ldg d3 nocollapse", t, bytes, d_bm, d_l1b, d_cnt);
    run<2, 1, 1, 1>("stride2 tma cta", t, bytes, d_bm, d_l1b, d_cnt);
  }
  return 0;
}
