// scan_kernels.cu -- sm_100a kernels of the message-scan hot path.
//
//   scan_kernel     every byte of every message through the prefilter DFA that lives in shared
//                   memory (image staged with one TMA bulk copy per CTA, completion on an mbarrier).
//                   One lane owns one message; candidates (message, rule) are queued in HBM.
//   verify_kernel   exact ECMAScript semantics for the queued candidates: a Pike VM over UTF-16
//                   units decoded on the fly from the UTF-8 bytes (leftmost-first, global-exec
//                   iteration of registry.ts:225-236, RegExp.test of context.ts:9-25).
//   finalize_kernel per-message result words from the verified-hit bitmaps.
//
// Replaces the two hot loops of the reference: gov/src/conditions/context.ts:9-25 (rules x
// RegExp.test) and gov/src/redaction/registry.ts:212-242 (P global-exec passes per string).
#include "kernels.h"
#include "rulec.h"
#include "pike_vm.h"

namespace cg {

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copy (global -> shared)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ uint4 ldg_stream(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ------------------------------------------------------------------------------------------
// candidate queue (lane-private slot; cand bitmap dedupes (message, rule) at the source)
// ------------------------------------------------------------------------------------------
struct Emitter {
  const DevRuleset& rs; const ScanWork& w; uint32_t msg; uint32_t slot;
  __device__ Emitter(const DevRuleset& r, const ScanWork& wk, uint32_t m) : rs(r), w(wk), msg(m), slot(0xffffffffu) {}
  __device__ bool ensure_slot() {
    if (slot != 0xffffffffu) return slot < w.slot_cap;
    slot = atomicAdd(&w.counters[0], 1u);
    if (slot >= w.slot_cap) { atomicOr(&w.counters[3], ERR_SLOT_OVERFLOW); return false; }
    w.slot_msg[slot] = msg;
    for (uint32_t k = 0; k < rs.rw; k++) { w.cand[(size_t)slot * rs.rw + k] = 0; w.hit[(size_t)slot * rs.rw + k] = 0; }
    return true;
  }
  __device__ void rule(uint32_t r) {
    if (!ensure_slot()) return;
    uint32_t* cw = &w.cand[(size_t)slot * rs.rw + (r >> 5)];
    uint32_t v = *cw, bit = 1u << (r & 31);
    if (v & bit) return;
    *cw = v | bit;
    uint32_t e = atomicAdd(&w.counters[1], 1u);
    if (e < w.event_cap) w.events[e] = make_uint2(slot, r); else atomicOr(&w.counters[3], ERR_EVENT_OVERFLOW);
  }
  __device__ void accept_state(uint32_t s) {
    uint32_t a = s - rs.first_accept;
    for (uint32_t k = rs.out_offsets[a]; k < rs.out_offsets[a + 1]; k++) rule(rs.out_rules[k]);
  }
};

// ------------------------------------------------------------------------------------------
// prefilter scan, v1: lane-per-message, streaming 16-byte loads, table in shared memory
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 512;

template <int MODE>
__device__ __forceinline__ uint32_t dfa_step(const uint16_t* __restrict__ table, const uint8_t* __restrict__ lut,
                                             uint32_t state, uint32_t b, uint32_t ncols_log2) {
  if (MODE == 0) return table[(state << 7) + (b & 0x7fu)];
  return table[(state << ncols_log2) + lut[b]];
}

template <int MODE>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_kernel(DevRuleset rs, ScanWork w, const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ off, uint32_t n,
            uint64_t* __restrict__ words) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, rs.image_bytes);
    for (uint32_t o = 0; o < rs.image_bytes; o += 32768u) {
      uint32_t len = rs.image_bytes - o < 32768u ? rs.image_bytes - o : 32768u;
      tma_bulk_g2s(smem + o, rs.image + o, len, &bar);
    }
  }
  mbar_wait(&bar, 0);
  const uint8_t* lut = smem;
  const uint16_t* table = reinterpret_cast<const uint16_t*>(smem + 256);
  const uint32_t ncl = rs.ncols_log2, first_accept = rs.first_accept;

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = kScanThreads / 32;
  const uint32_t ntiles = (n + 31) / 32;
  for (uint32_t tile = blockIdx.x * wpb + warp; tile < ntiles; tile += gridDim.x * wpb) {
    uint32_t msg = tile * 32 + lane;
    if (msg >= n) continue;
    uint32_t b = off[msg], e = off[msg + 1];
    uint32_t state = 0, acc = 0, p = b;
    while (p < e && (p & 15u)) { state = dfa_step<MODE>(table, lut, state, bytes[p], ncl); acc = max(acc, state); p++; }
    while (p + 16 <= e) {
      uint4 v = ldg_stream(bytes + p);
      uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          state = dfa_step<MODE>(table, lut, state, (wd[q] >> (8 * k)) & 0xffu, ncl);
          acc = max(acc, state);
        }
      }
      p += 16;
    }
    while (p < e) { state = dfa_step<MODE>(table, lut, state, bytes[p], ncl); acc = max(acc, state); p++; }

    if (acc >= first_accept || rs.n_always) {
      Emitter em(rs, w, msg);
      for (uint32_t k = 0; k < rs.n_always; k++) em.rule(rs.always_rules[k]);
      if (acc >= first_accept) {
        state = 0;
        for (p = b; p < e; p++) {
          state = dfa_step<MODE>(table, lut, state, bytes[p], ncl);
          if (state >= first_accept) em.accept_state(state);
        }
      }
    }
    words[msg] = 0ull;
  }
}

constexpr int kVerifyThreads = 64;

struct GlobalSpanSink {
  const ScanWork& w; uint32_t msg, rule;
  __device__ void span(uint32_t sb, uint32_t eb, uint32_t s16, uint32_t e16) {
    uint32_t k = atomicAdd(&w.counters[2], 1u);
    if (k < w.span_cap) { uint32_t* o = w.spans + (size_t)k * 6; o[0] = msg; o[1] = rule; o[2] = sb; o[3] = eb; o[4] = s16; o[5] = e16; }
    else atomicOr(&w.counters[3], ERR_SPAN_OVERFLOW);
  }
};

template <bool SPANS>
__global__ void __launch_bounds__(kVerifyThreads)
verify_kernel(DevRuleset rs, ScanWork w, const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ off) {
  const uint32_t n_events = min(w.counters[1], w.event_cap);
  VM vm(rs);
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_events; e += gridDim.x * blockDim.x) {
    uint2 ev = w.events[e];
    uint32_t slot = ev.x, rule = ev.y, msg = w.slot_msg[slot];
    GlobalSpanSink sink{w, msg, rule};
    bool any = run_rule<SPANS>(vm, rs, rule, bytes + off[msg], off[msg + 1] - off[msg], sink);
    if (any) atomicOr(&w.hit[(size_t)slot * rs.rw + (rule >> 5)], 1u << (rule & 31));
  }
  if (vm.err) atomicOr(&w.counters[3], vm.err);
}

__global__ void finalize_kernel(DevRuleset rs, ScanWork w, uint64_t* __restrict__ words) {
  const uint32_t n_slots = min(w.counters[0], w.slot_cap);
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
    uint32_t count = 0, first = 0xffffffffu;
    for (uint32_t k = 0; k < rs.rw; k++) {
      uint32_t v = w.hit[(size_t)s * rs.rw + k];
      if (v) { if (first == 0xffffffffu) first = k * 32 + (__ffs(v) - 1); count += __popc(v); }
    }
    if (count) words[w.slot_msg[s]] = (1ull << 63) | ((uint64_t)count << 32) | first;
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int launch_scan(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off, uint32_t n,
                uint64_t* d_words, int sm_count, cudaStream_t stream) {
  if (n == 0) return 0;
  size_t smem = rs.image_bytes;
  uint32_t ntiles = (n + 31) / 32, wpb = kScanThreads / 32;
  uint32_t grid = (ntiles + wpb - 1) / wpb; if (grid > (uint32_t)sm_count) grid = sm_count;
  if (rs.mode == 0) {
    cudaFuncSetAttribute(scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_kernel<0><<<grid, kScanThreads, smem, stream>>>(rs, w, d_bytes, d_off, n, d_words);
  } else {
    cudaFuncSetAttribute(scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scan_kernel<1><<<grid, kScanThreads, smem, stream>>>(rs, w, d_bytes, d_off, n, d_words);
  }
  return 1;
}

int launch_verify(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off,
                  bool want_spans, int sm_count, cudaStream_t stream) {
  int grid = sm_count * 8;
  if (want_spans) verify_kernel<true><<<grid, kVerifyThreads, 0, stream>>>(rs, w, d_bytes, d_off);
  else verify_kernel<false><<<grid, kVerifyThreads, 0, stream>>>(rs, w, d_bytes, d_off);
  return 1;
}

int launch_finalize(const DevRuleset& rs, const ScanWork& w, uint64_t* d_words, int sm_count, cudaStream_t stream) {
  finalize_kernel<<<sm_count * 2, 256, 0, stream>>>(rs, w, d_words);
  return 1;
}

}  // namespace cg
