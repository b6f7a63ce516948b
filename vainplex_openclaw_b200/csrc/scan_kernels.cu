// scan_kernels.cu -- sm_100a kernels of the message-scan hot path.
//
//   scan_kernel     every byte of every message through the level-1 prefilter DFA that lives in
//                   shared memory (image staged with TMA bulk copies, completion on an mbarrier).
//                   One lane owns one message; accepting transitions are compacted with warp
//                   ballots into an event queue in HBM (one atomic per warp and word).
//   confirm_kernel  level 2: one thread per level-1 event confirms the full factor exactly and
//                   queues the surviving (message, rule) pairs (or records a direct hit).
//   verify_kernel   exact ECMAScript semantics for the queued pairs: a Pike VM over UTF-16 units
//                   decoded on the fly from the UTF-8 bytes (leftmost-first, global-exec
//                   iteration of registry.ts:225-236, RegExp.test of context.ts:9-25).
//   finalize_kernel per-message result words from the verified-hit bitmaps.
//
// Replaces the two hot loops of the reference: gov/src/conditions/context.ts:9-25 (rules x
// RegExp.test) and gov/src/redaction/registry.ts:212-242 (P global-exec passes per string).
#include "kernels.h"
#include "rulec.h"
#include "pike_vm.h"
#include "prefilter_dev.h"
#include <algorithm>

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
// level 1: prefilter scan.  lane-per-message (NS interleaved message streams per lane), streaming
// 16-byte loads, table in shared memory
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr uint32_t kL1Always = 0xffffffffu;
constexpr uint32_t kAccept = 0x8000u, kCold = 0x4000u, kStateMask = 0x3fffu;

// pre-doubled column indices of the four bytes of a word (one byte each), SWAR
template <int MODE> __device__ __forceinline__ uint32_t cols2_of_word(uint32_t w);
template <> __device__ __forceinline__ uint32_t cols2_of_word<0>(uint32_t w) { return (w + w) & 0xfefefefeu; }
template <> __device__ __forceinline__ uint32_t cols2_of_word<2>(uint32_t w) { return ((w & 0x1f1f1f1fu) | ((w >> 1) & 0x20202020u)) << 1; }
template <> __device__ __forceinline__ uint32_t cols2_of_word<3>(uint32_t w) { return (w & 0x1f1f1f1fu) << 1; }
template <> __device__ __forceinline__ uint32_t cols2_of_word<1>(uint32_t w) { return w; }   // LUT mode resolves per byte

template <int MODE> struct RowBytes { static constexpr uint32_t v = MODE == 0 ? 256u : MODE == 3 ? 64u : 128u; };

__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) { uint16_t v; asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }

// One level-1 transition out of the shared-memory image (32-bit shared addresses, no generic
// loads).  Rows [0, hot) are real; row `hot` is a trap row (every entry = hot | kCold) and every
// hot-row transition into a deep state is stored as hot | kCold, so the fast path never leaves
// shared memory: an excursion into deep states only raises kCold, and the flagged word is then
// re-walked on the full table in HBM/L2 by the (rare) slow path, which also restores the state.
// In mode 2 the 32 words of a row are XOR-swizzled by the low five bits of the row index (the image is
// stored that way): lanes sitting in different states but reading the same frequent column (' ', 'e',
// ...) then land in different banks instead of all colliding in one.
template <int MODE>
__device__ __forceinline__ uint32_t l1_fast(uint32_t tbl_s, uint32_t lut_s, uint32_t row_shift, uint32_t state, uint32_t c2) {
  if (MODE == 1) return lds_u16(tbl_s + (state << row_shift) + 2u * lds_u8(lut_s + c2));
  if (MODE == 2) return lds_u16(tbl_s + state * 128u + (c2 ^ ((state << 2) & 0x7cu)));
  return lds_u16(tbl_s + state * RowBytes<MODE>::v + c2);
}

// the complete table (deep states included), L2-resident
__device__ __forceinline__ uint32_t l1_full(const DevRuleset& rs, uint32_t state, uint32_t col) {
  return __ldg(rs.table_full + ((size_t)state << rs.ncols_log2) + col);
}

__device__ __forceinline__ void l1_push_one(const ScanWork& w, uint32_t msg, uint32_t pos, uint32_t sc) {
  uint32_t k = atomicAdd(&w.counters[4], 1u);
  if (k < w.l1_cap) { w.l1_msg[k] = msg; w.l1_pos[k] = pos; w.l1_sc[k] = sc; } else atomicOr(&w.counters[3], ERR_L1_OVERFLOW);
}

template <int MODE, int NS>
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
  const uint32_t lut_s = smem_u32(smem), tbl_s = lut_s + 256u;
  const uint32_t row_shift = rs.ncols_log2 + 1, hot = rs.hot_states;
  const uint32_t FULL = 0xffffffffu;

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = kScanThreads / 32;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t per_tile = 32u * NS;
  const uint32_t ntiles = (n + per_tile - 1) / per_tile;
  for (uint32_t tile = blockIdx.x * wpb + warp; tile < ntiles; tile += gridDim.x * wpb) {
    uint32_t msg[NS], b[NS], e[NS], p[NS], state[NS], nch[NS];
    uint32_t maxch = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) {
      msg[s] = tile * per_tile + s * 32 + lane;
      const bool valid = msg[s] < n;
      b[s] = valid ? off[msg[s]] : 0u; e[s] = valid ? off[msg[s] + 1] : 0u;
      if (valid && rs.n_always) l1_push_one(w, msg[s], 0, kL1Always);
      state[s] = 0; p[s] = b[s];
      // unaligned head: byte-wise on the full table
      uint32_t head_end = (b[s] + 15u) & ~15u; if (head_end > e[s]) head_end = e[s];
      for (; p[s] < head_end; p[s]++) {
        uint32_t col = l1_col(rs.mode, lut, bytes[p[s]]);
        uint32_t ent = l1_full(rs, state[s], col);
        if (ent & kAccept) l1_push_one(w, msg[s], p[s] - b[s], (state[s] << 8) | col);
        state[s] = ent & kStateMask;
      }
      nch[s] = (e[s] - p[s]) >> 4;
      maxch = max(maxch, nch[s]);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) maxch = max(maxch, __shfl_xor_sync(FULL, maxch, d));

    // warp-uniform main loop over 16-byte chunks of NS message streams; the chunks are software-
    // pipelined kPrefetch deep in registers so the HBM latency of chunk c+kPrefetch hides behind
    // the table walk of chunk c
    constexpr int kPrefetch = 3;
    uint4 buf[NS][kPrefetch];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
      for (int d = 0; d < kPrefetch; d++) { buf[s][d] = make_uint4(0, 0, 0, 0); if ((uint32_t)d < nch[s]) buf[s][d] = ldg_stream(bytes + p[s] + 16 * d); }
    for (uint32_t c = 0; c < maxch; c++) {
      uint4 v[NS]; bool act[NS];
#pragma unroll
      for (int s = 0; s < NS; s++) {
        act[s] = c < nch[s]; v[s] = buf[s][0];
#pragma unroll
        for (int d = 0; d + 1 < kPrefetch; d++) buf[s][d] = buf[s][d + 1];
        buf[s][kPrefetch - 1] = make_uint4(0, 0, 0, 0);
        if (c + kPrefetch < nch[s]) buf[s][kPrefetch - 1] = ldg_stream(bytes + p[s] + 16 * kPrefetch);
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t s0[NS], acc[NS], wd[NS], fs[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
          wd[s] = q == 0 ? v[s].x : q == 1 ? v[s].y : q == 2 ? v[s].z : v[s].w;
          s0[s] = state[s];
          acc[s] = state[s] >= hot ? kCold : 0u;          // a deep state at word start: this word goes the slow way
          fs[s] = min(state[s], hot);
        }
        uint32_t c2[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) c2[s] = cols2_of_word<MODE>(wd[s]);
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
          for (int s = 0; s < NS; s++) {
            uint32_t ent = l1_fast<MODE>(tbl_s, lut_s, row_shift, fs[s], __byte_perm(c2[s], 0, 0x4440 + k));
            acc[s] |= ent; fs[s] = ent & kStateMask;
          }
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
          const bool flagged = act[s] && (acc[s] & (kAccept | kCold));
          if (act[s]) state[s] = fs[s];
          // rare: an accepting transition or an excursion into deep states in this word
          if (__ballot_sync(FULL, flagged)) {
            uint32_t cnt = 0, ev_pos[4], ev_sc[4];
            if (flagged) {
              uint32_t st = s0[s];
#pragma unroll
              for (int k = 0; k < 4; k++) {
                uint32_t col = l1_col(rs.mode, lut, (wd[s] >> (8 * k)) & 0xffu);
                uint32_t ent = l1_full(rs, st, col);
                if (ent & kAccept) { ev_pos[cnt] = p[s] + 4 * q + k - b[s]; ev_sc[cnt] = (st << 8) | col; cnt++; }
                st = ent & kStateMask;
              }
              state[s] = st;                                 // the true state (may be a deep one)
            }
            uint32_t idx = 0, total = 0;
#pragma unroll
            for (uint32_t j = 1; j <= 4; j++) { uint32_t bj = __ballot_sync(FULL, cnt >= j); idx += __popc(bj & lt_mask); total += __popc(bj); }
            if (total) {
              uint32_t base = 0;
              if (lane == 0) base = atomicAdd(&w.counters[4], total);
              base = __shfl_sync(FULL, base, 0);
              if (base + total > w.l1_cap) { if (lane == 0) atomicOr(&w.counters[3], ERR_L1_OVERFLOW); }
              else {
#pragma unroll
                for (uint32_t j = 0; j < 4; j++) if (j < cnt) { uint32_t k2 = base + idx + j; w.l1_msg[k2] = msg[s]; w.l1_pos[k2] = ev_pos[j]; w.l1_sc[k2] = ev_sc[j]; }
              }
            }
          }
        }
      }
#pragma unroll
      for (int s = 0; s < NS; s++) if (act[s]) p[s] += 16;
    }
    // tails
#pragma unroll
    for (int s = 0; s < NS; s++) {
      for (; p[s] < e[s]; p[s]++) {
        uint32_t col = l1_col(rs.mode, lut, bytes[p[s]]);
        uint32_t ent = l1_full(rs, state[s], col);
        if (ent & kAccept) l1_push_one(w, msg[s], p[s] - b[s], (state[s] << 8) | col);
        state[s] = ent & kStateMask;
      }
      if (msg[s] < n) words[msg[s]] = 0ull;
    }
  }
}

// ------------------------------------------------------------------------------------------
// level 1, mode 4: stateless fingerprint probe.  Every byte position hashes the last four symbols
// (fold6 + digit collapse, SWAR) and reads ONE word from the lane's own bank of the replicated table
// -- conflict-free by construction -- and compares two 16-bit fingerprints.  Up to two single-byte
// triggers (e.g. '@') are matched SWAR-style in registers.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fp_fold_word(uint32_t w) {
  uint32_t f = (w & 0x1f1f1f1fu) | ((w >> 1) & 0x20202020u);
  uint32_t t = (f >> 4) & ~(f >> 5) & 0x01010101u;        // digit columns 0x10..0x19 -> 0x10 / 0x18
  return f & ~(t * 7u);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ uint32_t fp_probe(uint32_t tabs, uint32_t buckets, uint32_t mult, uint32_t win) {
  uint32_t h = win * mult;
  uint32_t x = lds_u32(tabs + (__umulhi(h, buckets) << 7)) ^ __byte_perm(h, 0, 0x2121);        // fingerprint = bits 8..23 of h, in both halves
  return (x - 0x00010001u) & ~x & 0x80008000u;              // non-zero <=> one half equals the fingerprint
}
__device__ __forceinline__ uint32_t has_byte(uint32_t w, uint32_t splat) { uint32_t x = w ^ splat; return (x - 0x01010101u) & ~x & 0x80808080u; }

__global__ void __launch_bounds__(kScanThreads, 1)
scan_fp_kernel(DevRuleset rs, ScanWork w, const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ off, uint32_t n,
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
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = kScanThreads / 32;
  const uint32_t tabs = smem_u32(smem) + 256u + lane * 4u;      // this lane's bank
  const uint32_t B = rs.fp_buckets, mult = rs.fp_mult, ntrig = rs.n_trig;
  const uint32_t sp0 = rs.trig_byte[0] * 0x01010101u, sp1 = rs.trig_byte[1] * 0x01010101u;
  const uint32_t FULL = 0xffffffffu, lt_mask = (1u << lane) - 1u;
  const uint32_t ntiles = (n + 31) / 32;
  for (uint32_t tile = blockIdx.x * wpb + warp; tile < ntiles; tile += gridDim.x * wpb) {
    const uint32_t msg = tile * 32 + lane;
    const bool valid = msg < n;
    const uint32_t b = valid ? off[msg] : 0u, e = valid ? off[msg + 1] : 0u;
    if (valid && rs.n_always) l1_push_one(w, msg, 0, kL1Always);
    uint32_t p = b, win = 0;                                 // win: last four symbols, oldest in the low byte
    auto bytewise = [&](uint32_t upto) {
      for (; p < upto; p++) {
        const uint32_t byte = bytes[p];
        win = (win >> 8) | (fp_fold_word(byte) << 24);
        if (fp_probe(tabs, B, mult, win)) l1_push_one(w, msg, p - b, win * mult);
        for (uint32_t t = 0; t < ntrig; t++) if (byte == rs.trig_byte[t]) l1_push_one(w, msg, (p - b) | 0x80000000u, t);
      }
    };
    uint32_t head_end = (b + 15u) & ~15u; if (head_end > e) head_end = e;
    bytewise(head_end);
    uint32_t nch = (e - p) >> 4, maxch = nch;
#pragma unroll
    for (int d = 16; d; d >>= 1) maxch = max(maxch, __shfl_xor_sync(FULL, maxch, d));
    constexpr int kPrefetch = 3;
    uint4 buf[kPrefetch];
#pragma unroll
    for (int d = 0; d < kPrefetch; d++) { buf[d] = make_uint4(0, 0, 0, 0); if ((uint32_t)d < nch) buf[d] = ldg_stream(bytes + p + 16 * d); }
    for (uint32_t c = 0; c < maxch; c++) {
      const bool act = c < nch;
      const uint4 v = buf[0];
#pragma unroll
      for (int d = 0; d + 1 < kPrefetch; d++) buf[d] = buf[d + 1];
      buf[kPrefetch - 1] = make_uint4(0, 0, 0, 0);
      if (c + kPrefetch < nch) buf[kPrefetch - 1] = ldg_stream(bytes + p + 16 * kPrefetch);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint32_t wd = q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w;
        const uint32_t P = win, F = fp_fold_word(wd);
        uint32_t acc = fp_probe(tabs, B, mult, __funnelshift_r(P, F, 8)) | fp_probe(tabs, B, mult, __funnelshift_r(P, F, 16)) |
                       fp_probe(tabs, B, mult, __funnelshift_r(P, F, 24)) | fp_probe(tabs, B, mult, F);
        if (ntrig) { acc |= has_byte(wd, sp0); if (ntrig > 1) acc |= has_byte(wd, sp1); }
        const bool flagged = act && acc != 0;
        if (act) win = F;
        if (__ballot_sync(FULL, flagged)) {
          uint32_t cnt = 0, ev_pos[6], ev_sc[6];
          if (flagged) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const uint32_t wk = k == 3 ? F : __funnelshift_r(P, F, 8 * (k + 1));
              const uint32_t pos = p + 4 * q + k - b;
              if (fp_probe(tabs, B, mult, wk) && cnt < 6) { ev_pos[cnt] = pos; ev_sc[cnt] = wk * mult; cnt++; }
              const uint32_t byte = (wd >> (8 * k)) & 0xffu;
              for (uint32_t t = 0; t < ntrig; t++) if (byte == rs.trig_byte[t] && cnt < 6) { ev_pos[cnt] = pos | 0x80000000u; ev_sc[cnt] = t; cnt++; }
            }
          }
          uint32_t idx = 0, total = 0;
#pragma unroll
          for (uint32_t j = 1; j <= 6; j++) { uint32_t bj = __ballot_sync(FULL, cnt >= j); idx += __popc(bj & lt_mask); total += __popc(bj); }
          if (total) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&w.counters[4], total);
            base = __shfl_sync(FULL, base, 0);
            if (base + total > w.l1_cap) { if (lane == 0) atomicOr(&w.counters[3], ERR_L1_OVERFLOW); }
            else {
#pragma unroll
              for (uint32_t j = 0; j < 6; j++) if (j < cnt) { uint32_t k2 = base + idx + j; w.l1_msg[k2] = msg; w.l1_pos[k2] = ev_pos[j]; w.l1_sc[k2] = ev_sc[j]; }
            }
          }
        }
      }
      if (act) p += 16;
    }
    bytewise(e);
    if (valid) words[msg] = 0ull;
  }
}

// ------------------------------------------------------------------------------------------
// level 2: confirm the full factor for every level-1 event; queue survivors for the VM
// ------------------------------------------------------------------------------------------
struct SlotSink {
  const DevRuleset& rs; const ScanWork& w; uint32_t msg; uint32_t slot;
  __device__ SlotSink(const DevRuleset& r, const ScanWork& wk, uint32_t m, bool ws) : rs(r), w(wk), msg(m), slot(0xffffffffu), want_spans(ws) {}
  __device__ bool ensure_slot() {
    if (slot != 0xffffffffu) return slot < w.slot_cap;
    uint32_t s = *reinterpret_cast<volatile uint32_t*>(&w.slot_of_msg[msg]);
    if (s == 0xffffffffu) {
      uint32_t mine = atomicAdd(&w.counters[0], 1u);
      if (mine >= w.slot_cap) { atomicOr(&w.counters[3], ERR_SLOT_OVERFLOW); slot = mine; return false; }
      for (uint32_t k = 0; k < rs.rw; k++) { w.cand[(size_t)mine * rs.rw + k] = 0; w.hit[(size_t)mine * rs.rw + k] = 0; }
      w.slot_msg[mine] = msg;
      __threadfence();                                   // rows are zero before the slot becomes visible
      uint32_t old = atomicCAS(&w.slot_of_msg[msg], 0xffffffffu, mine);
      if (old == 0xffffffffu) s = mine; else { s = old; w.slot_msg[mine] = 0xffffffffu; }   // lost the race: slot stays unused
    }
    slot = s;
    return true;
  }
  // spans mode: one VM run per (message, rule), over the whole message (cand bitmap dedupes).
  // policy mode: one VM run per confirmed factor occurrence, restricted to its island.
  bool want_spans;
  __device__ void candidate(uint32_t r, uint32_t t0, uint32_t pre) {
    if (!ensure_slot()) return;
    uint32_t bit = 1u << (r & 31);
    uint32_t old = atomicOr(&w.cand[(size_t)slot * rs.rw + (r >> 5)], bit);
    if (want_spans && (old & bit)) return;
    uint32_t e = atomicAdd(&w.counters[1], 1u);
    if (e < w.event_cap) { w.events[e] = make_uint2(slot, r); w.event_pos[e] = t0; w.event_pre[e] = pre; } else atomicOr(&w.counters[3], ERR_EVENT_OVERFLOW);
  }
  // rules without factors: one whole-message run (event_pos = 0xffffffff)
  __device__ void candidate_always(uint32_t r) {
    if (!ensure_slot()) return;
    uint32_t bit = 1u << (r & 31);
    uint32_t old = atomicOr(&w.cand[(size_t)slot * rs.rw + (r >> 5)], bit);
    if (old & bit) return;
    uint32_t e = atomicAdd(&w.counters[1], 1u);
    if (e < w.event_cap) { w.events[e] = make_uint2(slot, r); w.event_pos[e] = 0xffffffffu; } else atomicOr(&w.counters[3], ERR_EVENT_OVERFLOW);
  }
  // a confirmed exact factor already proves RegExp.test(message) for this rule
  __device__ void direct(uint32_t r) {
    if (!ensure_slot()) return;
    uint32_t bit = 1u << (r & 31);
    atomicOr(&w.cand[(size_t)slot * rs.rw + (r >> 5)], bit);
    atomicOr(&w.hit[(size_t)slot * rs.rw + (r >> 5)], bit);
  }
};

__global__ void __launch_bounds__(256)
confirm_kernel(DevRuleset rs, ScanWork w, const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ off, int want_spans) {
  const uint32_t n1 = min(w.counters[4], w.l1_cap);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
    const uint32_t msg = w.l1_msg[i], pos = w.l1_pos[i], sc = w.l1_sc[i];
    SlotSink sink(rs, w, msg, want_spans != 0);
    if (sc == kL1Always) { for (uint32_t k = 0; k < rs.n_always; k++) sink.candidate_always(rs.always_rules[k]); continue; }
    const uint32_t b = off[msg], len = off[msg + 1] - b;
    if (rs.mode == 4) {
      if (pos & 0x80000000u) accept_id(rs, rs.trig_acc[sc & 1u], bytes + b, len, pos & 0x7fffffffu, want_spans != 0, sink);
      else fp_accept(rs, sc, bytes + b, len, pos, want_spans != 0, sink);
    } else l1_accept(rs, sc >> 8, sc & 0xffu, bytes + b, len, pos, want_spans != 0, sink);
  }
}

// ------------------------------------------------------------------------------------------
// level 3: exact verification by the Pike VM
// ------------------------------------------------------------------------------------------
constexpr int kVerifyThreads = 64;            // verify_large_kernel
// verify_small_kernel: VM runs diverge completely (different program, different text per lane), so a warp
// pays the SUM of its lanes' instruction streams.  Only kVerifyLanes lanes per warp carry a run.
constexpr int kVerifyBlock = 256, kVerifyLanes = 2, kVerifySlots = (kVerifyBlock / 32) * kVerifyLanes;
constexpr int kSmallProg = 192;          // VM capacity that covers every built-in rule (longest: key-value-credential, 183 instructions)

struct GlobalSpanSink {
  const ScanWork& w; uint32_t msg, rule;
  __device__ void span(uint32_t sb, uint32_t eb, uint32_t s16, uint32_t e16) {
    uint32_t k = atomicAdd(&w.counters[2], 1u);
    if (k < w.span_cap) { uint32_t* o = w.spans + (size_t)k * 6; o[0] = msg; o[1] = rule; o[2] = sb; o[3] = eb; o[4] = s16; o[5] = e16; }
    else atomicOr(&w.counters[3], ERR_SPAN_OVERFLOW);
  }
};

// thread-interleaved shared-memory VM storage: element i of thread t lives at base[i * blockDim + t]
struct SmemStore {
  static constexpr uint32_t cap = kSmallProg;
  uint16_t* mark_; uint16_t* pcs_; uint32_t* sts_; uint16_t* stk_; uint32_t stride, t;
  __device__ __forceinline__ uint16_t& mark(uint32_t i) { return mark_[i * stride + t]; }
  __device__ __forceinline__ uint16_t& pc(int L, uint32_t i) { return pcs_[(L * cap + i) * stride + t]; }
  __device__ __forceinline__ uint32_t& st(int L, uint32_t i) { return sts_[(L * cap + i) * stride + t]; }
  __device__ __forceinline__ uint16_t& stk(uint32_t i) { return stk_[i * stride + t]; }
};
constexpr int kStageMsg = 256;                       // messages up to this many bytes are staged into shared memory
constexpr size_t kVerifyVmBytes = (size_t)kVerifySlots * (kSmallProg * 2 + 2 * kSmallProg * 2 + 2 * kSmallProg * 4 + kVmStack * 2);
constexpr size_t kVerifyProgBytes = (size_t)kVerifySlots * (kSmallProg + 1) * 4;     // odd word stride per slot: conflict-free
constexpr size_t kVerifyMsgBytes = (size_t)kVerifySlots * (kStageMsg + 4);
constexpr size_t kVerifySmem = kVerifyVmBytes + kVerifyProgBytes + kVerifyMsgBytes;

// small programs (<= kSmallProg instructions): VM state in shared memory
template <bool SPANS>
__global__ void __launch_bounds__(kVerifyBlock)
verify_small_kernel(DevRuleset rs, ScanWork w, const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ off) {
  extern __shared__ __align__(16) uint8_t vsm[];
  const uint32_t n_events = min(w.counters[1], w.event_cap);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane >= (uint32_t)kVerifyLanes) return;
  const uint32_t slot_id = warp * kVerifyLanes + lane;                 // this thread's VM slot inside the block
  SmemStore st;
  st.stride = kVerifySlots; st.t = slot_id;
  st.sts_ = reinterpret_cast<uint32_t*>(vsm);
  st.pcs_ = reinterpret_cast<uint16_t*>(vsm + (size_t)kVerifySlots * 2 * kSmallProg * 4);
  st.mark_ = st.pcs_ + (size_t)kVerifySlots * 2 * kSmallProg;
  st.stk_ = st.mark_ + (size_t)kVerifySlots * kSmallProg;
  VMS<SmemStore> vm(rs, st);
  // per-thread staging areas: the rule's program and (for short messages) the message bytes, so the
  // VM's dependent loads hit shared memory instead of chains of L2 round trips
  uint32_t* my_prog = reinterpret_cast<uint32_t*>(vsm + kVerifyVmBytes) + (size_t)slot_id * (kSmallProg + 1);
  uint8_t* my_msg = vsm + kVerifyVmBytes + kVerifyProgBytes + (size_t)slot_id * (kStageMsg + 4);
  for (uint32_t e = blockIdx.x * kVerifySlots + slot_id; e < n_events; e += gridDim.x * kVerifySlots) {
    uint2 ev = w.events[e];
    uint32_t slot = ev.x, rule = ev.y;
    const uint32_t poff = rs.rule_prog_off[rule], plen = rs.rule_prog_off[rule + 1] - poff;
    if (plen > (uint32_t)kSmallProg) continue;          // handled by verify_large_kernel
    uint32_t msg = w.slot_msg[slot];
    GlobalSpanSink sink{w, msg, rule};
    const uint32_t t0 = w.event_pos[e];
    if (!SPANS && t0 != 0xffffffffu && ((w.hit[(size_t)slot * rs.rw + (rule >> 5)] >> (rule & 31)) & 1u)) continue;   // another occurrence already proved it
    // (staging the program / message into shared memory was measured: the copy loops cost more than the
    //  L1-cached global loads they replace, so the VM reads both through the read-only path)
    (void)my_prog; (void)my_msg; (void)poff;
    const uint8_t* m = bytes + off[msg]; const uint32_t len = off[msg + 1] - off[msg];
    bool any;
    if (SPANS || t0 == 0xffffffffu) any = run_rule<SPANS>(vm, rs, rule, m, len, sink);
    else any = test_at_factor(vm, rs, rule, m, len, t0, w.event_pre[e]);
    if (any) atomicOr(&w.hit[(size_t)slot * rs.rw + (rule >> 5)], 1u << (rule & 31));
  }
  if (vm.err) atomicOr(&w.counters[3], vm.err);
}

// programs longer than kSmallProg: per-thread local arrays
template <bool SPANS>
__global__ void __launch_bounds__(kVerifyThreads)
verify_large_kernel(DevRuleset rs, ScanWork w, const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ off) {
  const uint32_t n_events = min(w.counters[1], w.event_cap);
  VMT<kMaxProgLen> vm(rs);
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_events; e += gridDim.x * blockDim.x) {
    uint2 ev = w.events[e];
    uint32_t slot = ev.x, rule = ev.y;
    uint32_t plen = rs.rule_prog_off[rule + 1] - rs.rule_prog_off[rule];
    if (plen <= (uint32_t)kSmallProg) continue;
    uint32_t msg = w.slot_msg[slot];
    GlobalSpanSink sink{w, msg, rule};
    const uint32_t t0 = w.event_pos[e];
    bool any;
    if (SPANS || t0 == 0xffffffffu) any = run_rule<SPANS>(vm, rs, rule, bytes + off[msg], off[msg + 1] - off[msg], sink);
    else {
      if ((w.hit[(size_t)slot * rs.rw + (rule >> 5)] >> (rule & 31)) & 1u) continue;     // another occurrence already proved it
      any = test_at_factor(vm, rs, rule, bytes + off[msg], off[msg + 1] - off[msg], t0, w.event_pre[e]);
    }
    if (any) atomicOr(&w.hit[(size_t)slot * rs.rw + (rule >> 5)], 1u << (rule & 31));
  }
  if (vm.err) atomicOr(&w.counters[3], vm.err);
}

__global__ void finalize_kernel(DevRuleset rs, ScanWork w, uint64_t* __restrict__ words) {
  const uint32_t n_slots = min(w.counters[0], w.slot_cap);
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
    const uint32_t msg = w.slot_msg[s];
    if (msg == 0xffffffffu) continue;                    // slot lost an allocation race: unused
    uint32_t count = 0, first = 0xffffffffu;
    for (uint32_t k = 0; k < rs.rw; k++) {
      uint32_t v = w.hit[(size_t)s * rs.rw + k];
      if (v) { if (first == 0xffffffffu) first = k * 32 + (__ffs(v) - 1); count += __popc(v); }
    }
    if (count) words[msg] = (1ull << 63) | ((uint64_t)count << 32) | first;
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void prepare_scan_kernels() {
  int dev = 0, optin = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int kMaxSmem = (optin > 0 ? optin : 227 * 1024) - 1024;     // leave room for the kernels' static shared memory
#define CG_PREP(M) cudaFuncSetAttribute(scan_kernel<M, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); cudaFuncSetAttribute(scan_kernel<M, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)
  CG_PREP(0); CG_PREP(1); CG_PREP(2); CG_PREP(3);
#undef CG_PREP
  cudaFuncSetAttribute(scan_fp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  cudaFuncSetAttribute(verify_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVerifySmem);
  cudaFuncSetAttribute(verify_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVerifySmem);
}

int launch_scan(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off, uint32_t n,
                uint64_t* d_words, bool want_spans, int sm_count, cudaStream_t stream) {
  (void)want_spans;
  if (n == 0) return 0;
  size_t smem = rs.image_bytes;
  uint32_t ntiles = (n + 31) / 32, wpb = kScanThreads / 32;
  uint32_t grid = (ntiles + wpb - 1) / wpb; if (grid > (uint32_t)sm_count) grid = sm_count;
  if (rs.mode == 4) { scan_fp_kernel<<<std::min<uint32_t>((n + 32 * wpb - 1) / (32 * wpb), (uint32_t)sm_count), kScanThreads, smem, stream>>>(rs, w, d_bytes, d_off, n, d_words); return 1; }
  const int ns = rs.scan_streams == 1 ? 1 : 2;
#define CG_LAUNCH_SCAN(M) do { if (ns == 1) { \
      scan_kernel<M, 1><<<grid, kScanThreads, smem, stream>>>(rs, w, d_bytes, d_off, n, d_words); } \
    else { \
      scan_kernel<M, 2><<<grid, kScanThreads, smem, stream>>>(rs, w, d_bytes, d_off, n, d_words); } } while (0)
  switch (rs.mode) { case 0: CG_LAUNCH_SCAN(0); break; case 2: CG_LAUNCH_SCAN(2); break; case 3: CG_LAUNCH_SCAN(3); break; default: CG_LAUNCH_SCAN(1); break; }
#undef CG_LAUNCH_SCAN
  return 1;
}

int launch_confirm(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off,
                   bool want_spans, int sm_count, cudaStream_t stream) {
  confirm_kernel<<<sm_count * 8, 256, 0, stream>>>(rs, w, d_bytes, d_off, want_spans ? 1 : 0);
  return 1;
}

int launch_verify(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off,
                  bool want_spans, int sm_count, cudaStream_t stream) {
  int k = 1;
  if (want_spans) {
    verify_small_kernel<true><<<sm_count * 2, kVerifyBlock, kVerifySmem, stream>>>(rs, w, d_bytes, d_off);
  } else {
    verify_small_kernel<false><<<sm_count * 2, kVerifyBlock, kVerifySmem, stream>>>(rs, w, d_bytes, d_off);
  }
  if (rs.max_prog_len > (uint32_t)kSmallProg) {
    k++;
    if (want_spans) verify_large_kernel<true><<<sm_count * 4, kVerifyThreads, 0, stream>>>(rs, w, d_bytes, d_off);
    else verify_large_kernel<false><<<sm_count * 4, kVerifyThreads, 0, stream>>>(rs, w, d_bytes, d_off);
  }
  return k;
}

int launch_finalize(const DevRuleset& rs, const ScanWork& w, uint64_t* d_words, int sm_count, cudaStream_t stream) {
  finalize_kernel<<<sm_count * 2, 256, 0, stream>>>(rs, w, d_words);
  return 1;
}

}  // namespace cg
