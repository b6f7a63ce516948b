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
constexpr int kScanThreads = 1024;      // 32 warps, one CTA per SM (768 threads next to a co-resident verify CTA was measured: scan 10% slower, no overlap gain)
constexpr uint32_t kL1Always = 0xffffffffu;
constexpr uint32_t kAccept = 0x8000u, kCold = 0x4000u, kStateMask = 0x3fffu;

// pre-doubled column indices of the four bytes of a word (one byte each), SWAR
template <int MODE> __device__ __forceinline__ uint32_t cols2_of_word(uint32_t w);
template <> __device__ __forceinline__ uint32_t cols2_of_word<0>(uint32_t w) { return (w + w) & 0xfefefefeu; }
template <> __device__ __forceinline__ uint32_t cols2_of_word<2>(uint32_t w) { return ((w << 1) & 0x3e3e3e3eu) | (w & 0x40404040u); }
template <> __device__ __forceinline__ uint32_t cols2_of_word<3>(uint32_t w) { return (w << 1) & 0x3e3e3e3eu; }
template <> __device__ __forceinline__ uint32_t cols2_of_word<1>(uint32_t w) { return w; }   // LUT mode resolves per byte

// bytes between two rows of the shared-memory table: 2 * ncols + 4 (ruleset_image.cpp)
template <int MODE> struct RowStride { static constexpr uint32_t v = (MODE == 0 ? 256u : MODE == 3 ? 64u : 128u) + 4u; };

// the complete table (deep states and accept flags included), L2-resident
__device__ __forceinline__ uint32_t l1_full(const DevRuleset& rs, uint32_t state, uint32_t col) {
  return __ldg(rs.table_full + ((size_t)state << rs.ncols_log2) + col);
}

__device__ __forceinline__ void l1_push_one(const ScanWork& w, uint32_t msg, uint32_t pos, uint32_t sc) {
  uint32_t k = atomicAdd(&w.counters[4], 1u);
  if (k < w.l1_cap) { w.l1_msg[k] = msg; w.l1_pos[k] = pos; w.l1_sc[k] = sc; } else atomicOr(&w.counters[3], ERR_L1_OVERFLOW);
}

// One level-1 transition out of the shared-memory table: entries are plain state indices, so the whole
// step is  PRMT (column of byte k) ; IMAD (row * stride + column) ; LDS.U16.  Rows [0, hot) are real, row
// `hot` is an absorbing trap row that every accepting transition and every transition into a non-resident
// state leads to: the fast path never tests a flag, it only looks where it ended up after 16 bytes.
template <int MODE>
__device__ __forceinline__ uint32_t l1_step(const uint8_t* __restrict__ tbl, const uint8_t* __restrict__ lut, uint32_t stride, uint32_t state, uint32_t c2w, int k) {
  uint32_t c2 = __byte_perm(c2w, 0, 0x4440 + k);
  if (MODE == 1) c2 = 2u * lut[c2];
  return *reinterpret_cast<const uint16_t*>(tbl + state * stride + c2);
}

// Staging buffer of one warp: 64 bytes (four 16-byte units) of each of its 32 messages.  Unit u of message m
// lives at m * 64 + ((u ^ ((m >> 1) & 3)) << 4): with that XOR both the loader's STS.128 (lane -> message
// lane/4 + 8j, unit lane%4) and the walker's LDS.128 (lane -> its own message, unit cc) are bank-conflict free.
constexpr uint32_t kStageBytesPerWarp = 32u * 64u;
constexpr uint32_t kScanStageBytes = (kScanThreads / 32) * kStageBytesPerWarp;
constexpr uint32_t kEvBuf = 16;                                 // events per warp buffered in shared memory (3 words each)
constexpr uint32_t kScanEvBytes = (kScanThreads / 32) * kEvBuf * 12u;
__device__ __forceinline__ uint32_t stage_off(uint32_t m, uint32_t u) { return m * 64u + ((u ^ ((m >> 1) & 3u)) << 4); }

// Lane-per-message walk, but the message bytes do not arrive lane-per-message: a lane reading 16 bytes of its
// own message makes every LDG.128 touch 32 different 128-byte lines (32 L1 wavefronts per instruction, more
// load/store-pipe time than the table walk itself).  Instead the warp fetches "group g" = 64 bytes of each of
// its 32 messages with four LDG.128 whose lanes cover 8 messages x 64 contiguous bytes each (8 lines per
// instruction, every 32-byte sector fully used), transposes through the staging buffer, and each lane then
// pulls its own 16-byte chunks out of shared memory.  The loads of group g+1 are in flight while group g is
// walked.
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
  const uint8_t* tbl = smem;
  const uint8_t* lut = smem + rs.lut_off;
  const uint32_t hot = rs.hot_states;
  const uint32_t stride = MODE == 1 ? rs.row_stride : RowStride<MODE>::v;     // compile-time constant except in LUT mode
  const uint32_t FULL = 0xffffffffu;

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = kScanThreads / 32;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint8_t* stage = smem + ((rs.image_bytes + 127u) & ~127u) + warp * kStageBytesPerWarp;
  const uint32_t ld_unit = lane & 3u;                          // loader role: unit of the group this lane fetches
  // Level-1 accept events are collected in a small per-warp buffer in shared memory and appended to the global
  // queue kEvBuf at a time: one returning atomic (a ~1 us round trip to L2 that stalls the whole warp) per
  // flush instead of one per flagged word.
  uint32_t* evb = reinterpret_cast<uint32_t*>(smem + ((rs.image_bytes + 127u) & ~127u) + kScanStageBytes) + warp * (3u * kEvBuf);
  uint32_t evn = 0;                                            // warp-uniform fill level
  uint32_t slow_entries = 0;                                   // (same-address atomics from every slow-path entry would serialise in L2)
  auto flush_events = [&]() {
    if (evn == 0) return;
    __syncwarp();
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&w.counters[4], evn);
    base = __shfl_sync(FULL, base, 0);
    if (base + evn > w.l1_cap) { if (lane == 0) atomicOr(&w.counters[3], ERR_L1_OVERFLOW); }
    else if (lane < evn) { w.l1_msg[base + lane] = evb[3u * lane]; w.l1_pos[base + lane] = evb[3u * lane + 1]; w.l1_sc[base + lane] = evb[3u * lane + 2]; }
    evn = 0;
    __syncwarp();
  };
  // units: one per message, or (segmented scans) kSegBytes-sized pieces of long messages.  A piece other than the
  // first starts kSegWarm bytes early in state 0 and reports nothing before its own first byte: the level-1 automaton
  // is definite (its state depends on the last kMaxWindow - 1 bytes only), so by then it is in the true state.
  const bool segmented = w.units != nullptr && !(w.counters[3] & ERR_UNIT_OVERFLOW);
  const uint32_t n_units = segmented ? min(w.counters[16], w.unit_cap) : n;
  const uint32_t ntiles = (n_units + 31u) / 32u;
  for (uint32_t tile = blockIdx.x * wpb + warp; tile < ntiles; tile += gridDim.x * wpb) {
    const uint32_t unit = tile * 32u + lane;
    const bool valid = unit < n_units;
    uint32_t msg = unit, seg = 0;
    if (segmented && valid) { const uint2 un = w.units[unit]; msg = un.x; seg = un.y; }
    const uint32_t mb = valid ? off[msg] : 0u, me = valid ? off[msg + 1] : 0u;      // the message; p - mb = event position
    const uint32_t lo = seg * kSegBytes;                                            // first position this unit reports
    const uint32_t b = seg ? mb + lo - kSegWarm : mb;                               // first byte walked
    const uint32_t e = segmented ? min(me, mb + lo + kSegBytes) : me;
    if (valid && rs.n_always && lo == 0) l1_push_one(w, msg, 0, kL1Always);
    uint32_t state = 0, p = b;
    // unaligned head: byte-wise on the full table
    uint32_t head_end = (b + 15u) & ~15u; if (head_end > e) head_end = e;
    for (; p < head_end; p++) {
      uint32_t col = l1_col(rs.mode, lut, bytes[p]);
      uint32_t ent = l1_full(rs, state, col);
      if ((ent & kAccept) && p - mb >= lo) l1_push_one(w, msg, p - mb, (state << 8) | col);
      state = ent & kStateMask;
    }
    const uint32_t nch = (e - p) >> 4;
    uint32_t maxch = nch;
#pragma unroll
    for (int d = 16; d; d >>= 1) maxch = max(maxch, __shfl_xor_sync(FULL, maxch, d));
    const uint32_t ngroups = (maxch + 3u) >> 2;

    // loader role: this lane fetches unit ld_unit of messages (lane / 4) + 8 j, j = 0..3 (their chunk base and
    // count come by shuffle each time: eight registers less than keeping them)
    const uint32_t p0 = p;
    uint4 r[4];
    auto load_group = [&](uint32_t g) {
      const uint32_t cidx = 4u * g + ld_unit;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t lp = __shfl_sync(FULL, p0, (lane >> 2) + 8 * j), ln = __shfl_sync(FULL, nch, (lane >> 2) + 8 * j);
        r[j] = make_uint4(0, 0, 0, 0);
        if (cidx < ln) r[j] = ldg_stream(bytes + lp + 16u * cidx);
      }
    };
    auto store_group = [&]() {
#pragma unroll
      for (int j = 0; j < 4; j++) *reinterpret_cast<uint4*>(stage + stage_off((lane >> 2) + 8 * j, ld_unit)) = r[j];
    };
    if (ngroups) { load_group(0); __syncwarp(); store_group(); __syncwarp(); if (ngroups > 1) load_group(1); }
    for (uint32_t g = 0; g < ngroups; g++) {
#pragma unroll 1
      for (uint32_t cc = 0; cc < 4; cc++) {
        const bool act = 4u * g + cc < nch;
        const uint4 v = *reinterpret_cast<const uint4*>(stage + stage_off(lane, cc));
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint32_t c2 = cols2_of_word<MODE>(q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w);
          uint32_t fs = min(state, hot);                     // a non-resident state at word start: straight into the trap row
#pragma unroll
          for (int k = 0; k < 4; k++) fs = l1_step<MODE>(tbl, lut, stride, fs, c2, k);
          const bool flagged = act && fs == hot;
          const uint32_t fl = __ballot_sync(FULL, flagged);
          if (fl == 0) { if (act) state = fs; continue; }
          // rare: an accepting transition or an excursion into non-resident states somewhere in this word.
          // Flagged lanes re-walk the four bytes from the word's start state: resident transitions from shared
          // memory, trapped ones from the full table in L2 (accept flag + true successor).
          if (rs.debug_flags & 1u) continue;
          slow_entries++;                                    // folded into counters[6] once per warp
          uint32_t cnt = 0, ev_pos[4], ev_sc[4];
          if (flagged) {
            uint32_t st = state;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              uint32_t c2k = __byte_perm(c2, 0, 0x4440 + k);
              if (MODE == 1) c2k = 2u * lut[c2k];
              uint32_t nx = hot;
              if (st < hot) nx = *reinterpret_cast<const uint16_t*>(tbl + st * stride + c2k);
              if (nx == hot) {
                const uint32_t ent = l1_full(rs, st, c2k >> 1);
                if ((ent & kAccept) && p + 4 * q + k - mb >= lo) {
#pragma unroll
                  for (int j = 0; j <= k; j++) if (cnt == (uint32_t)j) { ev_pos[j] = p + 4 * q + k - mb; ev_sc[j] = (st << 8) | (c2k >> 1); }
                  cnt++;
                }
                nx = ent & kStateMask;
              }
              st = nx;
            }
            state = st;                                      // the true state (may be a non-resident one)
          } else if (act) state = fs;
          uint32_t idx = 0, total = 0;
#pragma unroll
          for (uint32_t j = 1; j <= 4; j++) { uint32_t bj = __ballot_sync(FULL, cnt >= j); idx += __popc(bj & lt_mask); total += __popc(bj); }
          if (total) {
            if (evn + total > kEvBuf) flush_events();
            if (total > kEvBuf) {                              // more events in one word than the buffer holds: straight to the queue
              uint32_t base = 0;
              if (lane == 0) base = atomicAdd(&w.counters[4], total);
              base = __shfl_sync(FULL, base, 0);
              if (base + total > w.l1_cap) { if (lane == 0) atomicOr(&w.counters[3], ERR_L1_OVERFLOW); }
              else {
#pragma unroll
                for (uint32_t j = 0; j < 4; j++) if (j < cnt) { uint32_t k2 = base + idx + j; w.l1_msg[k2] = msg; w.l1_pos[k2] = ev_pos[j]; w.l1_sc[k2] = ev_sc[j]; }
              }
            } else {
#pragma unroll
              for (uint32_t j = 0; j < 4; j++) if (j < cnt) { uint32_t* d = evb + 3u * (evn + idx + j); d[0] = msg; d[1] = ev_pos[j]; d[2] = ev_sc[j]; }
              evn += total;
            }
          }
        }
        if (act) p += 16;
      }
      if (g + 1 < ngroups) {
        __syncwarp(); store_group(); __syncwarp();
        if (g + 2 < ngroups) load_group(g + 2);
      }
    }
    // tail
    for (; p < e; p++) {
      uint32_t col = l1_col(rs.mode, lut, bytes[p]);
      uint32_t ent = l1_full(rs, state, col);
      if ((ent & kAccept) && p - mb >= lo) l1_push_one(w, msg, p - mb, (state << 8) | col);
      state = ent & kStateMask;
    }
    if (valid && lo == 0) words[msg] = 0ull;
  }
  flush_events();
  if (lane == 0 && slow_entries) atomicAdd(&w.counters[6], slow_entries);
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
// segmented scans: the unit table.  A message of len bytes becomes max(1, ceil(len / kSegBytes)) units; slots are
// reserved with one atomic per warp (the order of units does not matter).  counters[16] = units needed; if that
// exceeds unit_cap the scan falls back to one unit per message (ERR_UNIT_OVERFLOW) and the host grows the table.
// ------------------------------------------------------------------------------------------
__global__ void plan_units_kernel(ScanWork w, const uint32_t* __restrict__ off, uint32_t n) {
  const uint32_t FULL = 0xffffffffu, lane = threadIdx.x & 31u;
  for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t msg = i0 + threadIdx.x;
    uint32_t k = 0;
    if (msg < n) { const uint32_t len = off[msg + 1] - off[msg]; k = len <= kSegBytes ? 1u : (len + kSegBytes - 1u) / kSegBytes; }
    uint32_t incl = k;                                      // inclusive warp prefix sum
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, d); if (lane >= (uint32_t)d) incl += t; }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    uint32_t base = 0;
    if (lane == 0 && total) base = atomicAdd(&w.counters[16], total);
    base = __shfl_sync(FULL, base, 0) + incl - k;
    if (base + k > w.unit_cap) { if (k) atomicOr(&w.counters[3], ERR_UNIT_OVERFLOW); }
    else for (uint32_t j = 0; j < k; j++) w.units[base + j] = make_uint2(msg, j);
  }
}
__global__ void max_len_kernel(const uint32_t* __restrict__ off, uint32_t n, uint32_t* __restrict__ out) {
  uint32_t m = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, off[i + 1] - off[i]);
#pragma unroll
  for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31u) == 0 && m) atomicMax(out, m);
}

// ------------------------------------------------------------------------------------------
// profile-guided residency: how often is each level-1 state visited on a sample of the traffic?
// One thread per sampled message walks the full table (L2) and counts; run once per rule set (and on request),
// the host then renumbers the states so the most visited ones are the shared-memory resident ones.
// ------------------------------------------------------------------------------------------
__global__ void l1_profile_kernel(DevRuleset rs, const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ off, uint32_t n,
                                  uint32_t n_sample, uint32_t* __restrict__ visits) {
  const uint8_t* lut = rs.image + rs.lut_off;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_sample; i += gridDim.x * blockDim.x) {
    const uint32_t msg = (uint32_t)(((uint64_t)i * n) / n_sample);
    uint32_t state = 0;
    for (uint32_t p = off[msg], e = off[msg + 1]; p < e; p++) {
      atomicAdd(&visits[state], 1u);
      state = l1_full(rs, state, l1_col(rs.mode, lut, bytes[p])) & kStateMask;
    }
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
    if (e < w.event_cap) {
      // a long stretch of pattern before the factor (or an unbounded one) means many start positions for the VM: such
      // runs are an order of magnitude longer than the rest and are handed out first, so that none of them starts last
      const bool heavy = !want_spans && (pre & 0xffffu) >= 24u;
      w.events[e] = make_uint2(slot, r | (heavy ? 0x80000000u : 0u)); w.event_pos[e] = t0; w.event_pre[e] = pre;
      if (heavy) w.heavy_idx[atomicAdd(&w.counters[17], 1u)] = e;
    } else atomicOr(&w.counters[3], ERR_EVENT_OVERFLOW);
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
// verify_small_kernel: VM runs diverge completely (different program, different text), so two runs in one warp cost the
// SUM of their instruction streams: one run per warp, whose lanes cooperate inside the run instead.
constexpr int kVerifyBlock = 256, kVerifyCtasPerSm = 4, kVerifySlots = kVerifyBlock / 32;     // one VM run per warp
constexpr int kSmallProg = 192;          // VM capacity that covers every built-in rule (longest: key-value-credential, 183 instructions)

struct GlobalSpanSink {
  const ScanWork& w; uint32_t msg, rule; bool lane0_only;
  __device__ void span(uint32_t sb, uint32_t eb, uint32_t s16, uint32_t e16) {
    if (lane0_only && (threadIdx.x & 31u)) return;         // warp-redundant runs (verify_small_kernel): one lane reports
    uint32_t k = atomicAdd(&w.counters[2], 1u);
    if (k < w.span_cap) { uint32_t* o = w.spans + (size_t)k * 6; o[0] = msg; o[1] = rule; o[2] = sb; o[3] = eb; o[4] = s16; o[5] = e16; }
    else atomicOr(&w.counters[3], ERR_SPAN_OVERFLOW);
  }
};

// thread-interleaved shared-memory VM storage: element i of thread t lives at base[i * blockDim + t]
struct SmemStore {
  static constexpr uint32_t cap = kSmallProg;
  static constexpr bool coop = true;       // one run per WARP: every lane executes it redundantly and helps where it can (pike_vm.h)
  uint16_t* mark_; uint16_t* pcs_; uint32_t* sts_; uint16_t* stk_; uint32_t stride, t;
  __device__ __forceinline__ uint16_t& mark(uint32_t i) { return mark_[i * stride + t]; }
  __device__ __forceinline__ uint16_t& pc(int L, uint32_t i) { return pcs_[(L * cap + i) * stride + t]; }
  __device__ __forceinline__ uint32_t& st(int L, uint32_t i) { return sts_[(L * cap + i) * stride + t]; }
  __device__ __forceinline__ uint16_t& stk(uint32_t i) { return stk_[i * stride + t]; }
};
constexpr int kStageMsg = 256;                       // messages up to this many bytes are staged into shared memory
constexpr size_t kVerifyVmBytes = (size_t)kVerifySlots * (kSmallProg * 2 + 2 * kSmallProg * 2 + 2 * kSmallProg * 4 + kVmStack * 2);
constexpr size_t kVerifyProgBytes = 0;      // (staging areas: measured, not worth their copy loops)
constexpr size_t kVerifyMsgBytes = 0;
constexpr size_t kVerifySmem = kVerifyVmBytes + kVerifyProgBytes + kVerifyMsgBytes;

// small programs (<= kSmallProg instructions): VM state in shared memory
template <bool SPANS>
__global__ void __launch_bounds__(kVerifyBlock, kVerifyCtasPerSm)
verify_small_kernel(DevRuleset rs, ScanWork w, const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ off) {
  extern __shared__ __align__(16) uint8_t vsm[];
  const uint32_t n_events = min(w.counters[1], w.event_cap);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // One run per warp.  All 32 lanes execute it redundantly -- identical registers, uniform control flow, the VM
  // state in shared memory written with identical values -- so that the lanes can split the work wherever a step
  // is data-parallel (SIMD literal / class stretches, mark clearing; see pike_vm.h).  Only lane 0 reports.
  const uint32_t slot_id = warp;
  SmemStore st;
  st.stride = kVerifySlots; st.t = slot_id;
  st.sts_ = reinterpret_cast<uint32_t*>(vsm);
  st.pcs_ = reinterpret_cast<uint16_t*>(vsm + (size_t)kVerifySlots * 2 * kSmallProg * 4);
  st.mark_ = st.pcs_ + (size_t)kVerifySlots * 2 * kSmallProg;
  st.stk_ = st.mark_ + (size_t)kVerifySlots * kSmallProg;
  VMS<SmemStore> vm(rs, st);
  // events are handed out one at a time (their cost varies by two orders of magnitude): a static split leaves most
  // warps idle while a few finish their second or third long run
  const uint32_t n_heavy = min(w.counters[17], n_events);
  bool heavy_phase = true;
  for (;;) {
    uint32_t e = 0;
    if (lane == 0) e = atomicAdd(&w.counters[heavy_phase ? 5 : 18], 1u);
    e = __shfl_sync(0xffffffffu, e, 0);
    if (heavy_phase) {
      if (e >= n_heavy) { heavy_phase = false; continue; }
      e = w.heavy_idx[e];
    } else if (e >= n_events) break;
    uint2 ev = w.events[e];
    if (!heavy_phase && (ev.y >> 31)) continue;             // done (or being done) by the heavy phase
    uint32_t slot = ev.x, rule = ev.y & 0x7fffffffu;
    const uint32_t plen = rs.rule_prog_off[rule + 1] - rs.rule_prog_off[rule];
    if (plen > (uint32_t)kSmallProg) continue;          // handled by verify_large_kernel
    uint32_t msg = w.slot_msg[slot];
    GlobalSpanSink sink{w, msg, rule, true};
    const uint32_t t0 = w.event_pos[e];
    if (!SPANS && t0 != 0xffffffffu) {                  // another occurrence already proved it?  (one read, broadcast: the warp must agree)
      uint32_t hv = 0;
      if (lane == 0) hv = w.hit[(size_t)slot * rs.rw + (rule >> 5)];
      hv = __shfl_sync(0xffffffffu, hv, 0);
      if ((hv >> (rule & 31)) & 1u) continue;
    }
    const uint8_t* m = bytes + off[msg]; const uint32_t len = off[msg + 1] - off[msg];
    bool any;
    const long long tc0 = (rs.debug_flags & 2u) ? clock64() : 0;
    if (SPANS || t0 == 0xffffffffu) any = run_rule<SPANS>(vm, rs, rule, m, len, sink);
    else any = test_at_factor(vm, rs, rule, m, len, t0, w.event_pre[e]);
    __syncwarp();
    if (any && lane == 0) atomicOr(&w.hit[(size_t)slot * rs.rw + (rule >> 5)], 1u << (rule & 31));
    if ((rs.debug_flags & 2u) && lane == 0) {       // CG_SCAN_DEBUG=2: histogram of VM cycles per event in counters[8..15] (<2^11, <2^12, ... >=2^17), max in [7]
      const uint32_t cyc = (uint32_t)(clock64() - tc0);
      int bkt = 31 - __clz(cyc | 1u) - 10; bkt = bkt < 0 ? 0 : bkt > 7 ? 7 : bkt;
      atomicAdd(&w.counters[8 + bkt], 1u); atomicMax(&w.counters[7], cyc);
    }
    __syncwarp();                                          // lane 0 reported; everyone together again before the next run touches shared memory
  }
  if (vm.err && lane == 0) atomicOr(&w.counters[3], vm.err);
}

// programs longer than kSmallProg: per-thread local arrays
template <bool SPANS>
__global__ void __launch_bounds__(kVerifyThreads)
verify_large_kernel(DevRuleset rs, ScanWork w, const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ off) {
  const uint32_t n_events = min(w.counters[1], w.event_cap);
  VMT<kMaxProgLen> vm(rs);
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_events; e += gridDim.x * blockDim.x) {
    uint2 ev = w.events[e];
    uint32_t slot = ev.x, rule = ev.y & 0x7fffffffu;
    uint32_t plen = rs.rule_prog_off[rule + 1] - rs.rule_prog_off[rule];
    if (plen <= (uint32_t)kSmallProg) continue;
    uint32_t msg = w.slot_msg[slot];
    GlobalSpanSink sink{w, msg, rule, false};
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

// Verdict of a message from its hit bitmap (aggregateMatches, policy-evaluator.ts:110-146, for messageContains rules):
// policies in priority order, per policy the FIRST matching rule counts, deny beats audit beats allow, and the first
// policy that produced the winning action is the one whose reason is reported.  Rules are stored policy by policy, so an
// ascending walk over the hit bits meets every policy's deciding rule first.
__global__ void verdict_kernel(DevRuleset rs, ScanWork w, uint32_t* __restrict__ verdicts) {
  const uint32_t n_slots = min(w.counters[0], w.slot_cap);
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
    const uint32_t msg = w.slot_msg[s];
    if (msg == 0xffffffffu) continue;
    uint32_t prev = 0xffffffffu, best = 0, decide = 0xfffffu, matched = 0;
    for (uint32_t k = 0; k < rs.rw; k++) {
      uint32_t v = w.hit[(size_t)s * rs.rw + k];
      while (v) {
        const uint32_t r = k * 32 + (__ffs(v) - 1); v &= v - 1;
        const uint32_t p = rs.rule_policy[r];
        if (p == prev) continue;
        prev = p; matched++;
        const uint32_t a = rs.rule_action[r];
        if (a == 2u && best != 2u) { best = 2u; decide = r; } else if (a == 1u && best == 0u) { best = 1u; decide = r; }
      }
    }
    if (matched) verdicts[msg] = best | (min(matched, 1023u) << 2) | (decide << 12);
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void prepare_scan_kernels() {
  int dev = 0, optin = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int kMaxSmem = (optin > 0 ? optin : 227 * 1024) - 1024;     // leave room for the kernels' static shared memory
#define CG_PREP(M) cudaFuncSetAttribute(scan_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem)
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
  uint32_t grid = (ntiles + wpb - 1) / wpb; if (grid > (uint32_t)sm_count || w.units) grid = sm_count;      // segmented: unit count is only known on the device
  if (rs.mode == 4) { scan_fp_kernel<<<std::min<uint32_t>((n + 32 * wpb - 1) / (32 * wpb), (uint32_t)sm_count), kScanThreads, smem, stream>>>(rs, w, d_bytes, d_off, n, d_words); return 1; }
  smem = ((smem + 127) & ~(size_t)127) + kScanStageBytes + kScanEvBytes;      // image + per warp: 2 KB staging buffer, event buffer
#define CG_LAUNCH_SCAN(M) scan_kernel<M><<<grid, kScanThreads, smem, stream>>>(rs, w, d_bytes, d_off, n, d_words)
  switch (rs.mode) { case 0: CG_LAUNCH_SCAN(0); break; case 2: CG_LAUNCH_SCAN(2); break; case 3: CG_LAUNCH_SCAN(3); break; default: CG_LAUNCH_SCAN(1); break; }
#undef CG_LAUNCH_SCAN
  return 1;
}

int launch_plan_units(const ScanWork& w, const uint32_t* d_off, uint32_t n, cudaStream_t stream) {
  if (!n || !w.units) return 0;
  plan_units_kernel<<<std::min<uint32_t>((n + 255) / 256, 1184u), 256, 0, stream>>>(w, d_off, n);
  return 1;
}
int launch_max_len(const uint32_t* d_off, uint32_t n, uint32_t* d_out_max, cudaStream_t stream) {
  if (!n) return 0;
  max_len_kernel<<<std::min<uint32_t>((n + 255) / 256, 1184u), 256, 0, stream>>>(d_off, n, d_out_max);
  return 1;
}

int launch_l1_profile(const DevRuleset& rs, const uint8_t* d_bytes, const uint32_t* d_off, uint32_t n, uint32_t n_sample,
                      uint32_t* d_visits, cudaStream_t stream) {
  if (!n || !n_sample) return 0;
  l1_profile_kernel<<<(n_sample + 127) / 128, 128, 0, stream>>>(rs, d_bytes, d_off, n, n_sample, d_visits);
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
    verify_small_kernel<true><<<sm_count * kVerifyCtasPerSm, kVerifyBlock, kVerifySmem, stream>>>(rs, w, d_bytes, d_off);
  } else {
    verify_small_kernel<false><<<sm_count * kVerifyCtasPerSm, kVerifyBlock, kVerifySmem, stream>>>(rs, w, d_bytes, d_off);
  }
  if (rs.max_prog_len > (uint32_t)kSmallProg) {
    k++;
    if (want_spans) verify_large_kernel<true><<<sm_count * 4, kVerifyThreads, 0, stream>>>(rs, w, d_bytes, d_off);
    else verify_large_kernel<false><<<sm_count * 4, kVerifyThreads, 0, stream>>>(rs, w, d_bytes, d_off);
  }
  return k;
}

int launch_verdicts(const DevRuleset& rs, const ScanWork& w, uint32_t* d_verdicts, int sm_count, cudaStream_t stream) {
  verdict_kernel<<<sm_count * 2, 256, 0, stream>>>(rs, w, d_verdicts);
  return 1;
}

int launch_finalize(const DevRuleset& rs, const ScanWork& w, uint64_t* d_words, int sm_count, cudaStream_t stream) {
  finalize_kernel<<<sm_count * 2, 256, 0, stream>>>(rs, w, d_words);
  return 1;
}

}  // namespace cg
