// bitprog.h -- RegExp.test around a confirmed factor occurrence for the rules whose Pike program has at most 63 consuming
// instructions and no lookaround: a bit-parallel simulation of the same program (one bit per consuming instruction, bit 63
// = MATCH), one THREAD per occurrence instead of one warp per Pike-VM run.
//
// What makes the boolean answer of gov/src/conditions/context.ts:9-25 (RegExp.test per rule) independent of thread
// priorities: a match exists iff MATCH is reachable from some start position, whatever the order alternatives are tried
// in; the empty-iteration check of ECMA-262 22.2.2.3.1 step 2.b only prunes iterations that consume nothing, which
// never changes the set of end positions reachable (no backreferences, no lookaround in these programs).  So the
// program's epsilon edges are closed once, at compile time, per context of the position between two units -- the
// assertions ^ $ \b \B only look at (word boundary?, start of message?, end of message?) -- and a step is
// "accepting instructions for this byte" & "live instructions", then OR of the follow masks of what is left.
//
// ASCII only: a byte >= 0x80 inside the island hands the occurrence to the Pike VM (UTF-16 unit semantics, pike_vm.h).
// The tables are built from the rule's own Pike program (rulec.cpp output), so both matchers interpret ONE compilation.
#pragma once
#include <cstdint>
#include "kernels.h"
#include "pike_vm.h"

namespace cg {

// table of one rule (uint64 words): accept[128] | start[8] | header | next | self | follow rows [n_rows][64]
//   header = row of context c in bits 4c..4c+3, n_rows in bits 32..35 (contexts with identical follow rows share one: a rule
//   without assertions has a single row), bit 40 = start[] is the same in all eight contexts
//   next / self = the instructions k whose follow set is, in every row, a subset of {k, k + 1}: bit k of `next` when it holds
//   k + 1 (the next unit of a literal or class run; bit 62 -> MATCH), bit k of `self` when it holds k (a loop on one unit).
//   Their successors are a shift and a mask -- no load; only the other instructions (alternations, group loops) go
//   through the follow rows.  One thread walks an island, so every load on this path is a round trip to L2.
constexpr uint32_t kBitAccept = 0, kBitStart = 128, kBitHeader = 136, kBitNext = 137, kBitSelf = 138, kBitRows = 139;
constexpr uint64_t kBitStartSame = 1ull << 40;
constexpr uint32_t kBitProgWords = kBitRows + 8 * 64;
constexpr uint32_t kBitProgNone = 0xffffffffu;
constexpr uint64_t kBitMatch = 1ull << 63;

// context of the position between the bytes `prev` and `next` (-1 = none): bit 0 word boundary, bit 1 start of message, bit 2 end
CG_HD uint32_t bitprog_ctx(int prev, int next, bool at_start) {
  return (uint32_t)(is_word(prev) != is_word(next)) | (at_start ? 2u : 0u) | (next < 0 ? 4u : 0u);
}

// -> 1 a match that starts in [s, t0] exists, 0 none, -1 cannot tell (non-ASCII byte in the island, or more than max_steps
// bytes to walk: ask the VM, whose warp-wide runs suit long islands better than one thread's chain of table loads).
// One thread walks the island, so what this costs is the length of its dependency chain, not its instruction count: the
// four next bytes and their accept masks are fetched together (loads that depend on the position alone), and only the
// state update is sequential -- a shift and a mask for most instructions, a follow row from L2 for the others.
CG_HD int bitprog_test(const uint64_t* __restrict__ bp, const uint8_t* __restrict__ m, uint32_t len, uint32_t s, uint32_t t0, uint32_t max_steps = 0xffffffffu) {
  const uint64_t* accept = bp + kBitAccept; const uint64_t* start = bp + kBitStart; const uint64_t* rows = bp + kBitRows; const uint64_t header = bp[kBitHeader];
  const uint64_t m_next = bp[kBitNext], m_self = bp[kBitSelf], m_table = ~(m_next | m_self);
  const bool start_same = (header & kBitStartSame) != 0, one_row = ((header >> 32) & 15u) == 1u;
  int prev = s > 0 ? (m[s - 1] < 0x80 ? (int)m[s - 1] : 0x80) : -1;          // (a unit >= 0x80 is not a word character, whatever it is)
  int cur = s < len ? (int)m[s] : -1;
  uint32_t ctx = bitprog_ctx(prev, cur, s == 0);
  const uint64_t start0 = start[ctx];
  uint64_t live = start0;
  uint64_t acc_cur = cur >= 0 && cur < 0x80 ? accept[cur] : 0;
  for (uint32_t pos = s;;) {
    // the block's four "next" bytes and their accept masks
    int nb[4]; uint64_t na[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) nb[i] = pos + 1 + i < len ? (int)m[pos + 1 + i] : -1;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) na[i] = nb[i] >= 0 && nb[i] < 0x80 ? accept[nb[i]] : 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) {
      if (live & kBitMatch) return 1;
      if (cur < 0) return 0;
      if (cur >= 0x80 || pos - s >= max_steps) return -1;
      const uint64_t hit_all = live & acc_cur;
      const int nxt = nb[i];
      live = ((hit_all & m_next) << 1) | (hit_all & m_self);
      uint64_t hit = hit_all & m_table;
      if (!start_same || hit) ctx = bitprog_ctx(cur, nxt < 0x80 ? nxt : 0x80, false);
      if (hit) {
        const uint64_t* fw = one_row ? rows : rows + ((header >> (4 * ctx)) & 15u) * 64;
        while (hit) {
#if defined(__CUDA_ARCH__)
          const int k = __ffsll((long long)hit) - 1;
#else
          const int k = __builtin_ctzll(hit);
#endif
          hit &= hit - 1; live |= fw[k];
        }
      }
      pos++;
      if (pos <= t0) live |= start_same ? start0 : start[ctx];             // a new start position (none beyond the factor occurrence)
      else if (!live) return 0;
      cur = nxt; acc_cur = na[i];
    }
  }
}

// one word of a 256-bit set held in eight registers
CG_HD uint32_t bitprog_sel8(const uint32_t a[8], uint32_t i) {
  const uint32_t lo = (i & 2u) ? ((i & 1u) ? a[3] : a[2]) : ((i & 1u) ? a[1] : a[0]);
  const uint32_t hi = (i & 2u) ? ((i & 1u) ? a[7] : a[6]) : ((i & 1u) ? a[5] : a[4]);
  return (i & 4u) ? hi : lo;
}

// the island's first byte: test_at_factor's start position (pike_vm.h), without the cursor (ASCII islands only matter here).
// Four bytes per round, the set in registers: the chain is the tests, not a load per byte.  Gives up `limit` bytes before t0
// (the caller then hands the occurrence to the VM: bitprog_test would not walk that far either).
CG_HD uint32_t island_start(const DevRuleset& rs, uint32_t rule, const uint8_t* __restrict__ m, uint32_t len, uint32_t t0, uint32_t pre_units, uint32_t limit = 0xffffffffu) {
  const uint32_t pa = pre_units >> 16; pre_units &= 0xffffu;
  const uint32_t* alpha = pa != 0xffffu ? rs.bytesets + (size_t)pa * 8 : rs.rule_alpha + (size_t)rule * 8;
  if (t0 > len) t0 = len;
  uint32_t s = t0;
  uint32_t floor_ = pre_units < 0xffffu ? (t0 > 3u * pre_units ? t0 - 3u * pre_units : 0u) : 0u;     // (earlier starts cannot reach the factor)
  if (limit != 0xffffffffu && t0 - floor_ > limit) floor_ = t0 - limit - (t0 > limit ? 1u : 0u);      // (one byte past the limit tells "too long" from "ends there")
  if (s <= floor_) return s;
  uint32_t a[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < 8; i++) a[i] = alpha[i];
  for (;;) {
    uint32_t b[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (uint32_t i = 0; i < 4; i++) b[i] = s > floor_ + i ? m[s - 1 - i] : 0x100u;        // 0x100: stop
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (uint32_t i = 0; i < 4; i++) {
      if (b[i] > 0xffu || !((bitprog_sel8(a, b[i] >> 5) >> (b[i] & 31u)) & 1u)) return s;
      s--;
    }
  }
}

}  // namespace cg
