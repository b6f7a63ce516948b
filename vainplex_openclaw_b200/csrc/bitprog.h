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

// table of one rule (uint64 words): accept[128] | start[8] | header | follow rows [n_rows][64]
//   header = row of context c in bits 4c..4c+3, n_rows in bits 32..35 (contexts with identical follow rows share one: a rule
//   without assertions has a single row)
constexpr uint32_t kBitAccept = 0, kBitStart = 128, kBitHeader = 136, kBitRows = 137;
constexpr uint32_t kBitProgWords = kBitRows + 8 * 64;
constexpr uint32_t kBitProgNone = 0xffffffffu;
constexpr uint64_t kBitMatch = 1ull << 63;

// context of the position between the bytes `prev` and `next` (-1 = none): bit 0 word boundary, bit 1 start of message, bit 2 end
CG_HD uint32_t bitprog_ctx(int prev, int next, bool at_start) {
  return (uint32_t)(is_word(prev) != is_word(next)) | (at_start ? 2u : 0u) | (next < 0 ? 4u : 0u);
}

// -> 1 a match that starts in [s, t0] exists, 0 none, -1 cannot tell (non-ASCII byte in the island, or more than max_steps
// bytes to walk: ask the VM, whose warp-wide runs suit long islands better than one thread's chain of table loads)
CG_HD int bitprog_test(const uint64_t* __restrict__ bp, const uint8_t* __restrict__ m, uint32_t len, uint32_t s, uint32_t t0, uint32_t max_steps = 0xffffffffu) {
  const uint64_t* accept = bp + kBitAccept; const uint64_t* start = bp + kBitStart; const uint64_t* rows = bp + kBitRows; const uint64_t header = bp[kBitHeader];
  int prev = s > 0 ? (m[s - 1] < 0x80 ? (int)m[s - 1] : 0x80) : -1;          // (a unit >= 0x80 is not a word character, whatever it is)
  int cur = s < len ? (int)m[s] : -1;
  uint32_t ctx = bitprog_ctx(prev, cur, s == 0);
  uint64_t live = start[ctx];
  uint64_t acc_cur = cur >= 0 && cur < 0x80 ? accept[cur] : 0;               // (the accept mask is fetched one byte ahead of the state it meets)
  for (uint32_t pos = s;; ) {
    if (live & kBitMatch) return 1;
    if (cur < 0) return 0;
    if (cur >= 0x80 || pos - s >= max_steps) return -1;
    uint64_t hit = live & acc_cur;
    const int nxt = pos + 1 < len ? (int)m[pos + 1] : -1;
    acc_cur = nxt >= 0 && nxt < 0x80 ? accept[nxt] : 0;
    ctx = bitprog_ctx(cur, nxt < 0x80 ? nxt : 0x80, false);
    const uint64_t* fw = rows + ((header >> (4 * ctx)) & 15u) * 64;
    live = 0;
    while (hit) {
#if defined(__CUDA_ARCH__)
      const int k = __ffsll((long long)hit) - 1;
#else
      const int k = __builtin_ctzll(hit);
#endif
      hit &= hit - 1; live |= fw[k];
    }
    pos++;
    if (pos <= t0) live |= start[ctx];                                     // a new start position (none beyond the factor occurrence)
    else if (!live) return 0;
    prev = cur; cur = nxt;
  }
}

// the island's first byte: test_at_factor's start position (pike_vm.h), without the cursor (ASCII islands only matter here)
CG_HD uint32_t island_start(const DevRuleset& rs, uint32_t rule, const uint8_t* __restrict__ m, uint32_t len, uint32_t t0, uint32_t pre_units) {
  const uint32_t pa = pre_units >> 16; pre_units &= 0xffffu;
  const uint32_t* alpha = pa != 0xffffu ? rs.bytesets + (size_t)pa * 8 : rs.rule_alpha + (size_t)rule * 8;
  if (t0 > len) t0 = len;
  uint32_t s = t0;
  const uint32_t floor_ = pre_units < 0xffffu ? (t0 > 3u * pre_units ? t0 - 3u * pre_units : 0u) : 0u;     // (earlier starts cannot reach the factor)
  while (s > floor_ && ((alpha[m[s - 1] >> 5] >> (m[s - 1] & 31)) & 1u)) s--;
  return s;
}

}  // namespace cg
