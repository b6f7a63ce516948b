// pike_vm.h -- the exact matcher: a Pike VM over UTF-16 code units decoded on the fly from UTF-8.
//
// ECMAScript leftmost-first semantics (ECMA-262 22.2.2) for the supported subset: priority-ordered
// thread list, closure with per-step visited marks (which is exactly the spec's empty-check for
// `*` loops), MATCH cuts every lower-priority thread.  The global-exec iteration of
// gov/src/redaction/registry.ts:225-236 and RegExp.test of gov/src/conditions/context.ts:9-25 are
// built on vm.search().  The functions are CG_HD so that tests/native/vm_harness.cpp can run the
// *same code* on the host against the oracle; the product only ever calls them from verify_kernel.
#pragma once
#include <cstdint>
#include "kernels.h"
#include "rulec.h"

#ifdef __CUDACC__
#define CG_HD __host__ __device__ __forceinline__
#define CG_HD_NOINLINE __host__ __device__
#else
#define CG_HD inline
#define CG_HD_NOINLINE inline
#endif

#if defined(CG_VM_STATS) && !defined(__CUDA_ARCH__)
#define CG_VM_STAT(i) (++cg_vm_stats[i])
extern "C" unsigned long long cg_vm_stats[8];     // [0] closure (add) calls, [1] closure instructions, [2] list-thread steps, [3] search steps, [4] stretch units
#else
#define CG_VM_STAT(i) ((void)0)
#endif

namespace cg {

struct Cursor { uint32_t pos; uint32_t pending; };   // pending = low surrogate still to deliver (0 = none)

// WHATWG UTF-8 decode of one UTF-16 unit at the cursor; returns -1 at end of message.
CG_HD int next_unit(const uint8_t* __restrict__ m, uint32_t len, Cursor& c) {
  if (c.pending) { int u = (int)c.pending; c.pending = 0; c.pos += 2; return u; }
  if (c.pos >= len) return -1;
  uint32_t b0 = m[c.pos];
  if (b0 < 0x80) { c.pos += 1; return (int)b0; }
  uint32_t need, cp, lo = 0x80, hi = 0xbf;
  if (b0 >= 0xc2 && b0 <= 0xdf) { need = 1; cp = b0 & 0x1f; }
  else if (b0 >= 0xe0 && b0 <= 0xef) { need = 2; cp = b0 & 0x0f; if (b0 == 0xe0) lo = 0xa0; if (b0 == 0xed) hi = 0x9f; }
  else if (b0 >= 0xf0 && b0 <= 0xf4) { need = 3; cp = b0 & 0x07; if (b0 == 0xf0) lo = 0x90; if (b0 == 0xf4) hi = 0x8f; }
  else { c.pos += 1; return 0xfffd; }
  uint32_t j = c.pos + 1;
  for (uint32_t t = 0; t < need; t++, j++) {
    if (j >= len) { c.pos = j; return 0xfffd; }
    uint32_t bj = m[j];
    if (bj < lo || bj > hi) { c.pos = (j > c.pos + 1) ? j : c.pos + 1; return 0xfffd; }
    cp = (cp << 6) | (bj & 0x3f); lo = 0x80; hi = 0xbf;
  }
  if (cp >= 0x10000) {   // astral: deliver the high surrogate now (2 bytes), the low one next
    cp -= 0x10000; c.pending = 0xdc00 + (cp & 0x3ff); c.pos += 2; return (int)(0xd800 + (cp >> 10));
  }
  c.pos = j; return (int)cp;
}

CG_HD bool in_set(const DevRuleset& rs, uint32_t sid, int u) {
  if (u < 0) return false;
  const uint32_t* s = rs.sets + (size_t)sid * 6;
  if (u < 128) return (s[u >> 5] >> (u & 31)) & 1u;
  const uint16_t* r = rs.set_ranges + s[4];
  for (uint32_t k = 0; k < s[5]; k++) if ((uint32_t)u >= r[2 * k] && (uint32_t)u <= r[2 * k + 1]) return true;
  return false;
}
CG_HD bool is_word(int u) { return (u >= '0' && u <= '9') || (u >= 'A' && u <= 'Z') || u == '_' || (u >= 'a' && u <= 'z'); }

constexpr int kVmStack = kVmStackLimit;

// VM working storage.  LocalStore: plain per-thread arrays (host harness, large programs).
// SmemStore (scan_kernels.cu): thread-interleaved shared memory for the common small programs.
template <int CAP>
struct LocalStore {
  static constexpr uint32_t cap = CAP;
  static constexpr bool coop = false;      // one run per thread
  uint16_t mark_[CAP]; uint16_t pcs_[2][CAP]; uint32_t sts_[2][CAP]; uint16_t stk_[kVmStack];
  CG_HD uint16_t& mark(uint32_t i) { return mark_[i]; }
  CG_HD uint16_t& pc(int L, uint32_t i) { return pcs_[L][i]; }
  CG_HD uint32_t& st(int L, uint32_t i) { return sts_[L][i]; }
  CG_HD uint16_t& stk(uint32_t i) { return stk_[i]; }
};

template <class Store>
struct VMS {
  const DevRuleset& rs; const uint32_t* prog; uint32_t plen;
  const uint32_t* prog_override = nullptr;    // the caller's staged copy of the current rule's program (shared memory), if any
  Store S; uint16_t gen; uint32_t cnt[2];
  uint32_t err;
  // best match of the current search
  bool matched; uint32_t m_start, m_end, m_end16; int m_prev;

  static CG_HD uint32_t capacity() { return Store::cap; }
  static constexpr bool coop = Store::coop;
  // warp-redundant runs: after any section only some lanes execute (lane 0 reporting a span, ...) the warp must be back
  // together before the next shared-memory access -- every lane reads what every lane wrote
  static CG_HD void rejoin() {
#if defined(__CUDA_ARCH__)
    if (Store::coop) __syncwarp();
#endif
  }
  CG_HD_NOINLINE VMS(const DevRuleset& r) : rs(r), gen(0), err(0) {}
  CG_HD_NOINLINE VMS(const DevRuleset& r, const Store& s) : rs(r), S(s), gen(0), err(0) {}

  // all marks to zero, generation restarted (cooperative stores: the warp's lanes split the array)
  CG_HD void clear_marks() {
#if defined(__CUDA_ARCH__)
    if (Store::coop) { for (uint32_t k = threadIdx.x & 31u; k < plen; k += 32u) S.mark(k) = 0; __syncwarp(); gen = 0; return; }
#endif
    for (uint32_t k = 0; k < plen; k++) S.mark(k) = 0;
    gen = 0;
  }
  CG_HD_NOINLINE void bump_gen() { if (++gen >= 0x7fff) { for (uint32_t k = 0; k < plen; k++) S.mark(k) = 0; gen = 1; } }

  // Closure from pc0 at a position whose context is (prev, next, byte `pos`, unit index `upos`):
  // a depth-first walk in priority order.  SPLIT nodes carry a per-step mark that is IN-PROGRESS
  // while their lower-priority alternative is still pending and DONE afterwards; consuming
  // instructions are marked DONE when their thread is queued.  Arriving at a DONE node is a
  // duplicate of a higher-priority path and is dropped.  A loop back-edge (JMP_BACK / EMPTYCHK)
  // that finds its own SPLIT in progress at this very position closes an iteration that consumed
  // nothing -- RepeatMatcher's empty check (ECMA-262 22.2.2.3.1 step 2.b) -- and dies, while a
  // *fresh* entry into an in-progress SPLIT (an inner loop re-entered by a new outer iteration) is
  // explored again at its higher priority.  Returns true when MATCH was reached (the caller must
  // then cut every lower-priority thread).
  CG_HD bool add(int L, uint32_t pc0, uint32_t start, int prev, int next, uint32_t pos, uint32_t upos) {
    const uint16_t INPROG = (uint16_t)(gen * 2), DONE = (uint16_t)(gen * 2 + 1);
    CG_VM_STAT(0);
    int sp = 0; S.stk(sp++) = (uint16_t)pc0;
    while (sp) {
      uint32_t x = S.stk(--sp);
      if (x & 0x8000u) { S.mark(x & 0x7fffu) = DONE; continue; }
      uint32_t pc = x;
      for (;;) {
        uint32_t ins = prog[pc], op = ins & 0xff, arg = ins >> 8;
        bool go = false; CG_VM_STAT(1);
        switch (op) {
          case OP_SPLIT_NEXT: case OP_SPLIT_JUMP:
            if (S.mark(pc) == DONE) break;
            S.mark(pc) = INPROG;
            if (sp + 2 <= kVmStack) { S.stk(sp++) = (uint16_t)(0x8000u | pc); S.stk(sp++) = (uint16_t)(op == OP_SPLIT_NEXT ? arg : pc + 1); }
            else err |= ERR_VM_STACK;
            pc = op == OP_SPLIT_NEXT ? pc + 1 : arg; go = true; break;
          case OP_JMP: pc = arg; go = true; break;
          case OP_JMP_BACK: if (S.mark(arg) != INPROG) { pc = arg; go = true; } break;
          case OP_EMPTYCHK: if (S.mark(arg) != INPROG) { pc++; go = true; } break;
          case OP_BOL: if (pos == 0) { pc++; go = true; } break;
          case OP_EOL: if (next < 0) { pc++; go = true; } break;
          case OP_WORDB: if (is_word(prev) != is_word(next)) { pc++; go = true; } break;
          case OP_NWORDB: if (is_word(prev) == is_word(next)) { pc++; go = true; } break;
          case OP_LOOKAHEAD: if (in_set(rs, arg, next)) { pc++; go = true; } break;
          case OP_NLOOKAHEAD: if (!in_set(rs, arg, next)) { pc++; go = true; } break;
          case OP_LOOKBEHIND: if (in_set(rs, arg, prev)) { pc++; go = true; } break;
          case OP_NLOOKBEHIND: if (!in_set(rs, arg, prev)) { pc++; go = true; } break;
          case OP_MATCH:
            matched = true; m_start = start; m_end = pos; m_end16 = upos; m_prev = prev;
            return true;
          default:   // consuming instruction: queue the thread once per step
            if (S.mark(pc) != DONE) {
              S.mark(pc) = DONE;
              uint32_t k = cnt[L];
              if (k < Store::cap) { S.pc(L, k) = (uint16_t)pc; S.st(L, k) = start; cnt[L] = k + 1; } else err |= ERR_VM_LIST;
            }
            break;
        }
        if (!go) break;
      }
    }
    return false;
  }

  // One leftmost-first search starting at byte `from` (unit index from16, unit before it `prev`).
  CG_HD_NOINLINE bool search(const uint8_t* __restrict__ m, uint32_t len, uint32_t from, uint32_t from16, int prev,
                         const uint32_t* __restrict__ first, Cursor cur_c, uint32_t start_limit = 0xffffffffu, bool first_match_only = false) {
    matched = false;
    Cursor c = cur_c;                 // cursor sits at `from`
    uint32_t pos = from, upos = from16;
    Cursor cn = c; int cur = next_unit(m, len, cn);      // unit at pos ; cn = cursor after it
    int L = 0; cnt[0] = 0; bump_gen();
    add(L, 0, pos, prev, cur, pos, upos);
    for (;;) {
      if (matched && first_match_only) break;
      if (cnt[L] == 0) {
        if (matched || cur < 0) break;
        // Nothing alive and the start at `pos` has already failed: move on, skipping ASCII units
        // that cannot begin a match (first-unit filter; nullable rules have an all-ones filter).
        do { prev = cur; pos = cn.pos; upos++; c = cn; cur = next_unit(m, len, cn); }
        while (pos <= start_limit && cur >= 0 && cur < 128 && !((first[cur >> 5] >> (cur & 31)) & 1u));
        if (pos > start_limit) break;
        bump_gen();
        add(L, 0, pos, prev, cur, pos, upos);
        continue;
      }
      if (cur < 0) break;
      // Deterministic stretch: one live thread, no further start positions, and straight-line consuming
      // instructions ahead (a literal, an expanded x{36}): walk it without the list / closure machinery.  The
      // last unit of the stretch -- the one whose successor instruction needs a closure -- is left to the
      // general step below.
      if (cnt[L] == 1 && !matched && cn.pos > start_limit) {
        uint32_t pc = S.pc(L, 0);
#if defined(__CUDA_ARCH__)
        if (Store::coop) {
          // Warp-cooperative form (verify_small_kernel: the 32 lanes of a warp execute ONE run redundantly, with
          // identical registers): lane j tests byte pos+j against instruction pc+j, a ballot finds how far the
          // stretch holds, and 32 units cost one round of loads instead of 32 dependent ones.  ASCII only -- a
          // byte >= 0x80 ends the batch and goes through the general step (full UTF-8 decode).
          const uint32_t lane = threadIdx.x & 31u;
          while (!c.pending) {
            const uint32_t bj = pos + lane < len ? (uint32_t)m[pos + lane] : 0x100u;
            const uint32_t ij = pc + lane < plen ? prog[pc + lane] : (uint32_t)OP_MATCH;
            const uint32_t nj = pc + lane + 1 < plen ? prog[pc + lane + 1] & 0xffu : (uint32_t)OP_MATCH;
            const uint32_t op = ij & 0xffu, arg = ij >> 8;
            const bool cons = op == OP_CHAR || op == OP_SET || op == OP_ANY;
            bool ok = false;
            if (cons && bj < 0x80u)
              ok = op == OP_CHAR ? bj == arg : op == OP_ANY ? !(bj == 0x0au || bj == 0x0du) : ((rs.sets[(size_t)arg * 6 + (bj >> 5)] >> (bj & 31u)) & 1u) != 0;
            const uint32_t okm = __ballot_sync(0xffffffffu, ok);
            const uint32_t good = okm & __ballot_sync(0xffffffffu, nj == OP_CHAR || nj == OP_SET || nj == OP_ANY);
            const uint32_t J = good == 0xffffffffu ? 32u : (uint32_t)__ffs(~good) - 1u;      // units consumed by this batch
            const uint32_t bJ = __shfl_sync(0xffffffffu, bj, J & 31u), bP = __shfl_sync(0xffffffffu, bj, (J + 31u) & 31u);
            if (J) { prev = (int)bP; pos += J; upos += J; pc += J; }
            if (J == 32u) {                                   // all 32 held: reload and go on (cur is fixed up on exit)
              c.pos = pos; c.pending = 0; cn = c; cur = next_unit(m, len, cn);
              if (cur < 0) { cnt[L] = 0; break; }
              continue;
            }
            // position J: its instruction is consuming (instruction pc + J - 1 said so, or it is the thread's own)
            c.pos = pos; c.pending = 0; cn = c;
            if (bJ == 0x100u) { cur = -1; cnt[L] = 0; break; }                      // end of message: the thread dies
            if (bJ < 0x80u) { cur = (int)bJ; cn.pos = pos + 1; if (!((okm >> J) & 1u)) cnt[L] = 0; }   // ASCII mismatch: dies
            else cur = next_unit(m, len, cn);                                       // non-ASCII: the general step decides
            break;
          }
          if (c.pending) {}                                   // (mid-surrogate: nothing done, general step)
        } else
#endif
        for (;;) {
          const uint32_t ins = prog[pc], op = ins & 0xff, arg = ins >> 8;
          const bool ok = op == OP_CHAR ? ((uint32_t)cur == arg)
                        : op == OP_ANY ? !(cur == 0x0a || cur == 0x0d || cur == 0x2028 || cur == 0x2029)
                        : in_set(rs, arg, cur);
          if (!ok) { cnt[L] = 0; break; }
          const uint32_t nop = prog[pc + 1] & 0xff;
          if (nop != OP_CHAR && nop != OP_SET && nop != OP_ANY) break;
          prev = cur; pos = cn.pos; upos++; c = cn; cur = next_unit(m, len, cn); pc++; CG_VM_STAT(4);
          if (cur < 0) { cnt[L] = 0; break; }
        }
        if (cnt[L] == 0) continue;                          // the thread died: the loop head ends the search
        S.pc(L, 0) = (uint16_t)pc;
      }
      Cursor cnn = cn; int nxt = next_unit(m, len, cnn);  // unit after cur
      int N = L ^ 1; cnt[N] = 0; bump_gen();
      uint32_t npos = cn.pos, nupos = upos + 1;
      bool cut = false;
      CG_VM_STAT(3);
      for (uint32_t k = 0; k < cnt[L] && !cut; k++) {
        CG_VM_STAT(2);
        uint32_t pc = S.pc(L, k); uint32_t ins = prog[pc], op = ins & 0xff, arg = ins >> 8;
        bool ok = op == OP_CHAR ? ((uint32_t)cur == arg)
                : op == OP_ANY ? !(cur == 0x0a || cur == 0x0d || cur == 0x2028 || cur == 0x2029)
                : in_set(rs, arg, cur);
        if (ok) {
          // successor is itself a consuming instruction (inside a literal / class run): queue it directly, no closure walk
          const uint32_t npc = pc + 1, nop = prog[npc] & 0xff;
          if (nop == OP_CHAR || nop == OP_SET || nop == OP_ANY) {
            const uint16_t DONE = (uint16_t)(gen * 2 + 1);
            if (S.mark(npc) != DONE) {
              S.mark(npc) = DONE;
              const uint32_t q = cnt[N];
              if (q < Store::cap) { S.pc(N, q) = (uint16_t)npc; S.st(N, q) = S.st(L, k); cnt[N] = q + 1; } else err |= ERR_VM_LIST;
            }
          } else cut = add(N, npc, S.st(L, k), cur, nxt, npos, nupos);
        }
      }
      // new lowest-priority start at npos -- skipped when the unit there cannot begin a match
      // (first-unit filter; a start thread that cannot consume its first unit dies immediately,
      // and rules that can match the empty string have an all-ones filter)
      if (!cut && !matched && npos <= start_limit && !(nxt >= 0 && nxt < 128 && !((first[nxt >> 5] >> (nxt & 31)) & 1u)))
        add(N, 0, npos, cur, nxt, npos, nupos);
      L = N; prev = cur; pos = npos; upos = nupos; c = cn; cn = cnn; cur = nxt;
    }
    return matched;
  }
};

template <int CAP> using VMT = VMS<LocalStore<CAP>>;
using VM = VMT<kMaxProgLen>;

// number of UTF-16 units in m[0, upto)
CG_HD_NOINLINE uint32_t count_units(const uint8_t* __restrict__ m, uint32_t len, uint32_t upto) {
  Cursor c{0, 0}; uint32_t k = 0;
  while (c.pos < upto) { if (next_unit(m, len, c) < 0) break; k++; }
  return k;
}
// cursor positioned at byte `at` (a unit boundary, possibly the middle of a 4-byte sequence)
CG_HD_NOINLINE Cursor cursor_at(const uint8_t* __restrict__ m, uint32_t len, uint32_t at) {
  uint32_t q = at; int back = 0;
  while (q > 0 && back < 3 && q < len && (m[q] & 0xc0) == 0x80) { q--; back++; }
  Cursor c{q, 0};
  while (c.pos < at) { Cursor t = c; if (next_unit(m, len, t) < 0) break; if (t.pos > at) break; c = t; }
  if (c.pos != at) { c.pos = at; c.pending = 0; }
  return c;
}


// All matches of one rule in one message -- the exec loop of registry.ts:225-236 (SPANS) or just
// RegExp.test (context.ts:9-25, !SPANS).  sink.span(start_byte, end_byte, start16, end16).
template <bool SPANS, class VMX, class Sink>
CG_HD_NOINLINE bool run_rule(VMX& vm, const DevRuleset& rs, uint32_t rule, const uint8_t* __restrict__ m, uint32_t len, Sink& sink) {
  vm.prog = vm.prog_override ? vm.prog_override : rs.prog + rs.rule_prog_off[rule]; vm.plen = rs.rule_prog_off[rule + 1] - rs.rule_prog_off[rule];
  if (vm.plen == 0) return false;                     // rule failed to compile: never matches
  if (vm.plen > VMX::capacity()) { vm.err |= ERR_VM_LIST; return false; }
  vm.clear_marks();
  const uint32_t* first = rs.rule_first + (size_t)rule * 8;
  uint32_t from = 0, from16 = 0; int prev = -1; Cursor c{0, 0};
  bool any = false;
  for (;;) {
    if (!vm.search(m, len, from, from16, prev, first, c)) break;
    any = true;
    if (!SPANS) break;
    sink.span(vm.m_start, vm.m_end, count_units(m, len, vm.m_start), vm.m_end16);
    VMX::rejoin();
    // lastIndex = end; after an empty match advance one code unit
    from = vm.m_end; from16 = vm.m_end16; prev = vm.m_prev; c = cursor_at(m, len, from);
    if (vm.m_end == vm.m_start) {
      Cursor t = c; int u = next_unit(m, len, t);
      if (u < 0) break;
      prev = u; from = t.pos; from16++; c = t;
    }
  }
  return any;
}

// the unit that ends at byte `at` (a unit boundary), or -1 at the start of the message
CG_HD_NOINLINE int unit_before(const uint8_t* __restrict__ m, uint32_t len, uint32_t at) {
  if (at == 0) return -1;
  uint32_t q = at - 1; int back = 0;
  while (q > 0 && back < 3 && (m[q] & 0xc0) == 0x80) { q--; back++; }
  Cursor c{q, 0}; int last = -1;
  while (c.pos < at || (c.pending && c.pos + 2 <= at)) { int u = next_unit(m, len, c); if (u < 0) break; last = u; }
  return last;
}

// RegExp.test restricted to the matches that contain the confirmed factor occurrence m[t0, t0+flen):
// such a match lies inside the "island" of bytes from the rule's alphabet around the factor, so the
// search starts at the island's first byte, spawns no start thread beyond t0 and stops at the first
// MATCH.  Every true match contains some confirmed occurrence of a factor, hence testing every
// occurrence this way is exactly RegExp.test(message).
template <class VMX>
CG_HD_NOINLINE bool test_at_factor(VMX& vm, const DevRuleset& rs, uint32_t rule, const uint8_t* __restrict__ m, uint32_t len,
                                   uint32_t t0, uint32_t pre_units) {
  vm.prog = vm.prog_override ? vm.prog_override : rs.prog + rs.rule_prog_off[rule]; vm.plen = rs.rule_prog_off[rule + 1] - rs.rule_prog_off[rule];
  if (vm.plen == 0) return false;
  if (vm.plen > VMX::capacity()) { vm.err |= ERR_VM_LIST; return false; }
  vm.clear_marks();
  const uint32_t* first = rs.rule_first + (size_t)rule * 8;
  // bytes the part of a match BEFORE this factor can consist of (per factor; falls back to the rule's alphabet)
  const uint32_t pa = pre_units >> 16; pre_units &= 0xffffu;
  const uint32_t* alpha = pa != 0xffffu ? rs.bytesets + (size_t)pa * 8 : rs.rule_alpha + (size_t)rule * 8;
  if (t0 > len) t0 = len;
  uint32_t s = t0;
  while (s > 0 && ((alpha[m[s - 1] >> 5] >> (m[s - 1] & 31)) & 1u)) s--;
  // s is a unit boundary: it is 0 or follows a byte outside the alphabet
  if (pre_units < 0xffffu) {
    // the pattern puts at most pre_units units (<= 3 bytes each) before the factor
    uint32_t sb = t0 > 3u * pre_units ? t0 - 3u * pre_units : 0u;
    for (int k = 0; k < 3 && sb > 0 && (m[sb] & 0xc0) == 0x80; k++) sb--;     // back to the start of a character
    if (sb > s) s = sb;
  }
  // a factor occurrence may begin in the middle of a character (byte-level wildcard shapes): start at the
  // character's first byte -- that only adds earlier, legitimate start positions
  for (int k = 0; k < 3 && s > 0 && s < len && (m[s] & 0xc0) == 0x80; k++) s--;
  Cursor c = cursor_at(m, len, s);
  return vm.search(m, len, s, 0, unit_before(m, len, s), first, c, t0, true);
}

}  // namespace cg
