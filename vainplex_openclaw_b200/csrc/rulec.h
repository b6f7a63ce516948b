// rulec.h -- host-side rule compiler for the B200 scan path.
//
// Replaces the reference's "rule-set compile" steps, which are plain `new RegExp(...)`:
//   gov/src/policy-loader.ts:88-133  (buildPolicyIndex -> regexCache, no flags)
//   gov/src/redaction/registry.ts:222-223, 249-281 (per-call RegExp + custom patterns)
// Output is a flat device image with two parts:
//   1. per rule: a unit-level Pike-VM program (exact ECMAScript leftmost-first semantics on
//      UTF-16 code units, decoded on the fly from UTF-8 by the verify kernel), and
//   2. for the whole set: one stateless gram filter over *necessary factors* of every rule
//      (byte-set sequences every match must contain) -- sound over-approximation, every
//      byte of every message is hashed against exactly this bitmap in shared memory.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace cg {

constexpr int kMaxProgLen = 1024;        // Pike instructions per rule (verify-kernel list bound)
constexpr int kVmStackLimit = 192;       // entries of the VM's closure stack (pike_vm.h); rules that could exceed it are rejected at compile time

// ---- Pike program encoding: one uint32 per instruction = op | arg << 8
enum Op : uint32_t {
  OP_CHAR = 0,        // arg = code unit
  OP_SET = 1,         // arg = set id (rule-set global)
  OP_ANY = 2,         // '.' : any unit except \n \r U+2028 U+2029
  OP_SPLIT_NEXT = 3,  // try pc+1 first, then arg
  OP_SPLIT_JUMP = 4,  // try arg first, then pc+1
  OP_JMP = 5,         // arg = target
  OP_MATCH = 6,
  OP_BOL = 7, OP_EOL = 8, OP_WORDB = 9, OP_NWORDB = 10,
  OP_LOOKAHEAD = 11,      // arg = set id ; next unit in set
  OP_NLOOKAHEAD = 12,
  OP_LOOKBEHIND = 13,     // previous unit in set
  OP_NLOOKBEHIND = 14,
  OP_EMPTYCHK = 15,       // arg = pc of the optional iteration's SPLIT: die if that SPLIT is in progress at this position
  OP_JMP_BACK = 16,       // loop back-edge to the SPLIT at arg: an iteration that consumed nothing dies here
};

struct UnitSet {            // set of UTF-16 code units
  uint32_t ascii[4];        // bitmap for units 0..127
  std::vector<uint16_t> ranges;  // sorted inclusive pairs lo,hi for units >= 128
  bool operator==(const UnitSet& o) const {
    return ascii[0] == o.ascii[0] && ascii[1] == o.ascii[1] && ascii[2] == o.ascii[2] && ascii[3] == o.ascii[3] && ranges == o.ranges;
  }
};

struct ByteSet {
  uint64_t w[4] = {0, 0, 0, 0};
  void set(int b) { w[b >> 6] |= 1ull << (b & 63); }
  bool has(int b) const { return (w[b >> 6] >> (b & 63)) & 1; }
  bool empty() const { return !(w[0] | w[1] | w[2] | w[3]); }
  bool operator==(const ByteSet& o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
  bool operator<(const ByteSet& o) const { for (int i = 0; i < 4; i++) if (w[i] != o.w[i]) return w[i] < o.w[i]; return false; }
};
using FactorSeq = std::vector<ByteSet>;

enum RuleStatus : int32_t {
  RULE_OK = 0,
  RULE_ERR_SYNTAX = -1,        // `new RegExp(src)` would throw SyntaxError
  RULE_ERR_UNSUPPORTED = -2,   // valid JS, outside the supported subset
  RULE_ERR_TOO_LARGE = -3,     // program longer than kMaxProgLen
};

struct CompiledRule {
  int32_t status = RULE_OK;
  std::string error;
  std::vector<uint32_t> prog;       // Pike program; entry pc = 0
  std::vector<UnitSet> sets;        // rule-local sets; OP_SET args index this (remapped globally later)
  bool nullable = false;            // matches the empty string somewhere
  ByteSet first_bytes;              // bytes that can start a match (all-ones when nullable)
  ByteSet alphabet;                 // every byte any consuming instruction could accept (over-approximation)
  std::vector<FactorSeq> factors;   // necessary factors; empty => rule is an "always candidate"
  int factor_pre = 1 << 20;         // max units of a match that can precede the factor occurrence (>= 0xffff: unbounded)
  ByteSet factor_pre_alpha;         // bytes the part of a match before the factor can consist of
  bool factors_exact = false;       // every match IS one of the factors (pure literal alternation, no assertions)
};

// flags: bit0 = ignoreCase ("i").  `src` is the JS pattern source encoded as UTF-8.
CompiledRule compile_rule(const char* src, size_t len, uint32_t flags);

// Bit-parallel form of a rule's Pike program (layout: bitprog.h).  false when the program has more than 127 consuming
// instructions, lookaround over sets with non-ASCII members or more than one set per direction: such rules stay with the Pike VM.
bool build_bitprog(const CompiledRule& r, std::vector<uint64_t>* out);

// ---- prefilter: a stateless two-level filter over every rule's *necessary factors* (byte-set sequences every
// match must contain; <= kMaxFactorElems elements are kept per factor).
//   level 1a  every aligned 4-byte gram of the batch buffer (stride 4), or every even-offset gram (stride 2), is
//             folded (gram_fold: case bit and bit 7 dropped, digits collapsed), hashed and tested against a bitmap
//             in shared memory.  For every factor and every residue r of its start position modulo the stride, ONE
//             gram position relative to the factor is registered: all folded byte combinations the factor allows
//             there (bytes outside the factor are wildcards) are set in the bitmap.
//   level 1b  a flagged gram is looked up in a hash table of (masked key -> factor, gram offset); the factor's byte
//             sets are then compared exactly at that position (what level 2 used to do in its own kernel).
// Factors no gram can cover within the key budget fall back to a single-byte trigger (at most two distinct bytes,
// compared SWAR-style), and rules without any usable factor are candidates for every message.
constexpr int kMaxFactorElems = 16;
constexpr int kGramLen = 4;

// byte -> 6-bit-ish symbol of the gram hash (49 distinct values).  z = (b & 0x5f) ^ 0x10; z < 0x10 (digits and
// : ; < = > ?, plus the C0/C1 bytes sharing those low bits) collapses to 0.  Mirrors gram_fold_word() in gram_filter.h.
inline uint32_t gram_fold(uint32_t b) { uint32_t z = (b & 0x5fu) ^ 0x10u; return z < 0x10u ? 0u : z; }
constexpr uint32_t kGramMult = 0x9E3779B1u;          // bitmap hash: h = key * kGramMult
constexpr uint32_t kGramMult2 = 0x85EBCA77u;         // level-1b bucket hash

struct PrefilterOptions {
  int stride = 0;                 // 0 = choose: 4 when every factor is coverable within max_keys, else 2
  uint32_t max_keys = 24576;      // bitmap keys that still allow stride 4
  uint32_t max_keys_per_gram = 4096;
};
struct FullFactor {
  uint32_t rule;
  uint8_t len, exact;                     // exact: a confirmed occurrence proves the rule matches (RegExp.test)
  uint16_t elem[kMaxFactorElems];         // byte-set ids
  uint16_t pre;                           // max units of a match before the factor (0xffff = unbounded)
  uint16_t pre_alpha;                     // byte-set id: what the part of a match before the factor can consist of
};
struct GramEntry {                        // one per (factor, residue of the factor's start modulo the stride)
  uint32_t key, mask;                     // folded gram with the class / wildcard bytes zeroed; 0xff per byte that is part of the key
  uint32_t factor; int32_t off;           // gram start - factor start, -3 .. len-1
};
struct Prefilter {
  int stride = 4;
  std::vector<uint32_t> keys;         // every folded gram the bitmap must flag (sorted, unique)
  std::vector<GramEntry> entries;
  std::vector<uint32_t> shapes;       // distinct entry masks
  std::vector<FullFactor> factors;
  std::vector<uint32_t> bytesets;     // 8 words per 256-bit set
  std::vector<uint32_t> always_rules; // rules without usable factors: candidates for every message
  std::vector<uint32_t> trig_bytes;   // up to two single-byte triggers for factors no gram covers
  std::vector<uint32_t> trig_offsets; // CSR over triggers -> trig_list
  std::vector<uint32_t> trig_list;    // factor | element index of the trigger byte << 20
  uint32_t min_factor_len = 0, max_factor_len = 0;
};

bool build_prefilter(const std::vector<CompiledRule>& rules, const PrefilterOptions& opt, Prefilter* out, std::string* err);

}  // namespace cg
