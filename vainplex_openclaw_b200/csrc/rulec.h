// rulec.h -- host-side rule compiler for the B200 scan path.
//
// Replaces the reference's "rule-set compile" steps, which are plain `new RegExp(...)`:
//   gov/src/policy-loader.ts:88-133  (buildPolicyIndex -> regexCache, no flags)
//   gov/src/redaction/registry.ts:222-223, 249-281 (per-call RegExp + custom patterns)
// Output is a flat device image with two parts:
//   1. per rule: a unit-level Pike-VM program (exact ECMAScript leftmost-first semantics on
//      UTF-16 code units, decoded on the fly from UTF-8 by the verify kernel), and
//   2. for the whole set: one prefilter DFA over *necessary factors* of every rule
//      (byte-set sequences every match must contain) -- sound over-approximation, every
//      byte of every message goes through exactly this table from shared memory.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace cg {

constexpr int kMaxProgLen = 1024;        // Pike instructions per rule (verify-kernel list bound)
constexpr int kMaxFactorLen = 8;         // elements per prefilter factor

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
  std::vector<FactorSeq> factors;   // necessary factors; empty => rule is an "always candidate"
  bool factors_exact = false;       // every match IS one of the factors (pure literal alternation, no assertions)
};

// flags: bit0 = ignoreCase ("i").  `src` is the JS pattern source encoded as UTF-8.
CompiledRule compile_rule(const char* src, size_t len, uint32_t flags);

// ---- prefilter DFA over all rules' factors
struct PrefilterOptions {
  int mode = 0;              // 0 = direct 7-bit index (128 columns, no LUT), 1 = byte->class LUT
  int max_states = 640;      // rows that fit the shared-memory budget
  int max_classes = 64;      // LUT mode only
  int max_factor_len = kMaxFactorLen;
};
struct Prefilter {
  int mode = 0;
  int ncols = 128;                    // columns per row (power of two)
  int nstates = 1;
  int first_accept = 1;               // states >= first_accept have outputs
  int factor_len = 0;                 // truncation that was needed to fit
  std::vector<uint16_t> table;        // nstates * ncols
  uint8_t lut[256];                   // LUT mode: byte -> column ; direct mode: b & 0x7f
  std::vector<uint32_t> out_offsets;  // per accept state (index s - first_accept): CSR into out_rules, size naccept+1
  std::vector<uint32_t> out_rules;    // rule ids
  std::vector<uint32_t> always_rules; // rules without usable factors: candidates for every message
};
bool build_prefilter(const std::vector<CompiledRule>& rules, const PrefilterOptions& opt, Prefilter* out, std::string* err);

}  // namespace cg
