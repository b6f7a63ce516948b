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

// ---- prefilter: level 1 = Mealy DFA over short *windows* of every rule's necessary factors
// (shared memory, one lookup per byte); level 2 = exact confirmation of the full factor (up to 16
// byte-set elements) at the flagged position; only confirmed (message, rule) pairs reach the VM.
constexpr int kMaxWindow = 8;
constexpr int kMaxFactorElems = 16;

struct PrefilterOptions {
  int mode = 2;              // 0 = direct 7-bit (128 cols), 1 = byte->class LUT (<=64), 2 = folded 6-bit (64, SWAR), 3 = folded 5-bit (32, SWAR),
                             // 4 = lane-private fingerprint table over 4-byte windows (falls back to 2 when a rule set does not fit)
  int fp_buckets = 1024;     // mode 4: 2-way buckets per lane-bank replica
  int max_states = 24576;    // total level-1 states (rows beyond the shared-memory budget stay in L2-resident HBM)
  int max_classes = 64;      // LUT mode only
  int max_window = kMaxWindow;
};
struct FullFactor {
  uint32_t rule;
  uint8_t len, win_off, win_len, exact;   // exact: a confirmed occurrence proves the rule matches (RegExp.test)
  uint16_t elem[kMaxFactorElems];         // byte-set ids
  uint16_t pre;                           // max units of a match before the factor (0xffff = unbounded)
  uint16_t pre_alpha;                     // byte-set id: what the part of a match before the factor can consist of
};
struct Prefilter {
  int mode = 2;
  int ncols = 64;                     // columns per row (power of two)
  int nstates = 1;                    // states are numbered breadth-first: shallow (hot) states first
  int window_min = 0, window_max = 0; // level-1 window lengths actually used
  std::vector<uint16_t> table;        // nstates * ncols; bit 15 = accepting transition, bits 0-14 = next state
  uint8_t lut[256];                   // byte -> column (used by the device only in LUT mode)
  std::vector<uint32_t> acc_index;    // nstates * ncols: accept id of an accepting transition, else 0xffffffff
  std::vector<uint32_t> acc_offsets;  // CSR over accept ids -> factor ids
  std::vector<uint32_t> acc_factors;
  std::vector<FullFactor> factors;
  std::vector<uint32_t> bytesets;     // 8 words per 256-bit set
  std::vector<uint32_t> always_rules; // rules without usable factors: candidates for every message
  // mode 4: fingerprint table.  bucket = umulhi(h, fp_buckets), fingerprint = (h >> 8) & 0xffff, h = window * fp_mult
  uint32_t fp_buckets = 0, fp_mult = 0, fp_keys = 0;
  std::vector<uint32_t> fp_table;     // fp_buckets words: two 16-bit fingerprints (0xffff = empty)
  std::vector<uint32_t> fp_acc;       // 2 * fp_buckets accept ids (0xffffffff = empty)
  std::vector<uint32_t> trig_bytes;   // mode 4: up to two single-byte triggers for factors without an enumerable window
  std::vector<uint32_t> trig_acc;     //         their accept ids
};

// the byte -> 6-bit symbol map of mode 4: fold6 with the digit columns collapsed to two symbols
inline uint32_t fp_fold(uint32_t b) { uint32_t c = (b & 0x1fu) | ((b >> 1) & 0x20u); if ((c & 0x30u) == 0x10u) c &= 0x38u; return c; }
bool build_prefilter(const std::vector<CompiledRule>& rules, const PrefilterOptions& opt, Prefilter* out, std::string* err);

}  // namespace cg
