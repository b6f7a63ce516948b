// kernels.h -- device-side structures and launchers shared by capi.cu and the .cu kernel files.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cg {

// Everything the scan / verify kernels need to know about a compiled rule set (device pointers).
struct DevRuleset {
  const uint8_t* image;          // shared-memory image (layout: ruleset_image.h), 16-byte aligned, size % 16 == 0
  uint32_t lut_off, row_stride;  // DFA modes: byte offset of the 256-byte LUT inside the image; bytes between table rows
  uint32_t image_bytes;
  uint32_t mode;                 // 0 = direct 7-bit columns, 1 = LUT columns, 2 = folded 6-bit, 3 = folded 5-bit columns, 4 = fingerprint table
  uint32_t max_prog_len;         // longest Pike program of the set (picks the VM capacity)
  uint32_t ncols_log2;
  uint32_t nstates;
  uint32_t hot_states;           // rows [0, hot_states) + one trap row are in the shared-memory image, the rest only in table_full
  uint32_t debug_flags;          // CG_SCAN_DEBUG (experiments only): bit 0 = skip the scan kernel's slow path (results wrong),
                                 //   bit 1 = histogram of VM cycles per event in counters[7..15]
  const uint16_t* table_full;    // complete level-1 table in HBM (L2-resident)
  const uint32_t* acc_index;     // nstates*ncols: accept id of an accepting transition
  const uint32_t* acc_offsets;   // CSR over accept ids -> factor ids
  const uint32_t* acc_factors;
  const uint32_t* factors;       // 12 words per full factor: rule, len|win_off<<8|win_len<<16|exact<<24, 16 x u16 set ids, max prefix units, pad
  const uint32_t* bytesets;      // 8 words per 256-bit byte set
  // mode 4 (fingerprint level 1): bucket = umulhi(window * fp_mult, fp_buckets); image = [256 B][fp_buckets x 32 replicated words]
  uint32_t fp_buckets, fp_mult;
  const uint32_t* fp_table;      // fp_buckets words (two 16-bit fingerprints each)
  const uint32_t* fp_acc;        // 2 * fp_buckets accept ids
  uint32_t n_trig, trig_byte[2], trig_acc[2];   // single-byte triggers (events carry pos | 0x80000000, sc = trigger index)
  const uint32_t* always_rules;  // candidates for every message
  uint32_t n_always;
  const uint32_t* prog;          // all Pike programs, concatenated
  const uint32_t* rule_prog_off; // n_rules + 1
  const uint32_t* sets;          // 6 words per unit set: ascii[4], range_off, n_ranges
  const uint16_t* set_ranges;    // inclusive lo,hi pairs
  const uint32_t* rule_first;    // 8 words per rule: first-byte bitmap
  const uint32_t* rule_alpha;    // 8 words per rule: alphabet bitmap (bytes a match can contain)
  // verdict aggregation (policy-evaluator.ts:44-146, messageContains slice): per rule its policy (index in priority order,
  // non-decreasing with the rule index) and its effect (0 allow, 1 audit, 2 deny); nullptr until cg_ruleset_set_policy
  const uint32_t* rule_policy; const uint32_t* rule_action;
  uint32_t n_rules;
  uint32_t rw;                   // bitmap words per slot = ceil(n_rules / 32)
};

// Per-call scratch in HBM.  A "slot" is one message with at least one confirmed candidate.
struct ScanWork {
  uint32_t* counters;            // 32 words: [0]=n_slots [1]=n_events [2]=n_spans [3]=error flags [4]=n_l1 (level-1 accept events) [5]=verify cursor
                                 //           [6]=slow-path entries [7..15]=debug [16]=n_units (segmented scans) [17]=n_heavy [18]=verify cursor (light events)
  uint32_t* l1_msg;              // [l1_cap] level-1 accept events queued by scan_kernel: message,
  uint32_t* l1_pos;              //          byte offset of the accepting byte inside the message,
  uint32_t* l1_sc;               //          state << 8 | column of the accepting transition (0xffffffff = "always" rules)
  uint32_t* slot_of_msg;         // [n] 0xffffffff = none yet (reset per step)
  uint32_t* slot_msg;            // [slot_cap]
  uint32_t* cand;                // [slot_cap * rw]  candidate (msg,rule) pairs already queued
  uint32_t* hit;                 // [slot_cap * rw]  verified pairs
  uint2* events;                 // [event_cap]  (slot, rule | heavy << 31) for the Pike VM
  uint32_t* heavy_idx;           // [event_cap]  indices of the events expected to run long (counters[17] of them): verified first
  uint32_t* event_pos;           // [event_cap]  policy mode: message offset of the confirmed factor's first byte
  uint32_t* event_pre;           // [event_cap]  policy mode: max units of a match before that factor (0xffff = unbounded)
  uint32_t* spans;               // [span_cap * 6] msg, rule, start_byte, end_byte, start16, end16
  // Long messages are cut into units of kSegBytes (+ kSegWarm bytes of warm-up) so that every lane has about the same
  // amount to walk: units[u] = (message, segment).  nullptr = one unit per message (short-message batches).
  uint2* units;
  uint32_t l1_cap, msg_cap, slot_cap, event_cap, span_cap, unit_cap;
};

enum : uint32_t { ERR_EVENT_OVERFLOW = 1, ERR_SPAN_OVERFLOW = 2, ERR_VM_STACK = 4, ERR_VM_LIST = 8, ERR_SLOT_OVERFLOW = 16, ERR_L1_OVERFLOW = 32,
                  ERR_UNIT_OVERFLOW = 64 /* not an error: the scan falls back to one unit per message */ };
constexpr uint32_t kSegBytes = 1024, kSegWarm = 16;     // kSegWarm >= longest level-1 window - 1 (kMaxWindow = 8), multiple of 16
constexpr uint32_t kCounterWords = 32;

// launchers (all asynchronous on `stream`); return the number of kernels launched
int launch_scan(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off, uint32_t n,
                uint64_t* d_words, bool want_spans, int sm_count, cudaStream_t stream);
// unit table of a segmented scan (counters[16] = number of units) and the longest message of a batch
int launch_plan_units(const ScanWork& w, const uint32_t* d_off, uint32_t n, cudaStream_t stream);
int launch_max_len(const uint32_t* d_off, uint32_t n, uint32_t* d_out_max, cudaStream_t stream);
// state visit histogram over n_sample evenly spaced messages (profile-guided residency)
int launch_l1_profile(const DevRuleset& rs, const uint8_t* d_bytes, const uint32_t* d_off, uint32_t n, uint32_t n_sample,
                      uint32_t* d_visits, cudaStream_t stream);
int launch_confirm(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off,
                   bool want_spans, int sm_count, cudaStream_t stream);
int launch_verify(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off,
                  bool want_spans, int sm_count, cudaStream_t stream);
// one verdict word per hit message: action | matched policies << 2 | deciding rule << 12 (cg_policy_verdict_batch)
int launch_verdicts(const DevRuleset& rs, const ScanWork& w, uint32_t* d_verdicts, int sm_count, cudaStream_t stream);
int launch_finalize(const DevRuleset& rs, const ScanWork& w, uint64_t* d_words, int sm_count, cudaStream_t stream);

// raises the dynamic shared-memory limits of every kernel once (not legal inside stream capture)
void prepare_scan_kernels();

// SHA-256 / Merkle
int launch_sha256_batch(const uint8_t* d_bytes, const uint64_t* d_off, uint32_t n, uint8_t* d_out, cudaStream_t stream);
// Merkle log: acc = node(left[k], acc) chains; audit-path verification
int launch_merkle_chain(const uint32_t* d_left, uint32_t n_left, const uint32_t* d_acc_in, uint32_t* d_acc_out, cudaStream_t stream);
int launch_merkle_verify(const uint32_t* d_leaf32, uint64_t index, uint64_t size, const uint32_t* d_path, uint32_t path_len, const uint32_t* d_root32, uint32_t* d_ok, cudaStream_t stream);
// redacted output: SHA-256 of every span, then copy + placeholder splice (one warp per message)
int launch_redact_digests(const uint8_t* d_bytes, const uint32_t* d_start, const uint32_t* d_len, uint32_t ns, uint32_t* d_out, cudaStream_t stream);
int launch_redact_splice(const uint8_t* d_bytes, const uint32_t* d_off, const uint32_t* d_out_off, const uint32_t* d_span_begin,
                         const uint32_t* d_span_start, const uint32_t* d_span_len, const uint32_t* d_span_cat, const uint32_t* d_digests,
                         uint8_t* d_out, uint32_t n, int sm_count, cudaStream_t stream);
// leaf digests of fixed-size leaves: out[i] = SHA-256(0x00 || leaf_i)
int launch_merkle_leaves_fixed(const uint8_t* d_bytes, uint64_t leaf_len, uint64_t n, uint32_t* d_out, cudaStream_t stream);
int launch_merkle_leaves_var(const uint8_t* d_bytes, const uint64_t* d_off, uint64_t n, uint32_t* d_out, cudaStream_t stream);
// one tree level: out[i] = node(in[2i], in[2i+1]) ; an unpaired last node is copied
int launch_merkle_level(const uint32_t* d_in, uint64_t n_in, uint32_t* d_out, cudaStream_t stream);
// reduce up to `levels` levels inside one kernel (aligned blocks of 2^levels nodes -> 1 node each)
int launch_merkle_reduce(const uint32_t* d_in, uint64_t n_in, uint32_t levels, uint32_t* d_out, cudaStream_t stream);

}  // namespace cg
