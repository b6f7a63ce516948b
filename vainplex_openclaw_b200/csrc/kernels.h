// kernels.h -- device-side structures and launchers shared by capi.cu and the .cu kernel files.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cg {

// Everything the scan / verify kernels need to know about a compiled rule set (device pointers).
struct DevRuleset {
  const uint8_t* image;          // [lut 256 B][table nstates*ncols u16], 16-byte aligned, size % 16 == 0
  uint32_t image_bytes;
  uint32_t mode;                 // 0 = direct 7-bit columns, 1 = LUT columns
  uint32_t ncols_log2;
  uint32_t nstates, first_accept;
  const uint32_t* out_offsets;   // CSR over accept states
  const uint32_t* out_rules;
  const uint32_t* always_rules;  // candidates for every message
  uint32_t n_always;
  const uint32_t* prog;          // all Pike programs, concatenated
  const uint32_t* rule_prog_off; // n_rules + 1
  const uint32_t* sets;          // 6 words per set: ascii[4], range_off, n_ranges
  const uint16_t* set_ranges;    // inclusive lo,hi pairs
  const uint32_t* rule_first;    // 8 words per rule: first-byte bitmap
  uint32_t n_rules;
  uint32_t rw;                   // bitmap words per slot = ceil(n_rules / 32)
};

// Per-call scratch in HBM.  A "slot" is one message that produced at least one candidate.
struct ScanWork {
  uint32_t* counters;            // [0]=n_slots [1]=n_events [2]=n_spans [3]=error flags
  uint32_t* slot_msg;            // [slot_cap]
  uint32_t* cand;                // [slot_cap * rw]  candidate (msg,rule) pairs already queued
  uint32_t* hit;                 // [slot_cap * rw]  verified pairs
  uint2* events;                 // [event_cap]  (slot, rule)
  uint32_t* spans;               // [span_cap * 6] msg, rule, start_byte, end_byte, start16, end16
  uint32_t slot_cap, event_cap, span_cap;
};

enum : uint32_t { ERR_EVENT_OVERFLOW = 1, ERR_SPAN_OVERFLOW = 2, ERR_VM_STACK = 4, ERR_VM_LIST = 8, ERR_SLOT_OVERFLOW = 16 };

// launchers (all asynchronous on `stream`); return the number of kernels launched
int launch_scan(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off, uint32_t n,
                uint64_t* d_words, int sm_count, cudaStream_t stream);
int launch_verify(const DevRuleset& rs, const ScanWork& w, const uint8_t* d_bytes, const uint32_t* d_off,
                  bool want_spans, int sm_count, cudaStream_t stream);
int launch_finalize(const DevRuleset& rs, const ScanWork& w, uint64_t* d_words, int sm_count, cudaStream_t stream);

// SHA-256 / Merkle
int launch_sha256_batch(const uint8_t* d_bytes, const uint64_t* d_off, uint32_t n, uint8_t* d_out, cudaStream_t stream);
// leaf digests of fixed-size leaves: out[i] = SHA-256(0x00 || leaf_i)
int launch_merkle_leaves_fixed(const uint8_t* d_bytes, uint64_t leaf_len, uint64_t n, uint32_t* d_out, cudaStream_t stream);
int launch_merkle_leaves_var(const uint8_t* d_bytes, const uint64_t* d_off, uint64_t n, uint32_t* d_out, cudaStream_t stream);
// one tree level: out[i] = node(in[2i], in[2i+1]) ; an unpaired last node is copied
int launch_merkle_level(const uint32_t* d_in, uint64_t n_in, uint32_t* d_out, cudaStream_t stream);
// reduce up to `levels` levels inside one kernel (aligned blocks of 2^levels nodes -> 1 node each)
int launch_merkle_reduce(const uint32_t* d_in, uint64_t n_in, uint32_t levels, uint32_t* d_out, cudaStream_t stream);

}  // namespace cg
