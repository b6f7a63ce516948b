// prefilter_dev.h -- level-2 of the prefilter: exact confirmation of a full factor at a position
// flagged by the level-1 DFA.  CG_HD so that tests/native/vm_harness.cpp restates the scan
// kernel's candidate logic with the very same code.
#pragma once
#include <cstdint>
#include "kernels.h"
#include "pike_vm.h"   // CG_HD macros

namespace cg {

CG_HD uint32_t l1_col(uint32_t mode, const uint8_t* __restrict__ lut, uint32_t b) {
  return mode == 0 ? (b & 0x7fu) : mode == 2 ? ((b & 0x1fu) | ((b >> 1) & 0x20u)) : mode == 3 ? (b & 0x1fu) : (uint32_t)lut[b];   // mode 4: lut = fp_fold
}

CG_HD bool byte_in_set(const DevRuleset& rs, uint32_t sid, uint32_t b) {
  return (rs.bytesets[(size_t)sid * 8 + (b >> 5)] >> (b & 31)) & 1u;
}

constexpr int kMaxFactorElemsDev = 16;      // = kMaxFactorElems (rulec.h): elements per full factor

// Does factor `fw` occur in message m[0,len) with its level-1 window ending at byte `pend`?
CG_HD_NOINLINE bool confirm_factor(const DevRuleset& rs, const uint32_t* __restrict__ fw, const uint8_t* __restrict__ m,
                                   uint32_t len, uint32_t pend) {
  const uint32_t meta = fw[1], flen = meta & 0xff, wend = ((meta >> 8) & 0xff) + ((meta >> 16) & 0xff);   // window end (exclusive) inside the factor
  if (pend + 1 < wend) return false;
  const uint32_t t0 = pend + 1 - wend;                 // message offset of factor element 0
  if (t0 + flen > len) return false;
  // No early exit on purpose: with all (up to 16) element checks unrolled and independent, their loads -- the factor's
  // set ids, the message bytes, one byte-set word each -- are issued back to back and overlap, instead of forming a
  // chain of 2 x flen dependent L2 round trips (this kernel is pure latency).
  bool ok = true;
#pragma unroll
  for (uint32_t k = 0; k < (uint32_t)kMaxFactorElemsDev; k++) {
    if (k < flen) {
      const uint32_t sid = (fw[2 + (k >> 1)] >> (16 * (k & 1))) & 0xffffu;
      ok = ok & byte_in_set(rs, sid, m[t0 + k]);
    }
  }
  return ok;
}

// the factors of accept id `aid` are confirmed at message offset `pend` (window's last byte)
template <class Sink>
CG_HD_NOINLINE void accept_id(const DevRuleset& rs, uint32_t aid, const uint8_t* __restrict__ m, uint32_t len, uint32_t pend,
                              bool want_spans, Sink& sink) {
  if (aid == 0xffffffffu) return;
  for (uint32_t k = rs.acc_offsets[aid]; k < rs.acc_offsets[aid + 1]; k++) {
    const uint32_t* fw = rs.factors + (size_t)rs.acc_factors[k] * 12;
    if (confirm_factor(rs, fw, m, len, pend)) {
      const uint32_t meta = fw[1], t0 = pend + 1 - (((meta >> 8) & 0xff) + ((meta >> 16) & 0xff));   // factor start in the message
      if ((meta >> 24) && !want_spans) sink.direct(fw[0]); else sink.candidate(fw[0], t0, fw[10] | (fw[11] << 16));     // max prefix units | prefix-alphabet set id << 16
    }
  }
}

// DFA modes: an accepting level-1 transition (state, col) was taken on the byte at message offset `pend`.
template <class Sink>
CG_HD_NOINLINE void l1_accept(const DevRuleset& rs, uint32_t state, uint32_t col, const uint8_t* __restrict__ m, uint32_t len,
                              uint32_t pend, bool want_spans, Sink& sink) {
  accept_id(rs, rs.acc_index[((size_t)state << rs.ncols_log2) + col], m, len, pend, want_spans, sink);
}

// mode 4: the window hash `h` hit a fingerprint at message offset `pend`
template <class Sink>
CG_HD_NOINLINE void fp_accept(const DevRuleset& rs, uint32_t h, const uint8_t* __restrict__ m, uint32_t len, uint32_t pend,
                              bool want_spans, Sink& sink) {
  const uint32_t b = (uint32_t)(((uint64_t)h * rs.fp_buckets) >> 32), fp = (h >> 8) & 0xffffu, val = rs.fp_table[b];
  for (uint32_t way = 0; way < 2; way++)
    if (((val >> (16 * way)) & 0xffffu) == fp) accept_id(rs, rs.fp_acc[2 * b + way], m, len, pend, want_spans, sink);
}

}  // namespace cg
