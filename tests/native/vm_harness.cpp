// tests/native/vm_harness.cpp -- TEST INFRASTRUCTURE ONLY (built into tests/native/_build/, never
// into the product library).  Runs the product's rule compiler, prefilter tables and the *same*
// Pike-VM source (csrc/pike_vm.h, compiled for the host here) on the CPU so that the CPU test tier
// (-m "not gpu") can compare them with the oracle before any GPU time is spent:
//   * compile status / program shape per rule
//   * prefilter soundness: every true (message, rule) hit must be a prefilter candidate
//   * exact spans of the VM == oracle/jsre.c spans
// The product path never links this file; libopenclaw_gov.so only runs the VM inside verify_kernel.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../vainplex_openclaw_b200/csrc/pike_vm.h"
#include "../../vainplex_openclaw_b200/csrc/prefilter_dev.h"
#include "../../vainplex_openclaw_b200/csrc/ruleset_image.h"

using namespace cg;

struct Harness {
  HostImage H;
  DevRuleset d{};
  std::vector<uint32_t> pad;
};

struct VecSink {
  std::vector<uint32_t>* v;
  void span(uint32_t sb, uint32_t eb, uint32_t s16, uint32_t e16) { v->push_back(sb); v->push_back(eb); v->push_back(s16); v->push_back(e16); }
};

extern "C" {

void* harness_create(const char** srcs, const uint32_t* lens, const uint32_t* flags, uint32_t n, int mode, uint32_t budget_kb,
                     int max_factor_len, int32_t* status) {
  Harness* h = new Harness();
  std::vector<RuleSrc> src(n);
  for (uint32_t i = 0; i < n; i++) src[i] = RuleSrc{srcs[i], lens[i], flags[i]};
  ImageOptions io; io.mode = mode; if (budget_kb) io.budget_bytes = (size_t)budget_kb * 1024; if (getenv("CG_MAX_STATES")) io.max_states = atoi(getenv("CG_MAX_STATES")); if (max_factor_len) io.max_window = max_factor_len;
  std::string err;
  if (!build_host_image(src.data(), n, io, &h->H, &err)) { delete h; return nullptr; }
  for (uint32_t i = 0; i < n; i++) if (status) status[i] = h->H.rules[i].status;
  HostImage& H = h->H; DevRuleset& d = h->d;
  // pad host vectors the same way capi.cu pads the device copies
  H.sets.resize(H.sets.size() + 8, 0); H.ranges.resize(H.ranges.size() + 8, 0); H.first.resize(H.first.size() + 8, 0);
  H.pf.acc_index.resize(H.pf.acc_index.size() + 4, 0xffffffffu); H.pf.acc_offsets.resize(H.pf.acc_offsets.size() + 4, 0);
  H.pf.acc_factors.resize(H.pf.acc_factors.size() + 4, 0); H.factor_words.resize(H.factor_words.size() + 24, 0);
  H.pf.bytesets.resize(H.pf.bytesets.size() + 8, 0);
  uint32_t n_always = (uint32_t)H.pf.always_rules.size();
  H.pf.always_rules.resize(H.pf.always_rules.size() + 4, 0);
  d.image = H.image.data(); d.image_bytes = (uint32_t)H.image.size(); d.mode = (uint32_t)H.pf.mode;
  d.ncols_log2 = 0; while ((1 << d.ncols_log2) < H.pf.ncols) d.ncols_log2++;
  d.nstates = (uint32_t)H.pf.nstates; d.hot_states = H.hot_states; d.lut_off = H.lut_off; d.row_stride = H.row_stride; d.table_full = H.pf.table.data();
  d.acc_index = H.pf.acc_index.data(); d.acc_offsets = H.pf.acc_offsets.data(); d.acc_factors = H.pf.acc_factors.data();
  d.factors = H.factor_words.data(); d.bytesets = H.pf.bytesets.data();
  d.always_rules = H.pf.always_rules.data(); d.n_always = n_always;
  H.pf.fp_table.resize(H.pf.fp_table.size() + 4, 0); H.pf.fp_acc.resize(H.pf.fp_acc.size() + 4, 0xffffffffu);
  d.n_trig = (uint32_t)H.pf.trig_bytes.size(); for (uint32_t q = 0; q < 2; q++) { d.trig_byte[q] = q < d.n_trig ? H.pf.trig_bytes[q] : 0; d.trig_acc[q] = q < d.n_trig ? H.pf.trig_acc[q] : 0xffffffffu; }
  d.fp_buckets = H.pf.fp_buckets; d.fp_mult = H.pf.fp_mult; d.fp_table = H.pf.fp_table.data(); d.fp_acc = H.pf.fp_acc.data();
  d.prog = H.prog.data(); d.rule_prog_off = H.prog_off.data(); d.sets = H.sets.data(); d.set_ranges = H.ranges.data();
  H.alpha.resize(H.alpha.size() + 8, 0);
  d.rule_first = H.first.data(); d.rule_alpha = H.alpha.data(); d.n_rules = n; d.rw = (n + 31) / 32; if (!d.rw) d.rw = 1;
  return h;
}
void harness_destroy(void* p) { delete (Harness*)p; }

// info: [0]=nstates [1]=n_factors [2]=ncols [3]=window_min|window_max<<8 [4]=n_always [5]=image_bytes [6]=program words
void harness_info(void* p, uint32_t* out) {
  Harness* h = (Harness*)p;
  out[0] = h->H.pf.mode == 4 ? h->H.pf.fp_keys : h->d.nstates; out[7] = (uint32_t)h->H.pf.mode; out[1] = (uint32_t)h->H.pf.factors.size(); out[2] = (uint32_t)h->H.pf.ncols;
  out[3] = (uint32_t)h->H.pf.window_min | ((uint32_t)h->H.pf.window_max << 8);
  out[4] = h->d.n_always; out[5] = h->d.image_bytes; out[6] = (uint32_t)h->H.prog.size();
}
const char* harness_rule_error(void* p, uint32_t rule) { return ((Harness*)p)->H.rules[rule].error.c_str(); }
int harness_rule_nfactors(void* p, uint32_t rule) { return (int)((Harness*)p)->H.rules[rule].factors.size(); }
int harness_rule_exact(void* p, uint32_t rule) { return ((Harness*)p)->H.rules[rule].factors_exact ? 1 : 0; }

// all matches of one rule (global exec loop); out quadruples sb,eb,s16,e16; returns count or -1 on VM overflow
int harness_find_all(void* p, uint32_t rule, const uint8_t* m, uint32_t len, uint32_t* out, uint32_t cap) {
  Harness* h = (Harness*)p;
  VM local(h->d); VM* vm = &local;
  std::vector<uint32_t> v; VecSink sink{&v};
  run_rule<true>(*vm, h->d, rule, m, len, sink);
  if (vm->err) return -1;
  uint32_t k = (uint32_t)(v.size() / 4);
  for (uint32_t i = 0; i < k && i < cap; i++) memcpy(out + 4 * i, v.data() + 4 * i, 16);
  return (int)k;
}
int harness_test(void* p, uint32_t rule, const uint8_t* m, uint32_t len) {
  Harness* h = (Harness*)p;
  VM vm(h->d);
  std::vector<uint32_t> v; VecSink sink{&v};
  bool any = run_rule<false>(vm, h->d, rule, m, len, sink);
  if (vm.err) return -1;
  return any ? 1 : 0;
}

// the scan kernel's candidate logic (level-1 walk + level-2 confirm, same device/host code),
// restated for one message: bitmaps (rw words each) of queued candidates and of direct hits
struct BitSink {
  uint32_t* cand; uint32_t* direct_;
  std::vector<uint32_t>* occ;   // (rule, t0, pre) triples of confirmed occurrences
  void candidate(uint32_t r, uint32_t t0 = 0xffffffffu, uint32_t pre = 0xffffffffu) { cand[r >> 5] |= 1u << (r & 31); if (occ) { occ->push_back(r); occ->push_back(t0); occ->push_back(pre); } }
  void direct(uint32_t r) { direct_[r >> 5] |= 1u << (r & 31); }
};
void harness_candidates(void* p, const uint8_t* m, uint32_t len, uint32_t* cand, uint32_t* direct, int want_spans, uint32_t* l1_hits) {
  Harness* h = (Harness*)p; const DevRuleset& d = h->d;
  for (uint32_t k = 0; k < d.rw; k++) { cand[k] = 0; direct[k] = 0; }
  BitSink sink{cand, direct, nullptr};
  for (uint32_t k = 0; k < d.n_always; k++) sink.candidate(d.always_rules[k]);
  const uint8_t* lut = d.image + d.lut_off; const uint16_t* table = h->H.pf.table.data();
  uint32_t state = 0, hits = 0;
  if (d.mode == 4) {        // the fingerprint scan restated: rolling 4-symbol window, hash, 2-way bucket
    uint32_t win = 0;
    for (uint32_t i = 0; i < len; i++) {
      win = (win >> 8) | (fp_fold(m[i]) << 24);
      uint32_t hsh = win * d.fp_mult, b = (uint32_t)(((uint64_t)hsh * d.fp_buckets) >> 32), val = d.fp_table[b], fp = (hsh >> 8) & 0xffffu;
      if ((val & 0xffffu) == fp || (val >> 16) == fp) { hits++; if (getenv("CG_FP_DEBUG")) fprintf(stderr, "hit win %08x h %08x b %u val %08x fp %04x acc %d %d\n", win, hsh, b, val, fp, (int)d.fp_acc[2*b], (int)d.fp_acc[2*b+1]); fp_accept(d, hsh, m, len, i, want_spans != 0, sink); }
      for (uint32_t q = 0; q < d.n_trig; q++) if (m[i] == d.trig_byte[q]) { hits++; accept_id(d, d.trig_acc[q], m, len, i, want_spans != 0, sink); }
    }
    if (l1_hits) *l1_hits = hits;
    return;
  }
  for (uint32_t i = 0; i < len; i++) {
    uint32_t col = l1_col(d.mode, lut, m[i]);
    uint32_t ent = table[(state << d.ncols_log2) + col];
    if (ent & 0x8000u) { hits++; l1_accept(d, state, col, m, len, i, want_spans != 0, sink); }
    state = ent & 0x3fffu;
  }
  if (l1_hits) *l1_hits = hits;
}


// policy-mode verification exactly as verify_*_kernel does it: for every confirmed factor occurrence
// run test_at_factor; returns the bitmap of rules that hit (direct hits included)
void harness_policy_hits(void* p, const uint8_t* m, uint32_t len, uint32_t* hits) {
  Harness* h = (Harness*)p; const DevRuleset& d = h->d;
  std::vector<uint32_t> cand(d.rw, 0), occ;
  for (uint32_t k = 0; k < d.rw; k++) hits[k] = 0;
  BitSink sink{cand.data(), hits, &occ};
  for (uint32_t k = 0; k < d.n_always; k++) sink.candidate(d.always_rules[k]);
  const uint8_t* lut = d.image + d.lut_off; const uint16_t* table = h->H.pf.table.data();
  uint32_t state = 0;
  if (d.mode == 4) {
    uint32_t win = 0;
    for (uint32_t i = 0; i < len; i++) {
      win = (win >> 8) | (fp_fold(m[i]) << 24);
      uint32_t hsh = win * d.fp_mult, b = (uint32_t)(((uint64_t)hsh * d.fp_buckets) >> 32), val = d.fp_table[b], fp = (hsh >> 8) & 0xffffu;
      if ((val & 0xffffu) == fp || (val >> 16) == fp) fp_accept(d, hsh, m, len, i, false, sink);
      for (uint32_t q = 0; q < d.n_trig; q++) if (m[i] == d.trig_byte[q]) accept_id(d, d.trig_acc[q], m, len, i, false, sink);
    }
  } else
  for (uint32_t i = 0; i < len; i++) {
    uint32_t col = l1_col(d.mode, lut, m[i]);
    uint32_t ent = table[(state << d.ncols_log2) + col];
    if (ent & 0x8000u) l1_accept(d, state, col, m, len, i, false, sink);
    state = ent & 0x3fffu;
  }
  VM vm(d);
  struct Null { void span(uint32_t, uint32_t, uint32_t, uint32_t) {} } ns;
  for (size_t k = 0; k + 2 < occ.size(); k += 3) {
    uint32_t r = occ[k], t0 = occ[k + 1], pre = occ[k + 2];
    if ((hits[r >> 5] >> (r & 31)) & 1u) continue;
    bool any = t0 == 0xffffffffu ? run_rule<false>(vm, d, r, m, len, ns) : test_at_factor(vm, d, r, m, len, t0, pre);
    if (any) hits[r >> 5] |= 1u << (r & 31);
  }
}

// debug: level-1 accepting transitions per factor over one message (adds into counts[n_factors])
void harness_l1_factor_counts(void* p, const uint8_t* m, uint32_t len, uint32_t* counts) {
  Harness* h = (Harness*)p; const DevRuleset& d = h->d;
  const uint8_t* lut = d.image + d.lut_off; const uint16_t* table = h->H.pf.table.data();
  uint32_t state = 0;
  if (d.mode == 4) {
    uint32_t win = 0;
    for (uint32_t i = 0; i < len; i++) {
      win = (win >> 8) | (fp_fold(m[i]) << 24);
      uint32_t hsh = win * d.fp_mult, b = (uint32_t)(((uint64_t)hsh * d.fp_buckets) >> 32), val = d.fp_table[b], fp = (hsh >> 8) & 0xffffu;
      for (uint32_t way = 0; way < 2; way++) if (((val >> (16 * way)) & 0xffffu) == fp) { uint32_t aid = d.fp_acc[2 * b + way]; if (aid != 0xffffffffu) for (uint32_t k = d.acc_offsets[aid]; k < d.acc_offsets[aid + 1]; k++) counts[d.acc_factors[k]]++; }
    }
    return;
  }
  for (uint32_t i = 0; i < len; i++) {
    uint32_t col = l1_col(d.mode, lut, m[i]);
    uint32_t ent = table[(state << d.ncols_log2) + col];
    if (ent & 0x8000u) { uint32_t aid = d.acc_index[((size_t)state << d.ncols_log2) + col]; for (uint32_t k = d.acc_offsets[aid]; k < d.acc_offsets[aid + 1]; k++) counts[d.acc_factors[k]]++; }
    state = ent & 0x3fffu;
  }
}

// debug: print every full factor with its level-1 window
void harness_dump_factors(void* p) {
  Harness* h = (Harness*)p; const Prefilter& P = h->H.pf;
  for (size_t f = 0; f < P.factors.size(); f++) {
    const FullFactor& ff = P.factors[f];
    printf("rule %u len %d win [%d,%d) exact %d : ", ff.rule, ff.len, ff.win_off, ff.win_off + ff.win_len, ff.exact);
    for (int k = 0; k < ff.len; k++) {
      int cnt = 0, last = -1; for (int b = 0; b < 256; b++) if ((P.bytesets[(size_t)ff.elem[k] * 8 + (b >> 5)] >> (b & 31)) & 1) { cnt++; last = b; }
      if (cnt == 1) { if (last >= 33 && last < 127) printf("%c", last); else printf("\\x%02x", last); } else printf("[%d]", cnt);
    }
    printf("\n");
  }
}

}  // extern "C"

// level-1 state visit histogram over a batch of fixed-length messages (state resets per message):
// hist[s] += 1 for every transition taken out of state s; returns the number of accepting transitions
extern "C" uint64_t harness_l1_hist(void* p, const uint8_t* data, uint64_t n_msgs, uint32_t msg_len, uint64_t* hist) {
  Harness* h = (Harness*)p; const DevRuleset& d = h->d;
  const uint8_t* lut = d.image + d.lut_off; const uint16_t* table = h->H.pf.table.data();
  uint64_t acc = 0;
  for (uint64_t m = 0; m < n_msgs; m++) {
    uint32_t state = 0;
    const uint8_t* s = data + m * msg_len;
    for (uint32_t i = 0; i < msg_len; i++) {
      uint32_t col = l1_col(d.mode, lut, s[i]);
      uint32_t ent = table[(state << d.ncols_log2) + col];
      hist[state]++;
      if (ent & 0x8000u) acc++;
      state = ent & 0x3fffu;
    }
  }
  return acc;
}

// profile-guided residency on the host copy: renumber the states by `visits` (see rank_states_by_visits)
extern "C" void harness_rank(void* p, const uint32_t* visits) {
  Harness* h = (Harness*)p; HostImage& H = h->H; DevRuleset& d = h->d;
  rank_states_by_visits(&H, visits);
  d.image = H.image.data(); d.image_bytes = (uint32_t)H.image.size(); d.hot_states = H.hot_states; d.lut_off = H.lut_off; d.row_stride = H.row_stride;
  d.table_full = H.pf.table.data(); d.acc_index = H.pf.acc_index.data();
}
// the shared-memory image as the scan kernel sees it (tests check the trap-table invariants)
extern "C" const uint8_t* harness_image(void* p, uint32_t* out4) {
  Harness* h = (Harness*)p;
  out4[0] = h->d.image_bytes; out4[1] = h->d.hot_states; out4[2] = h->d.row_stride; out4[3] = h->d.lut_off;
  return h->d.image;
}
extern "C" const uint16_t* harness_table(void* p) { return ((Harness*)p)->H.pf.table.data(); }

// VM work statistics per rule over a batch (built with -DCG_VM_STATS only): for every confirmed occurrence that reaches
// the VM, accumulate the counters of pike_vm.h into out[rule * 8 + i], i = 0..4; out[rule * 8 + 5] = events, [6] = island start distance
#ifdef CG_VM_STATS
extern "C" unsigned long long cg_vm_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
extern "C" void harness_vm_stats(void* p, const uint8_t* data, uint64_t n_msgs, uint32_t msg_len, unsigned long long* out) {
  Harness* h = (Harness*)p; const DevRuleset& d = h->d;
  const uint8_t* lut = d.image + d.lut_off; const uint16_t* table = h->H.pf.table.data();
  VM vm(d);
  for (uint64_t mi = 0; mi < n_msgs; mi++) {
    const uint8_t* m = data + mi * msg_len; uint32_t len = msg_len;
    std::vector<uint32_t> cand(d.rw, 0), hits(d.rw, 0), occ;
    BitSink sink{cand.data(), hits.data(), &occ};
    uint32_t state = 0;
    for (uint32_t i = 0; i < len; i++) {
      uint32_t col = l1_col(d.mode, lut, m[i]);
      uint32_t ent = table[(state << d.ncols_log2) + col];
      if (ent & 0x8000u) l1_accept(d, state, col, m, len, i, false, sink);
      state = ent & 0x3fffu;
    }
    for (size_t k = 0; k + 2 < occ.size(); k += 3) {
      uint32_t r = occ[k], t0 = occ[k + 1], pre = occ[k + 2];
      if ((hits[r >> 5] >> (r & 31)) & 1u) continue;
      if (t0 == 0xffffffffu) continue;
      for (int i = 0; i < 8; i++) cg_vm_stats[i] = 0;
      bool any = test_at_factor(vm, d, r, m, len, t0, pre);
      if (any) hits[r >> 5] |= 1u << (r & 31);
      for (int i = 0; i < 5; i++) out[(size_t)r * 8 + i] += cg_vm_stats[i];
      out[(size_t)r * 8 + 5]++;
      out[(size_t)r * 8 + 6] += (pre & 0xffffu);
    }
  }
}
#endif
