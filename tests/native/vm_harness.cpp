// tests/native/vm_harness.cpp -- TEST INFRASTRUCTURE ONLY (built into tests/native/_build/, never
// into the product library).  Runs the product's rule compiler, prefilter tables and the *same*
// Pike-VM source (csrc/pike_vm.h, compiled for the host here) on the CPU so that the CPU test tier
// (-m "not gpu") can compare them with the oracle before any GPU time is spent:
//   * compile status / program shape per rule
//   * prefilter soundness: every true (message, rule) hit must be a prefilter candidate
//   * exact spans of the VM == oracle/jsre.c spans
// The product path never links this file; libopenclaw_gov.so only runs the VM inside verify_kernel.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../vainplex_openclaw_b200/csrc/pike_vm.h"
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
  ImageOptions io; io.mode = mode; if (budget_kb) io.budget_bytes = (size_t)budget_kb * 1024; if (max_factor_len) io.max_factor_len = max_factor_len;
  std::string err;
  if (!build_host_image(src.data(), n, io, &h->H, &err)) { delete h; return nullptr; }
  for (uint32_t i = 0; i < n; i++) if (status) status[i] = h->H.rules[i].status;
  HostImage& H = h->H; DevRuleset& d = h->d;
  // pad host vectors the same way capi.cu pads the device copies
  H.sets.resize(H.sets.size() + 8, 0); H.ranges.resize(H.ranges.size() + 8, 0); H.first.resize(H.first.size() + 8, 0);
  H.pf.out_offsets.resize(H.pf.out_offsets.size() + 4, 0); H.pf.out_rules.resize(H.pf.out_rules.size() + 4, 0);
  H.pf.always_rules.resize(H.pf.always_rules.size() + 4, 0);
  d.image = H.image.data(); d.image_bytes = (uint32_t)H.image.size(); d.mode = (uint32_t)H.pf.mode;
  d.ncols_log2 = 0; while ((1 << d.ncols_log2) < H.pf.ncols) d.ncols_log2++;
  d.nstates = (uint32_t)H.pf.nstates; d.first_accept = (uint32_t)H.pf.first_accept;
  d.out_offsets = H.pf.out_offsets.data(); d.out_rules = H.pf.out_rules.data();
  d.always_rules = H.pf.always_rules.data(); d.n_always = (uint32_t)(H.pf.always_rules.size() - 4);
  d.prog = H.prog.data(); d.rule_prog_off = H.prog_off.data(); d.sets = H.sets.data(); d.set_ranges = H.ranges.data();
  d.rule_first = H.first.data(); d.n_rules = n; d.rw = (n + 31) / 32; if (!d.rw) d.rw = 1;
  return h;
}
void harness_destroy(void* p) { delete (Harness*)p; }

// info: [0]=nstates [1]=first_accept [2]=ncols [3]=factor_len [4]=n_always [5]=image_bytes [6]=program words
void harness_info(void* p, uint32_t* out) {
  Harness* h = (Harness*)p;
  out[0] = h->d.nstates; out[1] = h->d.first_accept; out[2] = (uint32_t)h->H.pf.ncols; out[3] = (uint32_t)h->H.pf.factor_len;
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

// the scan kernel's table walk, restated on the host: candidate bitmap (rw words) for one message
void harness_candidates(void* p, const uint8_t* m, uint32_t len, uint32_t* bits) {
  Harness* h = (Harness*)p; const DevRuleset& d = h->d;
  for (uint32_t k = 0; k < d.rw; k++) bits[k] = 0;
  for (uint32_t k = 0; k < d.n_always; k++) { uint32_t r = d.always_rules[k]; bits[r >> 5] |= 1u << (r & 31); }
  const uint8_t* lut = d.image; const uint16_t* table = (const uint16_t*)(d.image + 256);
  uint32_t state = 0;
  for (uint32_t i = 0; i < len; i++) {
    uint32_t col = d.mode == 0 ? (m[i] & 0x7fu) : lut[m[i]];
    state = table[(state << d.ncols_log2) + col];
    if (state >= d.first_accept) {
      uint32_t a = state - d.first_accept;
      for (uint32_t k = d.out_offsets[a]; k < d.out_offsets[a + 1]; k++) { uint32_t r = d.out_rules[k]; bits[r >> 5] |= 1u << (r & 31); }
    }
  }
}

}  // extern "C"
