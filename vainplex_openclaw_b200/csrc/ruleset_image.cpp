// ruleset_image.cpp -- see ruleset_image.h
#include "ruleset_image.h"

#include <algorithm>
#include <cstring>

namespace cg {

// (Re)builds the shared-memory image of a DFA-mode rule set from H.pf.table: see the layout comment below.
void build_dfa_image(HostImage* out) {
  // shared-memory image: [rows 0..hot of the table, `row_stride` bytes apart][pad to 16][lut 256 B].
  // Entries are plain u16 state indices.  Every transition that is accepting, or leads to a state that is
  // not resident (>= hot), is stored as `hot`, the index of an absorbing trap row: the fast path of
  // scan_kernel then needs no flag bits -- a chunk that ends in the trap row is re-walked on the full table.
  // row_stride = 2 * ncols + 4: the 4 pad bytes rotate consecutive rows by one bank, so lanes sitting in
  // different states but reading the same frequent column (' ', 'e', ...) land in different banks.
  HostImage& H = *out; size_t budget = H.budget_bytes; const uint32_t cols = H.image_cols; size_t hot_rows;
  H.row_stride = cols * 2 + 4;
  budget = std::min<size_t>(budget, (size_t)155 * 1024);            // + 64 KB staging + 6 KB event buffers + static shared memory <= 227 KB per CTA
  hot_rows = std::min<size_t>((budget - 256 - 16) / H.row_stride, 16383);
  H.hot_states = (uint32_t)std::min<size_t>(hot_rows > 1 ? hot_rows - 1 : 1, (size_t)H.pf.nstates);
  const uint32_t hot = H.hot_states; const size_t nc = (size_t)H.pf.ncols;
  size_t tbl_bytes = ((size_t)(hot + 1) * H.row_stride + 15) & ~(size_t)15;
  H.lut_off = (uint32_t)tbl_bytes;
  H.image.assign(tbl_bytes + 256, 0);
  auto put = [&](uint32_t row, size_t col, uint16_t v) { memcpy(H.image.data() + (size_t)row * H.row_stride + 2 * col, &v, 2); };
  for (uint32_t s = 0; s < hot; s++) for (size_t c = 0; c < nc; c++) {
    uint16_t e = H.pf.table[(size_t)s * nc + c];
    put(s, c, ((e & 0x8000) || (uint32_t)(e & 0x3fff) >= hot) ? (uint16_t)hot : (uint16_t)(e & 0x3fff));
  }
  for (size_t c = 0; c < (size_t)cols; c++) put(hot, c, (uint16_t)hot);
  memcpy(H.image.data() + H.lut_off, H.pf.lut, 256);
}

// Profile-guided residency: renumber the level-1 states so that the most visited ones get the lowest indices
// (= the rows that are resident in shared memory), keeping the start state at index 0.  `visits[s]` is how often
// state s was left on a sample of real traffic.  Everything indexed by state (table, acc_index) is permuted and
// the image rebuilt; the matcher's results do not depend on the numbering, only how often its slow path runs.
void rank_states_by_visits(HostImage* out, const uint32_t* visits) {
  HostImage& H = *out; Prefilter& P = H.pf;
  if (P.mode == 4 || P.nstates <= 1) return;
  const uint32_t ns = (uint32_t)P.nstates; const size_t nc = (size_t)P.ncols;
  std::vector<uint32_t> order(ns);
  for (uint32_t i = 0; i < ns; i++) order[i] = i;
  std::stable_sort(order.begin() + 1, order.end(), [&](uint32_t a, uint32_t b) { return visits[a] > visits[b]; });
  std::vector<uint32_t> new_of_old(ns);
  for (uint32_t i = 0; i < ns; i++) new_of_old[order[i]] = i;
  std::vector<uint16_t> table(P.table.size()); std::vector<uint32_t> acc(P.acc_index.size(), 0xffffffffu);
  for (uint32_t i = 0; i < ns; i++) {
    const uint32_t o = order[i];
    for (size_t c = 0; c < nc; c++) {
      const uint16_t e = P.table[(size_t)o * nc + c];
      table[(size_t)i * nc + c] = (uint16_t)((e & 0xc000) | new_of_old[e & 0x3fff]);
      acc[(size_t)i * nc + c] = P.acc_index[(size_t)o * nc + c];
    }
  }
  P.table.swap(table); P.acc_index.swap(acc);
  build_dfa_image(out);
}

bool build_host_image(const RuleSrc* src, uint32_t n, const ImageOptions& opt, HostImage* out, std::string* err) {
  HostImage& H = *out;
  H = HostImage();
  H.rules.resize(n);
  for (uint32_t i = 0; i < n; i++) H.rules[i] = compile_rule(src[i].src ? src[i].src : "", src[i].len, src[i].flags);
  H.prog_off.assign((size_t)n + 1, 0);
  H.first.assign(8 * (size_t)std::max<uint32_t>(n, 1), 0);
  H.alpha.assign(8 * (size_t)std::max<uint32_t>(n, 1), 0);
  for (uint32_t i = 0; i < n; i++) {
    H.prog_off[i] = (uint32_t)H.prog.size();
    CompiledRule& r = H.rules[i];
    if (r.status != RULE_OK) continue;     // empty program: never matches, no factors
    uint32_t base = (uint32_t)(H.sets.size() / 6);
    for (auto& s : r.sets) {
      for (int k = 0; k < 4; k++) H.sets.push_back(s.ascii[k]);
      H.sets.push_back((uint32_t)H.ranges.size());            // element offset of the first lo,hi pair
      H.sets.push_back((uint32_t)(s.ranges.size() / 2));
      H.ranges.insert(H.ranges.end(), s.ranges.begin(), s.ranges.end());
    }
    for (uint32_t ins : r.prog) {
      uint32_t op = ins & 0xff, arg = ins >> 8;
      if (op == OP_SET || op == OP_LOOKAHEAD || op == OP_NLOOKAHEAD || op == OP_LOOKBEHIND || op == OP_NLOOKBEHIND) arg += base;
      H.prog.push_back(op | (arg << 8));
    }
    for (int k = 0; k < 8; k++) H.first[(size_t)i * 8 + k] = (uint32_t)(r.first_bytes.w[k >> 1] >> (32 * (k & 1)));
    for (int k = 0; k < 8; k++) H.alpha[(size_t)i * 8 + k] = (uint32_t)(r.alphabet.w[k >> 1] >> (32 * (k & 1)));
  }
  H.prog_off[n] = (uint32_t)H.prog.size();
  H.n_sets = (uint32_t)(H.sets.size() / 6);

  PrefilterOptions po;
  po.mode = opt.mode; po.max_classes = opt.max_classes; po.max_window = opt.max_window;
  int cols = po.mode == 0 ? 128 : po.mode == 2 ? 64 : po.mode == 3 ? 32 : po.mode == 4 ? 64 : (po.max_classes <= 32 ? 32 : 64);
  size_t budget = std::max<size_t>(opt.budget_bytes, 256 + (size_t)cols * 2 * 2);
  size_t hot_rows = std::min<size_t>((budget - 256) / ((size_t)cols * 2), 32767);
  po.max_states = std::max<int>((int)std::min<size_t>((size_t)std::max(opt.max_states, 1), 16383), 1);
  po.fp_buckets = (int)std::max<size_t>((budget - 256) / 128, 16);
  if (!build_prefilter(H.rules, po, &H.pf, err)) return false;
  for (auto& f : H.pf.factors) {
    H.factor_words.push_back(f.rule);
    H.factor_words.push_back((uint32_t)f.len | ((uint32_t)f.win_off << 8) | ((uint32_t)f.win_len << 16) | ((uint32_t)f.exact << 24));
    for (int k = 0; k < kMaxFactorElems; k += 2) H.factor_words.push_back((uint32_t)f.elem[k] | ((uint32_t)f.elem[k + 1] << 16));
    H.factor_words.push_back(f.pre); H.factor_words.push_back(f.pre_alpha);
  }
  if (H.pf.mode == 4) {
    // fingerprint table replicated once per bank: word (bucket * 32 + lane) holds bucket's two fingerprints
    H.hot_states = 0; H.lut_off = 0; H.row_stride = 0;
    H.image.assign(256 + (size_t)H.pf.fp_buckets * 128, 0);
    memcpy(H.image.data(), H.pf.lut, 256);
    uint32_t* t = reinterpret_cast<uint32_t*>(H.image.data() + 256);
    for (uint32_t b = 0; b < H.pf.fp_buckets; b++) for (int l = 0; l < 32; l++) t[(size_t)b * 32 + l] = H.pf.fp_table[b];
    while (H.image.size() % 16) H.image.push_back(0);
    return true;
  }
  H.budget_bytes = budget; H.image_cols = (uint32_t)cols;
  build_dfa_image(&H);
  return true;
}

}  // namespace cg
