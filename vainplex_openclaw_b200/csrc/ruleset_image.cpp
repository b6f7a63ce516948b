// ruleset_image.cpp -- see ruleset_image.h
#include "ruleset_image.h"

#include <algorithm>
#include <cstring>

namespace cg {

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
  H.bit_off.assign(std::max<uint32_t>(n, 1), 0xffffffffu);
  for (uint32_t i = 0; i < n; i++) {
    std::vector<uint64_t> bp;
    if (H.rules[i].status == RULE_OK && build_bitprog(H.rules[i], &bp)) { H.bit_off[i] = (uint32_t)H.bit_words.size(); H.bit_words.insert(H.bit_words.end(), bp.begin(), bp.end()); }
  }
  H.n_sets = (uint32_t)(H.sets.size() / 6);

  PrefilterOptions po;
  po.stride = opt.stride; po.max_keys = opt.max_keys;
  if (!build_prefilter(H.rules, po, &H.pf, err)) return false;
  const Prefilter& P = H.pf;
  for (auto& f : P.factors) {
    H.factor_words.push_back(f.rule);
    H.factor_words.push_back((uint32_t)f.len | ((uint32_t)f.exact << 24));
    for (int k = 0; k < kMaxFactorElems; k += 2) H.factor_words.push_back((uint32_t)f.elem[k] | ((uint32_t)f.elem[k + 1] << 16));
    H.factor_words.push_back(f.pre); H.factor_words.push_back(f.pre_alpha);
  }
  // ---- what the island matcher may skip (bitprog.h: island_test): the rule's bit-parallel program run over the factor's own
  // elements, for factors that start the match in rules without assertions (one follow row, one start state)
  H.factor_skip.assign(3 * P.factors.size(), 0);
  for (size_t f = 0; f < P.factors.size(); f++) {
    const FullFactor& ff = P.factors[f];
    if (ff.pre != 0 || ff.exact || H.bit_off[ff.rule] == 0xffffffffu) continue;
    const uint64_t* bp = H.bit_words.data() + H.bit_off[ff.rule];
    const uint32_t W = (uint32_t)bp[0]; const uint64_t* t = bp + 1;
    const uint64_t meta = t[(128 + 32) * W];
    if ((meta & 15u) != 1u || !((meta >> 8) & 1u) || ((meta >> 9) & 1u)) continue;          // assertions / lookaround: the state depends on the text
    const uint64_t* rows = t + (128 + 32) * W + 7 + 2 * W;
    uint64_t live[2] = {t[128 * W], W == 2 ? t[128 * W + 1] : 0};
    uint32_t L = 0; uint64_t keep[2] = {0, 0};
    for (uint32_t k = 0; k < ff.len; k++) {
      const uint32_t* bs = P.bytesets.data() + (size_t)ff.elem[k] * 8;
      if (bs[4] | bs[5] | bs[6] | bs[7]) break;                                                // bytes >= 0x80: units, not bytes
      if (!(bs[0] | bs[1] | bs[2] | bs[3])) break;
      uint64_t all[2] = {~0ull, ~0ull}, some[2] = {0, 0};
      for (int b = 0; b < 128; b++) if ((bs[b >> 5] >> (b & 31)) & 1u) for (uint32_t w = 0; w < W; w++) { all[w] &= t[(size_t)b * W + w]; some[w] |= t[(size_t)b * W + w]; }
      bool ambiguous = false; uint64_t hit[2] = {0, 0};
      for (uint32_t w = 0; w < W; w++) { const uint64_t lw = w == W - 1 ? live[w] & ~(1ull << 63) : live[w]; hit[w] = lw & all[w]; if (lw & some[w] & ~all[w]) ambiguous = true; }
      if (ambiguous) break;
      uint64_t nl[2] = {0, 0};
      for (uint32_t w = 0; w < W; w++) for (uint64_t hw = hit[w]; hw; hw &= hw - 1) { const int j = __builtin_ctzll(hw); for (uint32_t v = 0; v < W; v++) nl[v] |= rows[(size_t)(w * 64 + j) * W + v]; }
      live[0] = nl[0]; live[1] = nl[1]; L = k + 1; keep[0] = live[0]; keep[1] = live[1];
      if (!(live[0] | live[1])) break;
    }
    H.factor_skip[3 * f] = L; H.factor_skip[3 * f + 1] = keep[0]; H.factor_skip[3 * f + 2] = keep[1];
  }
  // ---- level-1b hash table: entries sorted by bucket
  uint32_t nb = 16; while (nb < P.entries.size() * 2) nb *= 2;      // sparse: the lookup reads a bucket's first two entries without a loop
  H.n_buckets = nb; H.nb_shift = 32; { uint32_t t = nb; while (t > 1) { t >>= 1; H.nb_shift--; } }
  std::vector<uint32_t> shape_of(P.entries.size()), bucket_of(P.entries.size()), order(P.entries.size());
  for (size_t i = 0; i < P.entries.size(); i++) {
    uint32_t s = 0; while (P.shapes[s] != P.entries[i].mask) s++;
    shape_of[i] = s; bucket_of[i] = (uint32_t)((((P.entries[i].key & P.entries[i].mask) ^ (s * 0x9E3779B9u)) * kGramMult2) >> H.nb_shift); order[i] = (uint32_t)i;
  }
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return bucket_of[a] < bucket_of[b]; });
  H.bucket_start.assign((size_t)nb + 1, 0);
  for (uint32_t i : order) H.bucket_start[bucket_of[i] + 1]++;
  for (uint32_t b = 0; b < nb; b++) H.bucket_start[b + 1] += H.bucket_start[b];
  for (uint32_t i : order) {
    const GramEntry& e = P.entries[i];
    H.entry_words.push_back(e.key & e.mask);
    H.entry_words.push_back(e.factor | ((uint32_t)(e.off + 3) << 20) | (shape_of[i] << 25));
  }
  // ---- the same entries grouped by (masked key, shape) behind an open-addressing table (lookup_kernel)
  {
    std::vector<uint32_t> idx(P.entries.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = (uint32_t)i;
    auto kof = [&](uint32_t i) { return std::make_pair(shape_of[i], P.entries[i].key & P.entries[i].mask); };
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return kof(a) < kof(b); });
    uint32_t ns = 16; while (ns < idx.size() * 2) ns *= 2;
    H.n_slots = ns; H.slot_shift = 32; { uint32_t t = ns; while (t > 1) { t >>= 1; H.slot_shift--; } }
    H.slot_words.assign((size_t)ns * 4, 0);
    for (size_t i = 0; i < idx.size();) {
      size_t j = i; while (j < idx.size() && kof(idx[j]) == kof(idx[i])) j++;
      const uint32_t s = shape_of[idx[i]], km = P.entries[idx[i]].key & P.entries[idx[i]].mask;
      uint32_t slot = (uint32_t)(((km ^ (s * 0x9E3779B9u)) * kGramMult2) >> H.slot_shift);
      while (H.slot_words[(size_t)slot * 4 + 3]) slot = (slot + 1) & (ns - 1);
      H.slot_words[(size_t)slot * 4] = km; H.slot_words[(size_t)slot * 4 + 1] = s;
      H.slot_words[(size_t)slot * 4 + 2] = (uint32_t)H.group_entries.size(); H.slot_words[(size_t)slot * 4 + 3] = (uint32_t)(j - i);
      for (size_t k = i; k < j; k++) { const GramEntry& e = P.entries[idx[k]]; H.group_entries.push_back(e.factor | ((uint32_t)(e.off + 3) << 20) | (s << 25)); }
      i = j;
    }
  }
  // ---- bitmap: one bit per key (optionally two in the same word: blocked Bloom filter).  Every false positive costs a pass
  // through stage 1 of the slow path, staging the image costs next to nothing: 128 KB unless the level-1b tables then no
  // longer fit beside it and the rule set is small enough for a denser bitmap (>= 128 bits per key)
  uint32_t bm = opt.bitmap_kb ? opt.bitmap_kb * 1024u : 128u * 1024u;
  { uint32_t t = 4096; while (t < bm) t *= 2; bm = std::min<uint32_t>(t, 128u * 1024u); }       // power of two
  auto a16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  // recheck map: 32 bits per key (about 3 % of the first bitmap's false positives survive it), 4 .. 16 KB (confirm_kernel stages it into its own shared memory)
  uint32_t rk = 4096; while (rk < 16384u && (uint64_t)rk * 8 < (uint64_t)P.keys.size() * 32) rk *= 2;
  const size_t tables0 = a16(H.bucket_start.size() * 4) + a16(H.entry_words.size() * 4);     // (factor words and byte sets are read from HBM / L2: only the last, warp-parallel stage of the slow path needs them)
  const size_t tables = tables0 + rk;
  while (bm > 16u * 1024u && (size_t)bm + tables > opt.budget_bytes && (uint64_t)(bm / 2) * 8 >= (uint64_t)P.keys.size() * 128) bm /= 2;   // tables resident beats a sparser bitmap
  H.tables_resident = false;      // (scan_kernel stages the bitmap alone; the tables are confirm_kernel's, read through L1 / L2)
  if (!H.tables_resident && !opt.bitmap_kb) bm = 128u * 1024u;
  H.bm_bytes = bm; H.bm_mask = bm - 4; H.bloom2 = opt.bloom2 != 0;
  H.image.assign(bm, 0);
  for (uint32_t key : P.keys) {
    const uint32_t hi = (uint32_t)(((uint64_t)key * kGramMult) >> 32), addr = hi & H.bm_mask;
    uint32_t wv; memcpy(&wv, H.image.data() + addr, 4);
    wv |= (0x80000000u >> (key & 31u)) | (H.bloom2 ? (0x80000000u >> ((hi >> 17) & 31u)) : 0u);
    memcpy(H.image.data() + addr, &wv, 4);
  }
  auto append = [&](const void* p, size_t bytes, size_t pad) { uint32_t o = (uint32_t)H.image.size(); H.image.resize(a16(H.image.size() + bytes + pad), 0); if (bytes) memcpy(H.image.data() + o, p, bytes); return o; };
  {
    std::vector<uint8_t> map(rk, 0);
    for (uint32_t key : P.keys) {
      const uint32_t h = (uint32_t)(((uint64_t)key * kGramMult2) >> 32), addr = h & (rk - 4);
      uint32_t wv; memcpy(&wv, map.data() + addr, 4); wv |= 0x80000000u >> (h >> 27); memcpy(map.data() + addr, &wv, 4);
    }
    H.rk_bytes = rk; H.rk_off = append(map.data(), map.size(), 0);
  }
  if (H.tables_resident) {
    H.dir_off = append(H.bucket_start.data(), H.bucket_start.size() * 4, 0);
    H.ent_off = append(H.entry_words.data(), H.entry_words.size() * 4, 0);
  }
  return true;
}

}  // namespace cg
