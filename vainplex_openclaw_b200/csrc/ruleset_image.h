// ruleset_image.h -- host-side flat image of a compiled rule set (what gets uploaded to HBM).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "rulec.h"

namespace cg {

struct HostImage {
  std::vector<CompiledRule> rules;
  Prefilter pf;
  // what the scan kernel stages into shared memory: [bitmap][recheck map][bucket_start][entries]
  // (the level-1b tables only when they fit the budget: tables_resident)
  std::vector<uint8_t> image;
  uint32_t bm_bytes = 0, bm_mask = 0; bool bloom2 = false;
  uint32_t rk_off = 0, rk_bytes = 0;   // recheck map
  bool tables_resident = false;
  uint32_t dir_off = 0, ent_off = 0, nb_shift = 0, n_buckets = 0;
  std::vector<uint32_t> bucket_start;  // n_buckets + 1
  std::vector<uint32_t> entry_words;   // 2 words per level-1b entry, sorted by bucket: masked key, factor | (off + 3) << 20 | shape << 25
  // the same entries as lookup_kernel reads them: grouped by (masked key, shape); an open-addressing table of uint4 slots
  // {masked key, shape, first entry of the group, entries in the group (0 = empty slot)}, slot = hash >> slot_shift, linear probing
  std::vector<uint32_t> slot_words, group_entries; uint32_t slot_shift = 0, n_slots = 0;
  std::vector<uint64_t> bit_words; std::vector<uint32_t> bit_off;     // bitprog.h tables, 139 + 64 x rows words per eligible rule; bit_off[rule] = first word or 0xffffffff
  std::vector<uint64_t> factor_skip;   // 3 words per factor (bitprog.h: island_test)
  std::vector<uint32_t> prog, prog_off, sets, first, alpha;
  std::vector<uint32_t> factor_words;  // 12 words per full factor (device layout)
  std::vector<uint16_t> ranges;
  uint32_t n_sets = 0;
};

struct ImageOptions {
  int stride = 0;                       // 0 = choose (rulec.h)
  size_t budget_bytes = 185 * 1024;     // shared memory the image may take (scan_kernel adds 40 KB of rings + 1 KB; 227 KB per CTA)
  uint32_t bitmap_kb = 0;               // 0 = sized from the number of keys (16 .. 128 KB)
  int bloom2 = 0;                       // two bits per key in one bitmap word (costs three more ALU instructions per probe)
  uint32_t max_keys = 24576;
};

struct RuleSrc { const char* src; uint32_t len; uint32_t flags; };

// compiles every rule (failures stay in rules[i].status) and builds the prefilter + tables
bool build_host_image(const RuleSrc* rules, uint32_t n, const ImageOptions& opt, HostImage* out, std::string* err);

}  // namespace cg
