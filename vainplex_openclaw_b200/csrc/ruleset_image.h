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
  std::vector<uint8_t> image;          // DFA modes: [rows 0..hot_states, row_stride bytes apart][lut 256]; mode 4: [lut 256][replicated buckets]
  uint32_t hot_states = 0;             // rows resident in shared memory (row `hot_states` itself is the trap row)
  uint32_t lut_off = 0, row_stride = 0;
  size_t budget_bytes = 0; uint32_t image_cols = 0;   // what build_dfa_image() was given (kept for rebuilds)
  std::vector<uint32_t> prog, prog_off, sets, first, alpha;
  std::vector<uint32_t> factor_words;  // 12 words per full factor (device layout)
  std::vector<uint16_t> ranges;
  uint32_t n_sets = 0;
};

struct ImageOptions {
  int mode = 2;                 // level-1 mode (0 direct7, 1 LUT, 2 folded 6-bit, 3 folded 5-bit DFA; 4 fingerprint table, falls back to 2)
  size_t budget_bytes = 155 * 1024;   // shared-memory budget of the image (scan_kernel adds 64 KB of staging buffers; 227 KB per CTA)
  int max_states = 16384;             // total states (cold rows are read from L2-resident HBM)
  int max_classes = 64;
  int max_window = kMaxWindow;
};

struct RuleSrc { const char* src; uint32_t len; uint32_t flags; };

// compiles every rule (failures stay in rules[i].status) and builds the prefilter + tables
bool build_host_image(const RuleSrc* rules, uint32_t n, const ImageOptions& opt, HostImage* out, std::string* err);

void build_dfa_image(HostImage* H);
// profile-guided residency: most visited level-1 states first (see ruleset_image.cpp)
void rank_states_by_visits(HostImage* H, const uint32_t* visits);

}  // namespace cg
