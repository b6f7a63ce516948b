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
  std::vector<uint8_t> image;          // [lut 256][table], padded to 16 bytes
  std::vector<uint32_t> prog, prog_off, sets, first;
  std::vector<uint16_t> ranges;
  uint32_t n_sets = 0;
};

struct ImageOptions {
  int mode = 0;                 // prefilter column mode (0 direct7, 1 LUT)
  size_t budget_bytes = 192 * 1024;
  int max_classes = 64;
  int max_factor_len = kMaxFactorLen;
};

struct RuleSrc { const char* src; uint32_t len; uint32_t flags; };

// compiles every rule (failures stay in rules[i].status) and builds the prefilter + tables
bool build_host_image(const RuleSrc* rules, uint32_t n, const ImageOptions& opt, HostImage* out, std::string* err);

}  // namespace cg
