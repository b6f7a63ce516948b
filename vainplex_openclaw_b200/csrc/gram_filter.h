// gram_filter.h -- the data-independent half of the scan path: folding, the bitmap hash, the level-1b table lookup
// and the exact comparison of a factor at a text position.  CG_HD so that tests/native/vm_harness.cpp restates the
// scan kernel's candidate logic on the CPU with the very same code (the kernel's hot loop hand-inlines gram_fold_word
// and gram_bitmap_test; everything a flagged gram goes through afterwards is called from here).
//
// Replaces the inner loops of gov/src/conditions/context.ts:9-25 (rules x RegExp.test) and
// gov/src/redaction/registry.ts:212-242 (P global-exec passes per string) as far as *finding where to look* goes;
// the exact decision is the Pike VM's (pike_vm.h).
#pragma once
#include <cstdint>
#include "kernels.h"
#include "pike_vm.h"   // CG_HD macros

namespace cg {

// four bytes at once: z = (b & 0x5f) ^ 0x10 per byte, bytes with z < 0x10 (digits, : ; < = > ?) collapse to 0.
// z + 0x70 has bit 7 clear exactly for those bytes; PRMT in sign-replicate mode turns bit 7 into a byte mask.
CG_HD uint32_t gram_fold_word(uint32_t w) {
  const uint32_t z = (w & 0x5f5f5f5fu) ^ 0x10101010u;
#ifdef __CUDA_ARCH__
  uint32_t keep;       // prmt.b32, selector nibble bit 3 = replicate the selected byte's sign bit (__byte_perm masks that bit away)
  asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(keep) : "r"(z + 0x70707070u));
#else
  const uint32_t y = z + 0x70707070u; uint32_t keep = 0;
  for (int j = 0; j < 4; j++) if ((y >> (8 * j + 7)) & 1u) keep |= 0xffu << (8 * j);
#endif
  return z & (keep | 0xf0f0f0f0u);
}

// Bitmap geometry: word at byte address hi32(key * kGramMult) & bm_mask (bm_mask = bitmap bytes - 4, bitmap bytes a
// power of two); a key sets / needs bit 31 - (key & 31) -- the low five bits of its folded first byte -- and, for the
// optional second Bloom bit in the same word, bit 31 - ((hi32 >> 17) & 31).
CG_HD uint32_t gram_bitmap_addr(uint32_t key, uint32_t bm_mask) { return (uint32_t)(((uint64_t)key * kGramMult) >> 32) & bm_mask; }
CG_HD uint32_t gram_bitmap_bits(uint32_t key, bool bloom2) {
  const uint32_t hi = (uint32_t)(((uint64_t)key * kGramMult) >> 32);
  return (0x80000000u >> (key & 31u)) | (bloom2 ? (0x80000000u >> ((hi >> 17) & 31u)) : 0u);
}
CG_HD bool gram_bitmap_test(const uint8_t* bitmap, uint32_t key, uint32_t bm_mask, bool bloom2) {
  const uint32_t wv = *reinterpret_cast<const uint32_t*>(bitmap + gram_bitmap_addr(key, bm_mask));
  const uint32_t bits = gram_bitmap_bits(key, bloom2);
  return (wv & bits) == bits;
}

// The recheck map: a second, smaller bitmap under an independent hash (word = hi32(key * kGramMult2) & rk_mask, bit 31 - top
// five bits of the same product).  The hot loop never looks at it; a flagged gram is tested against it before the level-1b
// lookup, 32 flagged grams at a time, which sorts out the first bitmap's false positives for a handful of instructions each.
CG_HD uint32_t gram_recheck_hash(uint32_t key) { return (uint32_t)(((uint64_t)key * kGramMult2) >> 32); }
CG_HD bool gram_recheck_test(const uint8_t* rk, uint32_t key, uint32_t rk_mask) {
  const uint32_t h = gram_recheck_hash(key);
  return ((*reinterpret_cast<const uint32_t*>(rk + (h & rk_mask)) << (h >> 27)) & 0x80000000u) != 0;
}

// level-1b tables as the running side sees them (shared-memory copies inside the scan kernel when they fit, else HBM)
struct GramTables {
  const uint32_t* bucket_start;   // n_buckets + 1
  const uint2* entries;           // x = masked key, y = factor | (gram offset + 3) << 20 | shape << 25
  const uint32_t* factors;        // 12 words per factor (ruleset_image.cpp)
  const uint32_t* bytesets;       // 8 words per 256-bit set
};

CG_HD uint32_t gram_bucket(uint32_t masked_key, uint32_t shape, uint32_t nb_shift) { return ((masked_key ^ (shape * 0x9E3779B9u)) * kGramMult2) >> nb_shift; }

// one element of factor `fw` against the text byte at its position
CG_HD bool factor_elem_ok(const uint32_t* __restrict__ fw, const uint32_t* __restrict__ bytesets, const uint8_t* __restrict__ text, uint32_t k) {
  const uint32_t sid = (fw[2 + (k >> 1)] >> (16 * (k & 1))) & 0xffffu, b = text[k];
  return ((bytesets[(size_t)sid * 8 + (b >> 5)] >> (b & 31)) & 1u) != 0;
}
// exact comparison of factor `fw` with text[0, flen): every element's 256-bit byte set, first mismatch ends it
CG_HD bool factor_at(const uint32_t* __restrict__ fw, const uint32_t* __restrict__ bytesets, const uint8_t* __restrict__ text) {
  const uint32_t flen = fw[1] & 0xffu;
  for (uint32_t k = 0; k < flen; k++) if (!factor_elem_ok(fw, bytesets, text, k)) return false;
  return true;
}
// The same for a flagged gram that starts `goff` elements into the factor: the four elements FOLLOWING the gram
// (cyclically) first, all four loads in flight at once and no branch -- a flagged gram that is not an occurrence nearly
// always fails there, so the lanes of a draining warp stay together -- then the rest.
CG_HD bool factor_at_gram(const uint32_t* __restrict__ fw, const uint32_t* __restrict__ bytesets, const uint8_t* __restrict__ text, int goff) {
  const uint32_t flen = fw[1] & 0xffu;
  if (flen == 0) return true;
  uint32_t base = (uint32_t)(goff + kGramLen < 0 ? 0 : goff + kGramLen); if (base >= flen) base = 0;
  bool ok = true;
#pragma unroll
  for (uint32_t i = 0; i < 4; i++) {
    uint32_t k = base + i; if (k >= flen) k -= flen;
    if (i < flen) ok = ok & factor_elem_ok(fw, bytesets, text, k);
  }
  if (!ok) return false;
  for (uint32_t i = 4; i < flen; i++) { uint32_t k = base + i; if (k >= flen) k -= flen; if (!factor_elem_ok(fw, bytesets, text, k)) return false; }
  return true;
}

// A flagged gram: `key` = folded four bytes at buffer position `pos`.  Every registered (factor, gram offset) whose
// masked key equals the gram is compared exactly; emit(t0, factor) for each factor that really starts at t0, with
// [t0, t0 + len) inside [begin, end).
// Shaped for a draining warp (one flagged gram per lane): the walk over the shapes is uniform and loop-free per bucket
// (its first two entries), matches are only COLLECTED there; the factor comparison then runs once for every lane's first
// match and once more for second matches (guarded by warp votes on the device).  Buckets longer than two entries and
// third matches take the plain loop at the end.
template <class Emit>
CG_HD void gram_entry_check(const GramTables& T, uint32_t y, const uint8_t* __restrict__ buf, uint32_t begin, uint32_t end, uint32_t pos, Emit& emit) {
  const uint32_t f = y & 0xfffffu;
  const int goff = (int)((y >> 20) & 31u) - 3;
  const int64_t t0 = (int64_t)pos - goff;
  const uint32_t* fw = T.factors + (size_t)f * 12;
  if (t0 >= (int64_t)begin && t0 + (int64_t)(fw[1] & 0xffu) <= (int64_t)end && factor_at_gram(fw, T.bytesets, buf + t0, goff)) emit((uint32_t)t0, f);
}
template <class Emit>
CG_HD void gram_lookup(const DevRuleset& rs, const GramTables& T, uint32_t key, const uint8_t* __restrict__ buf, uint32_t begin, uint32_t end,
                       uint32_t pos, Emit& emit, bool active = true) {
  uint32_t ya = 0, yb = 0, nmatch = 0; bool spill = false;          // first two matches; spill: this lane has to take the plain loop
  for (uint32_t s = 0; s < rs.n_shapes; s++) {
    const uint32_t km = key & rs.shapes[s];
    const uint32_t b = gram_bucket(km, s, rs.nb_shift);
    const uint32_t e0 = T.bucket_start[b], e1 = active ? T.bucket_start[b + 1] : e0;
    if (e0 < e1) { const uint2 en = T.entries[e0]; if (en.x == km && (en.y >> 25) == s) { if (nmatch == 0) ya = en.y; else if (nmatch == 1) yb = en.y; nmatch++; } }
    if (e0 + 1 < e1) { const uint2 en = T.entries[e0 + 1]; if (en.x == km && (en.y >> 25) == s) { if (nmatch == 0) ya = en.y; else if (nmatch == 1) yb = en.y; nmatch++; } }
    spill = spill || e1 - e0 > 2;
  }
  spill = spill || nmatch > 2;
#ifdef __CUDA_ARCH__
  if (__any_sync(0xffffffffu, nmatch >= 1 && !spill)) { if (nmatch >= 1 && !spill) gram_entry_check(T, ya, buf, begin, end, pos, emit); }
  if (__any_sync(0xffffffffu, nmatch >= 2 && !spill)) { if (nmatch >= 2 && !spill) gram_entry_check(T, yb, buf, begin, end, pos, emit); }
  if (!__any_sync(0xffffffffu, spill)) return;
#else
  if (nmatch >= 1 && !spill) gram_entry_check(T, ya, buf, begin, end, pos, emit);
  if (nmatch >= 2 && !spill) gram_entry_check(T, yb, buf, begin, end, pos, emit);
#endif
  if (spill) for (uint32_t s = 0; s < rs.n_shapes; s++) {       // rare: every entry of every bucket
    const uint32_t km = key & rs.shapes[s];
    const uint32_t b = gram_bucket(km, s, rs.nb_shift);
    for (uint32_t e = T.bucket_start[b], e1 = T.bucket_start[b + 1]; e < e1; e++) {
      const uint2 en = T.entries[e];
      if (en.x == km && (en.y >> 25) == s) gram_entry_check(T, en.y, buf, begin, end, pos, emit);
    }
  }
}

// a trigger byte (trigger index ti) at buffer position pos: the factors it stands for, compared exactly
template <class Emit>
CG_HD void gram_trigger(const DevRuleset& rs, const GramTables& T, uint32_t ti, const uint8_t* __restrict__ buf, uint32_t begin, uint32_t end,
                        uint32_t pos, Emit& emit) {
  for (uint32_t k = rs.trig_offsets[ti], k1 = rs.trig_offsets[ti + 1]; k < k1; k++) {
    const uint32_t f = rs.trig_list[k] & 0xfffffu;
    const int64_t t0 = (int64_t)pos - (int64_t)(rs.trig_list[k] >> 20);
    const uint32_t* fw = T.factors + (size_t)f * 12;
    if (t0 < (int64_t)begin || t0 + (int64_t)(fw[1] & 0xffu) > (int64_t)end) continue;
    if (factor_at(fw, T.bytesets, buf + t0)) emit((uint32_t)t0, f);
  }
}

// message that owns buffer position pos (off[m] <= pos < off[m+1]); an interpolated first guess makes this one or two
// loads for batches of equal-length messages.  pos must lie in [off[0], off[n]).
CG_HD uint32_t message_of(const uint32_t* __restrict__ off, uint32_t n, uint32_t pos) {
  const uint32_t b0 = off[0], total = off[n] - b0;
  uint32_t lo = 0, hi = n;                     // invariant: off[lo] <= pos < off[hi]
  uint32_t g = (uint32_t)(((uint64_t)(pos - b0) * n) / (total ? total : 1u));
  if (g >= n) g = n - 1;
  if (off[g] <= pos) { lo = g; if (pos < off[g + 1]) return g; } else hi = g;
  while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (off[mid] <= pos) lo = mid; else hi = mid; }
  return lo;
}

// A confirmed factor occurrence inside message `msg`: direct hit when the factor IS the pattern (policy mode), else a
// candidate for the VM.  t0 = offset of the factor's first byte inside the message.
template <class Sink>
CG_HD void factor_confirmed(const DevRuleset& rs, uint32_t f, uint32_t t0, bool want_spans, Sink& sink) {
  const uint32_t* fw = rs.factors + (size_t)f * 12;
  if ((fw[1] >> 24) && !want_spans) sink.direct(fw[0]); else sink.candidate(fw[0], t0, fw[10] | (fw[11] << 16), f);   // max prefix units | prefix-alphabet set id << 16
}

}  // namespace cg
