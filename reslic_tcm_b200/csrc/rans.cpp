// rans.cpp — host-side rANS coder with CDF indexes.
//
// Native replacement for the un-vendored third-party extension `compressai.ans`
// (RansEncoder / BufferedRansEncoder / RansDecoder; reference call sites
// src/models/reference/tcm.py:522,564-565,604-605,621 and
// src/entropy_models/adaptive_gaussian_conditional.py:291-299,711-721).  The algorithm follows the
// published CompressAI 1.2.x coder as recalled (ryg_rans 64-bit state emitting 32-bit words,
// 16-bit probability precision, out-of-range symbols escaped through 4-bit bypass groups;
// SURVEY.md App. A.6).  compressai is not installed here, so byte-compatibility with its
// bitstreams is NOT verified; what is tested is exact encode -> decode round trip.
// The coder is sequential CPU code by nature; the GPU path produces its inputs (int32 symbols and
// CDF indexes, or the packed (start, range) words themselves: rans_slots.cu) in one fused pass.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>
#include "reslic_internal.h"

namespace {

constexpr int kPrecision = 16;
constexpr int kBypassPrecision = 4;
constexpr int32_t kMaxBypassVal = (1 << kBypassPrecision) - 1;
constexpr uint64_t kRansL = 1ull << 31;

// One 32-bit word per coded symbol, the format reslic_rans_slots_u32 emits: start << 16 | range.  range == 0 marks a
// 4-bit bypass group whose value sits in the upper half (a table symbol never has range 0), so the slots a GPU launch
// produced are appended with memcpy and the coding loop reads 4 bytes per symbol.
inline uint32_t sym_word(uint32_t start, uint32_t range) { return start << 16 | range; }
inline uint32_t bypass_word(uint32_t val) { return val << 16; }

struct Encoder { std::vector<uint32_t> syms; std::vector<uint32_t> out; const uint32_t* begin = nullptr; int64_t nbytes = 0; };
struct Decoder { std::vector<uint32_t> words; size_t pos = 0; uint64_t state = 0; std::vector<int32_t> lut; };

// x / freq without a divide (Alverson, "Integer division using reciprocals", as ryg_rans' RansEncSymbol does): for
// 2^(k-1) < freq <= 2^k, rcp = ceil(2^(k+63) / freq) fits 64 bits and floor(x * rcp / 2^(64+k-1)) == floor(x / freq) for
// every x < 2^63 — the error term x * e / (freq * 2^(63+k)) with e < freq stays below 2^-k <= 1/freq, too small to
// carry a fractional part of at most 1 - 1/freq over.  The state is below x_max <= 2^63 when it is divided.  freq == 1
// uses rcp = 2^64 - 1, which yields x - 1, and a bias that makes up for it.  One 16-byte entry per possible frequency
// (1 MB, built on first use; a stream touches the few hundred entries its tables contain).
struct Rcp { uint64_t rcp; uint32_t shift; uint32_t bias; };
const Rcp* rcp_table() {
  static const std::vector<Rcp> table = [] {
    std::vector<Rcp> v(1u << kPrecision, Rcp{0, 0, 0});
    v[1] = Rcp{~0ull, 0u, (1u << kPrecision) - 1u};
    for (uint32_t f = 2; f < (1u << kPrecision); ++f) {
      uint32_t k = 0;
      while ((1u << k) < f) ++k;                                   // ceil(log2 f) >= 1
      const unsigned __int128 num = (static_cast<unsigned __int128>(1) << (k + 63)) + (f - 1);
      v[f] = Rcp{static_cast<uint64_t>(num / f), k - 1, 0u};
    }
    return v;
  }();
  return table.data();
}
// x' = (x / freq) * 2^16 + x % freq + start  =  x + start + (x / freq) * (2^16 - freq)
inline void enc_put(uint64_t& x, uint32_t*& ptr, uint32_t start, uint32_t freq, const Rcp* rcp) {
  const uint64_t x_max = ((kRansL >> kPrecision) << 32) * freq;
  if (x >= x_max) { *--ptr = static_cast<uint32_t>(x); x >>= 32; }      // (a branch-free form measured the same)
  const Rcp r = rcp[freq];
  const uint64_t q = static_cast<uint64_t>((static_cast<unsigned __int128>(x) * r.rcp) >> 64) >> r.shift;
  x = x + start + r.bias + q * ((1u << kPrecision) - freq);
}
// the plain form, kept for the self-check below
inline uint64_t enc_step_div(uint64_t x, uint32_t start, uint32_t freq) {
  return ((x / freq) << kPrecision) + (x % freq) + start;
}
inline void enc_put_bits(uint64_t& x, uint32_t*& ptr, uint32_t val, int nbits) {
  const uint64_t freq = 1ull << (16 - nbits);
  const uint64_t x_max = ((kRansL >> 16) << 32) * freq;
  if (x >= x_max) { *--ptr = static_cast<uint32_t>(x); x >>= 32; }
  x = (x << nbits) | val;
}

// escape: the number of 4-bit groups (itself in groups of at most 15), then the groups, low bits first
inline void push_escape(std::vector<uint32_t>& syms, uint64_t raw) {
  int32_t n_bypass = 0;
  while ((raw >> (n_bypass * kBypassPrecision)) != 0) ++n_bypass;
  int32_t val = n_bypass;
  while (val >= kMaxBypassVal) {
    syms.push_back(bypass_word(static_cast<uint32_t>(kMaxBypassVal)));
    val -= kMaxBypassVal;
  }
  syms.push_back(bypass_word(static_cast<uint32_t>(val)));
  for (int32_t j = 0; j < n_bypass; ++j)
    syms.push_back(bypass_word(static_cast<uint32_t>((raw >> (j * kBypassPrecision)) & kMaxBypassVal)));
}

struct Tables { const int32_t* cdfs; int32_t n_cdfs, stride; const int32_t* sizes; const int32_t* offsets; };

int check_tables(const Tables& t) {
  if (!t.cdfs || !t.sizes || !t.offsets || t.n_cdfs < 1 || t.stride < 3)
    return reslic::set_error(RESLIC_ERR_ARG, "rans: CDF tables missing");
  for (int i = 0; i < t.n_cdfs; ++i)
    if (t.sizes[i] < 3 || t.sizes[i] > t.stride)
      return reslic::set_error(RESLIC_ERR_ARG, "rans: a cdf length is outside 3..stride");
  return RESLIC_OK;
}

}  // namespace

extern "C" {

// Self-check of the reciprocal table against the divide it replaces: for EVERY frequency 1..65535 the states at the
// edges of the admissible range ([2^31, 2^47 * freq): first, last, multiples of freq and their neighbours) and
// `samples_per_freq` pseudo-random states in between.  Returns the number of mismatches (0 = exact).
int64_t reslic_rans_check_reciprocals(int64_t samples_per_freq, uint64_t seed) {
  const Rcp* rcp = rcp_table();
  int64_t bad = 0;
  uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
  auto next = [&s]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (uint32_t f = 1; f < (1u << kPrecision); ++f) {
    const uint64_t lo = kRansL, hi = ((kRansL >> kPrecision) << 32) * f;      // x in [lo, hi) when it is divided
    auto check = [&](uint64_t x) {
      if (x < lo || x >= hi) return;
      uint64_t y = x;
      uint32_t* none = nullptr;
      enc_put(y, none, 7u, f, rcp);                      // x < x_max: no renormalisation, the pointer is not touched
      if (y != enc_step_div(x, 7u, f)) ++bad;
    };
    check(lo); check(lo + 1); check(hi - 1); check(hi - 2);
    const uint64_t m = (hi - 1) / f * f, m0 = (lo + f - 1) / f * f;
    check(m); check(m - 1); check(m + 1); check(m0); check(m0 - 1); check(m0 + 1);
    for (int64_t i = 0; i < samples_per_freq; ++i) {
      const uint64_t x = lo + next() % (hi - lo);
      check(x);
      check(x / f * f); check(x / f * f + f - 1);       // a multiple of freq and the state just below the next one
    }
  }
  return bad;
}

void* reslic_rans_encoder_create(void) { return new (std::nothrow) Encoder(); }
void reslic_rans_encoder_destroy(void* h) { delete static_cast<Encoder*>(h); }

int reslic_rans_encoder_push(void* h, const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                             int32_t n_cdfs, int32_t cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets) {
  auto* e = static_cast<Encoder*>(h);
  if (!e || n < 0 || (n > 0 && (!symbols || !indexes))) return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push: bad argument");
  const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
  if (int rc = check_tables(t)) return rc;
  e->syms.reserve(e->syms.size() + static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    if (ci < 0 || ci >= n_cdfs) return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push: cdf index out of range");
    const int32_t* cdf = cdfs + static_cast<int64_t>(ci) * cdf_stride;
    const int32_t max_value = cdf_sizes[ci] - 2;
    int64_t value = static_cast<int64_t>(symbols[i]) - offsets[ci];
    uint64_t raw = 0;
    if (value < 0) { raw = static_cast<uint64_t>(-2 * value - 1); value = max_value; }
    else if (value >= max_value) { raw = static_cast<uint64_t>(2 * (value - max_value)); value = max_value; }
    const int32_t start = cdf[value], range = cdf[value + 1] - cdf[value];
    if (range <= 0 || start < 0 || start + range > (1 << kPrecision))
      return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push: invalid cdf (zero or negative frequency)");
    if (range > 0xffff) return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push: a symbol with probability 1 (range 65536) cannot be coded");
    e->syms.push_back(sym_word(static_cast<uint32_t>(start), static_cast<uint32_t>(range)));
    if (value == max_value) push_escape(e->syms, raw);
  }
  return RESLIC_OK;
}

// The same with the table lookups already done on the device (rans_slots.cu): slots[i] = start << 16 | range,
// escapes (position ascending, raw bypass value) for the symbols that took their table's escape slot.
int reslic_rans_encoder_push_slots(void* h, const uint32_t* slots, int64_t n, const int32_t* esc_pos,
                                   const int64_t* esc_raw, int64_t n_esc) {
  auto* e = static_cast<Encoder*>(h);
  if (!e || n < 0 || n_esc < 0 || (n > 0 && !slots) || (n_esc > 0 && (!esc_pos || !esc_raw)))
    return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push_slots: bad argument");
  uint32_t any_empty = 0;
  for (int64_t i = 0; i < n; ++i) any_empty |= static_cast<uint32_t>((slots[i] & 0xffffu) == 0u);      // vectorises
  if (any_empty) return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push_slots: empty slot (zero frequency)");
  int64_t prev = -1;
  for (int64_t k = 0; k < n_esc; ++k) {
    if (esc_pos[k] <= prev || esc_pos[k] >= n)
      return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push_slots: escape positions not ascending or out of range");
    if (esc_raw[k] < 0) return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push_slots: negative bypass value");
    prev = esc_pos[k];
  }
  e->syms.reserve(e->syms.size() + static_cast<size_t>(n) + static_cast<size_t>(n_esc) * 4);
  // the slots ARE the coder's words: copied in runs, the bypass groups of an escape spliced in behind its symbol
  int64_t pos = 0;
  for (int64_t k = 0; k < n_esc; ++k) {
    const int64_t upto = static_cast<int64_t>(esc_pos[k]) + 1;
    e->syms.insert(e->syms.end(), slots + pos, slots + upto);
    push_escape(e->syms, static_cast<uint64_t>(esc_raw[k]));
    pos = upto;
  }
  e->syms.insert(e->syms.end(), slots + pos, slots + n);
  return RESLIC_OK;
}

// Encodes everything pushed so far (in reverse, as rANS requires) and returns the byte count;
// the bytes stay owned by the encoder until the next push/flush/destroy.
int64_t reslic_rans_encoder_flush(void* h, const uint8_t** data) {
  auto* e = static_cast<Encoder*>(h);
  if (!e || !data) { reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_flush: bad argument"); return -1; }
  e->out.assign(e->syms.size() + 4, 0u);
  uint32_t* end = e->out.data() + e->out.size();
  uint32_t* ptr = end;
  uint64_t x = kRansL;
  const Rcp* rcp = rcp_table();
  const uint32_t* syms = e->syms.data();
  for (size_t k = e->syms.size(); k-- > 0;) {
    const uint32_t w = syms[k], range = w & 0xffffu;
    if (range) enc_put(x, ptr, w >> 16, range, rcp);
    else enc_put_bits(x, ptr, w >> 16, kBypassPrecision);
  }
  ptr -= 2;
  ptr[0] = static_cast<uint32_t>(x);
  ptr[1] = static_cast<uint32_t>(x >> 32);
  e->syms.clear();
  e->begin = ptr;
  e->nbytes = static_cast<int64_t>(end - ptr) * static_cast<int64_t>(sizeof(uint32_t));
  *data = reinterpret_cast<const uint8_t*>(ptr);
  return e->nbytes;
}

void* reslic_rans_decoder_create(const uint8_t* data, int64_t nbytes) {
  if (!data || nbytes < 8 || (nbytes % 4) != 0) { reslic::set_error(RESLIC_ERR_ARG, "rans_decoder_create: bad stream"); return nullptr; }
  auto* d = new (std::nothrow) Decoder();
  if (!d) return nullptr;
  d->words.resize(static_cast<size_t>(nbytes / 4));
  std::memcpy(d->words.data(), data, static_cast<size_t>(nbytes));
  d->state = static_cast<uint64_t>(d->words[0]) | (static_cast<uint64_t>(d->words[1]) << 32);
  d->pos = 2;
  return d;
}
void reslic_rans_decoder_destroy(void* h) { delete static_cast<Decoder*>(h); }

// Decodes the next n symbols of the stream (decode_stream semantics: the decoder keeps its position).
int reslic_rans_decoder_decode(void* h, const int32_t* indexes, int64_t n, const int32_t* cdfs, int32_t n_cdfs,
                               int32_t cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* out) {
  auto* d = static_cast<Decoder*>(h);
  if (!d || n < 0 || (n > 0 && (!indexes || !out))) return reslic::set_error(RESLIC_ERR_ARG, "rans_decoder_decode: bad argument");
  const Tables t{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
  if (int rc = check_tables(t)) return rc;
  uint64_t x = d->state;
  size_t pos = d->pos;
  const size_t nw = d->words.size();
  auto renorm = [&]() -> bool {
    if (x < kRansL) {
      if (pos >= nw) return false;
      x = (x << 32) | d->words[pos++];
    }
    return true;
  };
  auto get_bits = [&](int nbits, int32_t& val) -> bool {
    val = static_cast<int32_t>(x & ((1u << nbits) - 1));
    x >>= nbits;
    return renorm();
  };
  const uint64_t mask = (1ull << kPrecision) - 1;
  // Symbol lookup: a binary search over a 3000-entry row is a dozen dependent cache misses per symbol.
  // For long runs a coarse table per row — the symbol that holds each multiple of 2^kLutShift — is built
  // first (n_cdfs * 256 entries) and the search becomes one table read plus a short forward scan (the
  // distributions are peaked: buckets that hold many symbols are the ones that are almost never hit).
  constexpr int kLutShift = 8;
  constexpr int kLutSize = 1 << (kPrecision - kLutShift);
  const bool use_lut = n >= 8LL * kLutSize;
  if (use_lut) {
    d->lut.resize(static_cast<size_t>(n_cdfs) * kLutSize);
    for (int32_t r = 0; r < n_cdfs; ++r) {
      const int32_t* cdf = cdfs + static_cast<int64_t>(r) * cdf_stride;
      const int32_t size = cdf_sizes[r];
      int32_t sy = 0;
      for (int k = 0; k < kLutSize; ++k) {
        const int32_t v = k << kLutShift;
        while (sy + 1 < size && cdf[sy + 1] <= v) ++sy;
        d->lut[static_cast<size_t>(r) * kLutSize + k] = sy;
      }
    }
  }
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    if (ci < 0 || ci >= n_cdfs) return reslic::set_error(RESLIC_ERR_ARG, "rans_decoder_decode: cdf index out of range");
    const int32_t* cdf = cdfs + static_cast<int64_t>(ci) * cdf_stride;
    const int32_t size = cdf_sizes[ci], max_value = size - 2;
    const int32_t cum = static_cast<int32_t>(x & mask);
    int32_t s;
    if (use_lut) {
      s = d->lut[static_cast<size_t>(ci) * kLutSize + (cum >> kLutShift)];
      while (s + 1 < size && cdf[s + 1] <= cum) ++s;
      if (cdf[s] > cum) s = -1;             // cum below the row's first entry: corrupt
    } else {
      // first entry > cum, minus one  (cdf is strictly increasing on [0, size))
      const int32_t* it = std::upper_bound(cdf, cdf + size, cum);
      s = static_cast<int32_t>(it - cdf) - 1;
    }
    if (s < 0 || s > max_value) return reslic::set_error(RESLIC_ERR_ARG, "rans_decoder_decode: corrupt stream");
    const uint64_t start = static_cast<uint64_t>(cdf[s]), freq = static_cast<uint64_t>(cdf[s + 1] - cdf[s]);
    x = freq * (x >> kPrecision) + (x & mask) - start;
    if (!renorm()) return reslic::set_error(RESLIC_ERR_ARG, "rans_decoder_decode: stream exhausted");
    int64_t value = s;
    if (s == max_value) {
      int32_t val;
      if (!get_bits(kBypassPrecision, val)) return reslic::set_error(RESLIC_ERR_ARG, "rans_decoder_decode: stream exhausted");
      int32_t n_bypass = val;
      while (val == kMaxBypassVal) {
        if (!get_bits(kBypassPrecision, val)) return reslic::set_error(RESLIC_ERR_ARG, "rans_decoder_decode: stream exhausted");
        n_bypass += val;
      }
      uint64_t raw = 0;
      for (int32_t j = 0; j < n_bypass; ++j) {
        if (!get_bits(kBypassPrecision, val)) return reslic::set_error(RESLIC_ERR_ARG, "rans_decoder_decode: stream exhausted");
        raw |= static_cast<uint64_t>(val) << (j * kBypassPrecision);
      }
      value = static_cast<int64_t>(raw >> 1);
      if (raw & 1) value = -value - 1;
      else value += max_value;
    }
    out[i] = static_cast<int32_t>(value + offsets[ci]);
  }
  d->state = x;
  d->pos = pos;
  return RESLIC_OK;
}

}  // extern "C"
