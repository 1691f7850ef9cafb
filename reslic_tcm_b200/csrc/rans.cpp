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
// CDF indexes) in one fused pass.
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

struct Sym { uint16_t start; uint16_t range; bool bypass; };

struct Encoder { std::vector<Sym> syms; std::vector<uint32_t> out; const uint32_t* begin = nullptr; int64_t nbytes = 0; };
struct Decoder { std::vector<uint32_t> words; size_t pos = 0; uint64_t state = 0; std::vector<int32_t> lut; };

inline void enc_put(uint64_t& x, uint32_t*& ptr, uint32_t start, uint32_t freq, int scale_bits) {
  const uint64_t x_max = ((kRansL >> scale_bits) << 32) * freq;
  if (x >= x_max) { *--ptr = static_cast<uint32_t>(x); x >>= 32; }
  x = ((x / freq) << scale_bits) + (x % freq) + start;
}
inline void enc_put_bits(uint64_t& x, uint32_t*& ptr, uint32_t val, int nbits) {
  const uint64_t freq = 1ull << (16 - nbits);
  const uint64_t x_max = ((kRansL >> 16) << 32) * freq;
  if (x >= x_max) { *--ptr = static_cast<uint32_t>(x); x >>= 32; }
  x = (x << nbits) | val;
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
    e->syms.push_back({static_cast<uint16_t>(start), static_cast<uint16_t>(range), false});
    if (value == max_value) {                       // escape: number of 4-bit groups, then the groups
      int32_t n_bypass = 0;
      while ((raw >> (n_bypass * kBypassPrecision)) != 0) ++n_bypass;
      int32_t val = n_bypass;
      while (val >= kMaxBypassVal) {
        e->syms.push_back({static_cast<uint16_t>(kMaxBypassVal), static_cast<uint16_t>(kMaxBypassVal + 1), true});
        val -= kMaxBypassVal;
      }
      e->syms.push_back({static_cast<uint16_t>(val), static_cast<uint16_t>(val + 1), true});
      for (int32_t j = 0; j < n_bypass; ++j) {
        const int32_t v = static_cast<int32_t>((raw >> (j * kBypassPrecision)) & kMaxBypassVal);
        e->syms.push_back({static_cast<uint16_t>(v), static_cast<uint16_t>(v + 1), true});
      }
    }
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
  e->syms.reserve(e->syms.size() + static_cast<size_t>(n) + static_cast<size_t>(n_esc) * 4);
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t s = slots[i];
    const uint32_t range = s & 0xffffu;
    if (range == 0) return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push_slots: empty slot (zero frequency)");
    e->syms.push_back({static_cast<uint16_t>(s >> 16), static_cast<uint16_t>(range), false});
    if (k < n_esc && esc_pos[k] == i) {
      if (esc_raw[k] < 0) return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push_slots: negative bypass value");
      const uint64_t raw = static_cast<uint64_t>(esc_raw[k]);
      ++k;
      int32_t n_bypass = 0;
      while ((raw >> (n_bypass * kBypassPrecision)) != 0) ++n_bypass;
      int32_t val = n_bypass;
      while (val >= kMaxBypassVal) {
        e->syms.push_back({static_cast<uint16_t>(kMaxBypassVal), static_cast<uint16_t>(kMaxBypassVal + 1), true});
        val -= kMaxBypassVal;
      }
      e->syms.push_back({static_cast<uint16_t>(val), static_cast<uint16_t>(val + 1), true});
      for (int32_t j = 0; j < n_bypass; ++j) {
        const int32_t v = static_cast<int32_t>((raw >> (j * kBypassPrecision)) & kMaxBypassVal);
        e->syms.push_back({static_cast<uint16_t>(v), static_cast<uint16_t>(v + 1), true});
      }
    }
  }
  if (k != n_esc) return reslic::set_error(RESLIC_ERR_ARG, "rans_encoder_push_slots: escape positions not ascending or out of range");
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
  for (size_t k = e->syms.size(); k-- > 0;) {
    const Sym s = e->syms[k];
    if (!s.bypass) enc_put(x, ptr, s.start, s.range, kPrecision);
    else enc_put_bits(x, ptr, s.start, kBypassPrecision);
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
