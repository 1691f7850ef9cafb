// cdf_tables.cpp — host-side pmf -> 16-bit quantised CDF (setup code, runs once per update()).
// Native replacement for the un-vendored third-party op compressai._CXX.pmf_to_quantized_cdf
// (reference call sites src/entropy_models/coder.py:53-56,
// src/entropy_models/adaptive_gaussian_conditional.py:197-205); algorithm: SURVEY.md App. A.5.
#include <cmath>
#include <cstdint>
#include <vector>
#include "reslic_internal.h"

extern "C" int reslic_pmf_to_quantized_cdf(const float* pmf, int32_t n, int32_t precision, uint32_t* cdf) {
  using reslic::set_error;
  if (!pmf || !cdf || n < 1) return set_error(RESLIC_ERR_ARG, "pmf_to_quantized_cdf: null or empty pmf");
  if (precision < 1 || precision > 31) return set_error(RESLIC_ERR_ARG, "pmf_to_quantized_cdf: precision outside 1..31");
  for (int i = 0; i < n; ++i)
    if (!(pmf[i] >= 0.0f) || !std::isfinite(pmf[i]))
      return set_error(RESLIC_ERR_ARG, "pmf_to_quantized_cdf: negative or non-finite probability");
  const uint64_t one = 1ull << precision;
  std::vector<uint64_t> c(static_cast<size_t>(n) + 1, 0);
  uint64_t total = 0;
  for (int i = 0; i < n; ++i) {
    c[i + 1] = static_cast<uint64_t>(std::round(pmf[i] * static_cast<float>(one)));
    total += c[i + 1];
  }
  if (total == 0) return set_error(RESLIC_ERR_ARG, "pmf_to_quantized_cdf: all probabilities are zero");
  for (auto& v : c) v = (one * v) / total;          // renormalise, floor
  for (int i = 1; i <= n; ++i) c[i] += c[i - 1];    // partial sums
  c[n] = one;
  // every symbol needs a non-zero frequency: steal one count from the cheapest donor
  for (int i = 0; i < n; ++i) {
    if (c[i] != c[i + 1]) continue;
    uint64_t best_freq = ~0ull;
    int donor = -1;
    for (int j = 0; j < n; ++j) {
      const uint64_t f = c[j + 1] - c[j];
      if (f > 1 && f < best_freq) { best_freq = f; donor = j; }
    }
    if (donor < 0) return set_error(RESLIC_ERR_ARG, "pmf_to_quantized_cdf: more symbols than 2^precision counts");
    if (donor < i) for (int j = donor + 1; j <= i; ++j) --c[j];
    else for (int j = i + 1; j <= donor; ++j) ++c[j];
  }
  for (int i = 0; i <= n; ++i) cdf[i] = static_cast<uint32_t>(c[i]);
  return RESLIC_OK;
}
