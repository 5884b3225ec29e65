// mirror_expand.cc — host side of the narrow mirror transport (mirror_compact.cu): widen one-byte codes into the float32
// values of the caller's layer.  Plain C++ (compiled by the host compiler, not nvcc): AVX2 path selected at run time,
// non-temporal stores (the mirror is written once and not read back by these threads; a regular store would first read
// every destination line into the cache).
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define AMB_HAVE_X86 1
#endif

namespace amb {

static void expand_scalar(const uint8_t* src, float* dst, size_t n, int nan_code) {
  float nanv;
  const uint32_t qnan = 0x7fc00000u;  // the canonical quiet NaN pack_codes_kernel accepted
  std::memcpy(&nanv, &qnan, sizeof(nanv));
  if (nan_code >= 0) {
    for (size_t k = 0; k < n; ++k) dst[k] = src[k] == nan_code ? nanv : static_cast<float>(src[k]);
  } else {
    for (size_t k = 0; k < n; ++k) dst[k] = static_cast<float>(src[k]);
  }
}

#ifdef AMB_HAVE_X86
__attribute__((target("avx2"))) static void expand_avx2(const uint8_t* src, float* dst, size_t n, int nan_code) {
  size_t k = 0;
  // scalar head up to a 32-byte aligned destination
  while (k < n && (reinterpret_cast<uintptr_t>(dst + k) & 31u) != 0) {
    expand_scalar(src + k, dst + k, 1, nan_code);
    ++k;
  }
  const __m256i nan_bits = _mm256_set1_epi32(0x7fc00000);
  const __m256i code = _mm256_set1_epi32(nan_code);
  for (; k + 16 <= n; k += 16) {
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + k));
    const __m256i lo = _mm256_cvtepu8_epi32(b);
    const __m256i hi = _mm256_cvtepu8_epi32(_mm_srli_si128(b, 8));
    __m256 flo = _mm256_cvtepi32_ps(lo), fhi = _mm256_cvtepi32_ps(hi);
    if (nan_code >= 0) {
      flo = _mm256_blendv_ps(flo, _mm256_castsi256_ps(nan_bits), _mm256_castsi256_ps(_mm256_cmpeq_epi32(lo, code)));
      fhi = _mm256_blendv_ps(fhi, _mm256_castsi256_ps(nan_bits), _mm256_castsi256_ps(_mm256_cmpeq_epi32(hi, code)));
    }
    _mm256_stream_ps(dst + k, flo);
    _mm256_stream_ps(dst + k + 8, fhi);
  }
  _mm_sfence();
  if (k < n) expand_scalar(src + k, dst + k, n - k, nan_code);
}
#endif

// dst[k] = (float)src[k], or the canonical NaN where src[k] == nan_code (nan_code < 0: no NaN code)
void expand_codes(const uint8_t* src, float* dst, size_t n, int nan_code) {
#ifdef AMB_HAVE_X86
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) {
    expand_avx2(src, dst, n, nan_code);
    return;
  }
#endif
  expand_scalar(src, dst, n, nan_code);
}

}  // namespace amb
