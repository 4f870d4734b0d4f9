// Host-side packer of the wire format (capi.cu forward_host_impl, bo_pack_rows_host): candidate rows in the layout BoFire
// hands to BoTorch (float64, fingerprint blocks as 0.0 / 1.0 doubles: molfeatures.py:31-48) -> the float64 columns that stay
// dense + one bit per fingerprint column.  Plain C++ (no CUDA): compiled by the host compiler so that the AVX2 variant can
// use intrinsics; the scalar variant is the portable path and the one the tail of every row takes.
#include <stddef.h>
#include <stdint.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "host_pack.h"

// 64 doubles -> one word (bit t = x[t] != 0); *bad is raised when a value is neither 0 nor 1
static inline uint64_t pack64_scalar(const double* x, int n, bool* bad) {
  uint64_t word = 0;
  bool b = false;
  for (int t = 0; t < n; ++t) {
    const double v = x[t];
    word |= (uint64_t)(v != 0.0) << t;
    b |= !((v == 0.0) | (v == 1.0));
  }
  *bad |= b;
  return word;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void pack_block_avx2(const double* xb, int nbits, uint64_t* br, bool* bad) {
  const __m256d zero = _mm256_setzero_pd(), one = _mm256_set1_pd(1.0);
  const int W = (nbits + 63) / 64;
  int ok_mask = 0xF;
  for (int w = 0; w < W; ++w) {
    const int n = nbits - w * 64 < 64 ? nbits - w * 64 : 64;
    if (n < 64) { br[w] = pack64_scalar(xb + w * 64, n, bad); continue; }
    const double* x = xb + w * 64;
    uint64_t word = 0;
    for (int t = 0; t < 64; t += 4) {
      const __m256d v = _mm256_loadu_pd(x + t);
      const int nz = _mm256_movemask_pd(_mm256_cmp_pd(v, zero, _CMP_NEQ_UQ));
      const int is1 = _mm256_movemask_pd(_mm256_cmp_pd(v, one, _CMP_EQ_OQ));
      ok_mask &= (is1 | ~nz);          // a non-zero lane must be exactly 1
      word |= (uint64_t)nz << t;
    }
    br[w] = word;
  }
  if ((ok_mask & 0xF) != 0xF) *bad = true;
}
#endif

bool everest_pack_rows(const double* X, size_t r0, size_t r1, int d, const int* dense_cols, int nd, const int* bit_cols, int nbits,
                       double* dense, unsigned long long* bits) {
  const int W = (nbits + 63) / 64;
  const bool contiguous = nbits > 0 && bit_cols[nbits - 1] - bit_cols[0] == nbits - 1;
  bool bad = false;
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
#else
  const bool avx2 = false;
#endif
  for (size_t r = r0; r < r1; ++r) {
    const double* x = X + r * (size_t)d;
    double* dr = dense + r * (size_t)nd;
    for (int k = 0; k < nd; ++k) dr[k] = x[dense_cols[k]];
    uint64_t* br = reinterpret_cast<uint64_t*>(bits) + r * (size_t)W;
    if (contiguous) {
      const double* xb = x + bit_cols[0];
#if defined(__x86_64__)
      if (avx2) { pack_block_avx2(xb, nbits, br, &bad); continue; }
#endif
      for (int w = 0; w < W; ++w) br[w] = pack64_scalar(xb + w * 64, nbits - w * 64 < 64 ? nbits - w * 64 : 64, &bad);
    } else {
      for (int w = 0; w < W; ++w) {
        const int n = nbits - w * 64 < 64 ? nbits - w * 64 : 64;
        double tmp[64];
        for (int t = 0; t < n; ++t) tmp[t] = x[bit_cols[w * 64 + t]];
        br[w] = pack64_scalar(tmp, n, &bad);
      }
    }
  }
  return !bad;
}
