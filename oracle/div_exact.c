/* TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the division the stencil kernels use for a grid
 * spacing h (ptv_interpolation_b200/csrc/bulk_pipe.cuh: div_by_spacing) -- reciprocal rh = RN(1 / h), then two
 * Newton corrections with exact FMA residuals (Markstein) -- checked here against the IEEE quotient the
 * reference computes (physics.py:26-53 and np.gradient divide by dx / dy / dz).  Returns the number of
 * pseudo-random numerators (exponents spread over +-40 binades, signed zeros now and then) whose result
 * differs from t / h in any bit. */
#include <math.h>
#include <stdint.h>
#include <string.h>

static double div_by_spacing(double t, double h, double rh) {
  double q = t * rh;
  double r = fma(-h, q, t);
  q = fma(r, rh, q);
  r = fma(-h, q, t);
  q = fma(r, rh, q);
  if (t == 0.0) q = h > 0.0 ? t : -t; /* 0 / h is a signed zero */
  return q;
}

int64_t div_exact_mismatches(double h, int64_t n, uint64_t seed) {
  const double rh = 1.0 / h;
  uint64_t st = seed ? seed : 88172645463325252ull;
  int64_t bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17; /* xorshift64 */
    const uint64_t mant = st & 0x000fffffffffffffull;
    const uint64_t ex = 1023 - 40 + ((st >> 52) % 81);
    const uint64_t sg = (st >> 63) << 63;
    uint64_t b = sg | (ex << 52) | mant;
    if ((st & 0xff000) == 0) b = sg;
    double t;
    memcpy(&t, &b, 8);
    const double a = div_by_spacing(t, h, rh), c = t / h;
    bad += memcmp(&a, &c, 8) != 0;
  }
  return bad;
}
