/* Brute-force k-nearest-neighbour search (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).
 *
 * Independent C restatement of the search at interpolator.py:97,139 (`KDTree.query(q, k)`, p = 2,
 * eps = 0): for every query the k particles of smallest squared distance
 *     d2 = (dx*dx + dy*dy) + dz*dz        (float64, no contraction: compile with -ffp-contract=off)
 * in ascending (d2, index) order -- the canonical tie rule shared by oracle/reference_port.py and the
 * CUDA kernels.  No SciPy, no tree: O(nq * n * log k), for cross-checking at sizes the NumPy
 * brute force cannot reach.  Built by oracle/Makefile into oracle/_build/libknn_brute.so.
 */
#include <stdint.h>
#include <stdlib.h>

static int greater(double ka, int64_t ia, double kb, int64_t ib) { return ka > kb || (ka == kb && ia > ib); }

static void sift_down(double* k, int64_t* id, int n, int pos) {
  const double nk = k[pos];
  const int64_t ni = id[pos];
  for (;;) {
    int c = 2 * pos + 1;
    if (c >= n) break;
    if (c + 1 < n && greater(k[c + 1], id[c + 1], k[c], id[c])) c++;
    if (!greater(k[c], id[c], nk, ni)) break;
    k[pos] = k[c];
    id[pos] = id[c];
    pos = c;
  }
  k[pos] = nk;
  id[pos] = ni;
}

/* points (n,3), queries (nq,3) row-major float64; idx_out (nq,k) int64; d2_out (nq,k) float64.
 * Returns 0, or 1 if k > n. */
int knn_brute(const double* points, int64_t n, const double* queries, int64_t nq, int k, int64_t* idx_out,
              double* d2_out) {
  if (k > n || k < 1) return 1;
  for (int64_t q = 0; q < nq; ++q) {
    const double qx = queries[3 * q], qy = queries[3 * q + 1], qz = queries[3 * q + 2];
    double* hk = d2_out + q * k;
    int64_t* hi = idx_out + q * k;
    int cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
      const double dx = qx - points[3 * i], dy = qy - points[3 * i + 1], dz = qz - points[3 * i + 2];
      const double d2 = (dx * dx + dy * dy) + dz * dz;
      if (cnt < k) {
        hk[cnt] = d2;
        hi[cnt] = i;
        if (++cnt == k)
          for (int h = k / 2 - 1; h >= 0; --h) sift_down(hk, hi, k, h);
      } else if (greater(hk[0], hi[0], d2, i)) {
        hk[0] = d2;
        hi[0] = i;
        sift_down(hk, hi, k, 0);
      }
    }
    for (int m = k; m > 1; --m) { /* heap sort -> ascending (d2, index) */
      const double tk = hk[m - 1];
      const int64_t ti = hi[m - 1];
      hk[m - 1] = hk[0];
      hi[m - 1] = hi[0];
      hk[0] = tk;
      hi[0] = ti;
      sift_down(hk, hi, m - 1, 0);
    }
  }
  return 0;
}
