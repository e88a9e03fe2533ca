// Fused kNN search + weights + accumulate + solid zeroing (replaces tree.query and the NumPy
// weighting at interpolator.py:97-122 / 139-153 and main.py:195-207).
//
// Mapping: one CTA = one TX x TY x TZ tile of voxels, one thread = one voxel.  The tile walks the
// particle cell list ring by ring outwards from its own cells; each ring's cell rows are
// contiguous ranges of 32-byte particle records that the CTA stages through shared memory and
// every thread scans (broadcast reads).  Each thread keeps its k best (d2, index) pairs in a
// shared-memory max-heap laid out [slot][thread] (bank-conflict free for any per-thread slot),
// with the current k-th distance in a register so most candidates are rejected with one compare.
// The search stops when every voxel's k-th distance is inside the scanned box (exact test).
//
// Exactness: d2 = (dx*dx + dy*dy) + dz*dz in float64 with no FMA contraction, ordering is
// (d2, original index) -- the same total order as the canonicalised oracle -- so the selected
// neighbour SET and its sorted order are bit-exact regardless of staging order.
#include "knn_common.cuh"

namespace ptv {

// Place (nk, ni) at `pos` of the size-n max-heap and sift it down.  Column `t` of [slot][T].
template <int T>
__device__ __forceinline__ void sift_down(double* __restrict__ hk, int* __restrict__ hi, int n, int pos,
                                          double nk, int ni) {
  for (;;) {
    int c = 2 * pos + 1;
    if (c >= n) break;
    double kc = hk[c * T];
    int ic = hi[c * T];
    if (c + 1 < n) {
      const double kr = hk[(c + 1) * T];
      const int ir = hi[(c + 1) * T];
      if (key_greater(kr, ir, kc, ic)) {
        kc = kr;
        ic = ir;
        c = c + 1;
      }
    }
    if (!key_greater(kc, ic, nk, ni)) break;
    hk[pos * T] = kc;
    hi[pos * T] = ic;
    pos = c;
  }
  hk[pos * T] = nk;
  hi[pos * T] = ni;
}

// ------------------------------------------------------------------------------------------
// Local RBF (interpolator.py:157-195 -> scipy RBFInterpolator(neighbors=k), thin-plate spline,
// degree-1 polynomial tail; scipy/interpolate/_rbfinterp_xp.py:139-266).  Per voxel the
// (k+4)x(k+4) saddle-point system [[K + sI, P], [P^T, 0]] c = [d; 0] is built and solved in float64
// by the voxel's warp: lane = matrix row, matrix in shared memory (row stride 33 doubles, bank
// conflict free), Gaussian elimination with partial pivoting (the system is symmetric indefinite,
// so no Cholesky; scipy uses LAPACK dsysv), then vec(x) . c with vec = [phi(|x-y_j|), 1, xh, yh, zh].
// RPL = matrix rows per lane: 1 for k + 4 <= 32, 2 for k + 4 <= 64.
__host__ __device__ constexpr int rbf_scratch_doubles(int rpl) {  // rpl 3: the register-resident variant below
  return rpl >= 3 ? 32 * 13 + 16 : (32 * rpl) * (32 * rpl + 1) + (32 * rpl) * 3 * 2 + (32 * rpl) / 2;
}

// Scale-invariant kernels of scipy.interpolate.RBFInterpolator (the only ones the reference can reach:
// it never passes epsilon, _rbfinterp.py:280-289) as functions of d2 = r^2 (_rbfinterp_xp.py:93-110):
// 0 thin_plate_spline r^2 log r (phi(0) = 0), 1 cubic r^3, 2 linear -r, 3 quintic -r^5.
__device__ __forceinline__ double rbf_phi_from_d2(int kern, double d2) {
  if (kern == 0) return d2 > 0.0 ? 0.5 * d2 * log(d2) : 0.0;
  const double r = sqrt(d2);
  if (kern == 1) return d2 * r;
  if (kern == 2) return -r;
  return -(d2 * d2 * r);
}
// Monomials of degree <= 2 in scipy's order (_rbfinterp_common.py:5-32): 1, x, y, z, xx, xy, xz, yy, yz, zz
__device__ __forceinline__ double rbf_monomial(int m, double x, double y, double z) {
  switch (m) {
    case 0: return 1.0;
    case 1: return x;
    case 2: return y;
    case 3: return z;
    case 4: return x * x;
    case 5: return x * y;
    case 6: return x * z;
    case 7: return y * y;
    case 8: return y * z;
    default: return z * z;
  }
}

template <int T, int RPL>
__device__ void rbf_tps_epilogue(const KnnParams& p, double* __restrict__ scratch, const double* hkey_all,
                                 const int* hidx_all, double qx, double qy, double qz, bool active, double& su,
                                 double& sv, double& sw) {
  constexpr int NMAX = 32 * RPL, LD = NMAX + 1;
  const unsigned full = 0xffffffffu;
  const int t = threadIdx.x, lane = t & 31, wbase = t & ~31;
  const int k = p.k, npoly = p.rbf_npoly, n = k + npoly, kern = p.rbf_kernel;
  double* A = scratch;                    // [NMAX][LD]
  double* B = A + NMAX * LD;              // [NMAX][3] right-hand sides -> coefficients
  double* Y = B + NMAX * 3;               // [NMAX][3] neighbour coordinates
  int* perm = reinterpret_cast<int*>(Y + NMAX * 3);  // [NMAX] pivot row of each column
  const HashGrid& g = p.g;
  for (int v = 0; v < 32; ++v) {
    if (!__shfl_sync(full, active ? 1 : 0, v)) continue;
    const double vx = __shfl_sync(full, qx, v), vy = __shfl_sync(full, qy, v), vz = __shfl_sync(full, qz, v);
    // row r = lane + 32*q; rows < k own neighbour r of voxel v
    double yx[RPL], yy[RPL], yz[RPL], d2q[RPL];
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
      const int r = lane + 32 * q;
      yx[q] = yy[q] = yz[q] = d2q[q] = 0.0;
      if (r < k) {
        const int pj = hidx_all[r * T + wbase + v];
        d2q[q] = hkey_all[r * T + wbase + v];
        yx[q] = g.pts[(int64_t)pj * 3 + 0];
        yy[q] = g.pts[(int64_t)pj * 3 + 1];
        yz[q] = g.pts[(int64_t)pj * 3 + 2];
        const Value4 val = g.vals[pj];
        B[r * 3 + 0] = val.u; B[r * 3 + 1] = val.v; B[r * 3 + 2] = val.w;
        Y[r * 3 + 0] = yx[q]; Y[r * 3 + 1] = yy[q]; Y[r * 3 + 2] = yz[q];
        mn[0] = fmin(mn[0], yx[q]); mn[1] = fmin(mn[1], yy[q]); mn[2] = fmin(mn[2], yz[q]);
        mx[0] = fmax(mx[0], yx[q]); mx[1] = fmax(mx[1], yy[q]); mx[2] = fmax(mx[2], yz[q]);
      } else if (r < n) {
        B[r * 3 + 0] = 0.0; B[r * 3 + 1] = 0.0; B[r * 3 + 2] = 0.0;
      }
    }
    // shift / scale of the neighbourhood (_rbfinterp_xp.py:186-193)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mn[c] = fmin(mn[c], __shfl_xor_sync(full, mn[c], o));
        mx[c] = fmax(mx[c], __shfl_xor_sync(full, mx[c], o));
      }
    }
    double shift[3], scale[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      shift[c] = (mx[c] + mn[c]) / 2.0;
      scale[c] = (mx[c] - mn[c]) / 2.0;
      if (scale[c] == 0.0) scale[c] = 1.0;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < RPL; ++q) {  // kernel rows: K + sI | P
      const int r = lane + 32 * q;
      if (r < k) {
        double* row = A + r * LD;
        int j = 0;
        for (; j + 4 <= k; j += 4) {  // four entries per trip: the loads of all four come before the stores
          double d2v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double dx = yx[q] - Y[(j + u) * 3 + 0], dy = yy[q] - Y[(j + u) * 3 + 1], dz = yz[q] - Y[(j + u) * 3 + 2];
            d2v[u] = (dx * dx + dy * dy) + dz * dz;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) row[j + u] = rbf_phi_from_d2(kern, d2v[u]);
        }
        for (; j < k; ++j) {
          const double dx = yx[q] - Y[j * 3 + 0], dy = yy[q] - Y[j * 3 + 1], dz = yz[q] - Y[j * 3 + 2];
          row[j] = rbf_phi_from_d2(kern, (dx * dx + dy * dy) + dz * dz);
        }
        row[r] += p.smoothing;
        const double hx = (yx[q] - shift[0]) / scale[0], hy = (yy[q] - shift[1]) / scale[1],
                     hz = (yz[q] - shift[2]) / scale[2];
        for (int m = 0; m < npoly; ++m) row[k + m] = rbf_monomial(m, hx, hy, hz);
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < RPL; ++q) {  // polynomial rows: P^T | 0
      const int r = lane + 32 * q;
      if (r >= k && r < n) {
        double* row = A + r * LD;
        for (int j = 0; j < k; ++j) row[j] = A[j * LD + r];
        for (int j = k; j < n; ++j) row[j] = 0.0;
      }
    }
    __syncwarp();
    // ---- elimination with partial pivoting; rows stay in place, perm[col] = pivot row
    bool used[RPL];
    int my_col[RPL];
#pragma unroll
    for (int q = 0; q < RPL; ++q) { used[q] = false; my_col[q] = n; }
    bool singular = false;
    for (int c = 0; c < n; ++c) {
      double bv = -1.0;
      int best = NMAX;
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const int r = lane + 32 * q;
        if (r < n && !used[q]) {
          const double a = fabs(A[r * LD + c]);
          if (a > bv) { bv = a; best = r; }  // NaN compares false: never chosen, caught below
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(full, bv, o);
        const int ol = __shfl_xor_sync(full, best, o);
        if (ov > bv || (ov == bv && ol < best)) { bv = ov; best = ol; }
      }
      if (!(bv > 0.0)) { singular = true; break; }  // exact zero (or NaN) pivot column: dsysv info > 0
#pragma unroll
      for (int q = 0; q < RPL; ++q)
        if (lane + 32 * q == best) { used[q] = true; my_col[q] = c; perm[c] = best; }
      const double* prow = A + best * LD;
      const double piv = prow[c];
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const int r = lane + 32 * q;
        if (r < n && !used[q]) {
          double* row = A + r * LD;
          const double f = row[c] / piv;
          if (f != 0.0) {
            // the lane's row is never the pivot row, but the compiler cannot know: four columns per trip with
            // every load issued before the first store, so the shared-memory round trips overlap
            int j = c + 1;
            for (; j + 4 <= n; j += 4) {
              const double p0 = prow[j], p1 = prow[j + 1], p2 = prow[j + 2], p3 = prow[j + 3];
              const double r0 = row[j], r1 = row[j + 1], r2 = row[j + 2], r3 = row[j + 3];
              row[j] = r0 - f * p0;
              row[j + 1] = r1 - f * p1;
              row[j + 2] = r2 - f * p2;
              row[j + 3] = r3 - f * p3;
            }
            for (; j < n; ++j) row[j] -= f * prow[j];
            B[r * 3 + 0] -= f * B[best * 3 + 0];
            B[r * 3 + 1] -= f * B[best * 3 + 1];
            B[r * 3 + 2] -= f * B[best * 3 + 2];
          }
        }
      }
      __syncwarp();
    }
    if (singular) {
      if (lane == 0) atomicExch(p.err_flag, 1);
      __syncwarp();
      continue;
    }
    // ---- back substitution: the row pivoting column c holds unknown c
    for (int c = n - 1; c >= 0; --c) {
      const int pr = perm[c];
      const double inv = 1.0 / A[pr * LD + c];
      const double x0 = B[pr * 3 + 0] * inv, x1 = B[pr * 3 + 1] * inv, x2 = B[pr * 3 + 2] * inv;
      __syncwarp();
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const int r = lane + 32 * q;
        if (r == pr) { B[r * 3 + 0] = x0; B[r * 3 + 1] = x1; B[r * 3 + 2] = x2; }
        else if (r < n && my_col[q] < c) {
          const double a = A[r * LD + c];
          B[r * 3 + 0] -= a * x0; B[r * 3 + 1] -= a * x1; B[r * 3 + 2] -= a * x2;
        }
      }
      __syncwarp();
    }
    // ---- evaluate vec(x) . coeffs (_rbfinterp_xp.py:213-266); row j holds unknown j's weight
    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
      const int r = lane + 32 * q;
      if (r < n) {
        double vj;
        if (r < k) vj = rbf_phi_from_d2(kern, d2q[q]);
        else vj = rbf_monomial(r - k, (vx - shift[0]) / scale[0], (vy - shift[1]) / scale[1], (vz - shift[2]) / scale[2]);
        const int pr = perm[r];
        e0 += vj * B[pr * 3 + 0]; e1 += vj * B[pr * 3 + 1]; e2 += vj * B[pr * 3 + 2];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e0 += __shfl_xor_sync(full, e0, o);
      e1 += __shfl_xor_sync(full, e1, o);
      e2 += __shfl_xor_sync(full, e2, o);
    }
    if (lane == v) { su = e0; sv = e1; sw = e2; }
    __syncwarp();
  }
}

// Register-resident variant for k + tail <= NR <= 32 (kRbf = 3): lane r keeps row r of the system in registers
// (NR doubles + 3 right-hand sides) and the elimination is shuffles and FMAs only -- no shared-memory matrix, and
// 3.4 KB of scratch per warp instead of 9 KB.  Registers cannot be indexed at run time and straight-line code for
// every column (tried first: 45k instructions, 1.7x SLOWER -- a dozen warps stream 100 KB of code per voxel
// through the instruction cache) is out, so the row ROTATES: every step works on column 0 and writes column j
// to j - 1, which makes the step one small loop body.  Rotation rules out back substitution (a frozen pivot row
// would have to be read at a lane-dependent offset), hence Gauss-Jordan: every row but the pivot row is
// eliminated in every step, the pivot rows end with their pivot alone and x = b / pivot.  Partial pivoting as
// above (largest |a|, ties to the lowest row).  Scratch per warp: Y[32][3] coordinates, P[32][10] monomials.
__device__ __noinline__ double rbf_phi_call(int kern, double d2) { return rbf_phi_from_d2(kern, d2); }

template <int T, int NR>
__device__ __noinline__ void rbf_reg_epilogue(const KnnParams& p, double* __restrict__ scratch, const double* hkey_all,
                                              const int* hidx_all, double qx, double qy, double qz, bool active,
                                              double& su, double& sv, double& sw) {
  static_assert(NR % 4 == 0 && NR <= 32, "rows are built four columns per trip");
  const unsigned full = 0xffffffffu;
  const int t = threadIdx.x, lane = t & 31, wbase = t & ~31;
  const int k = p.k, npoly = p.rbf_npoly, n = k + npoly, kern = p.rbf_kernel;
  double* Y = scratch;     // [32][3]
  double* P = Y + 32 * 3;  // [32][10] monomials of the neighbours (scaled coordinates)
  const HashGrid& g = p.g;
  double ru = 0.0, rv = 0.0, rw = 0.0;
#pragma unroll 1
  for (int v = 0; v < 32; ++v) {
    if (!__shfl_sync(full, active ? 1 : 0, v)) continue;
    const double vx = __shfl_sync(full, qx, v), vy = __shfl_sync(full, qy, v), vz = __shfl_sync(full, qz, v);
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    double yx = 0.0, yy = 0.0, yz = 0.0, d2q = 0.0;
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (lane < k) {
      const int pj = hidx_all[lane * T + wbase + v];
      d2q = hkey_all[lane * T + wbase + v];
      yx = g.pts[(int64_t)pj * 3 + 0];
      yy = g.pts[(int64_t)pj * 3 + 1];
      yz = g.pts[(int64_t)pj * 3 + 2];
      const Value4 val = g.vals[pj];
      b0 = val.u; b1 = val.v; b2 = val.w;
      Y[lane * 3 + 0] = yx; Y[lane * 3 + 1] = yy; Y[lane * 3 + 2] = yz;
      mn[0] = mx[0] = yx; mn[1] = mx[1] = yy; mn[2] = mx[2] = yz;
    }
    // shift / scale of the neighbourhood (_rbfinterp_xp.py:186-193)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mn[c] = fmin(mn[c], __shfl_xor_sync(full, mn[c], o));
        mx[c] = fmax(mx[c], __shfl_xor_sync(full, mx[c], o));
      }
    }
    double shift[3], scale[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      shift[c] = (mx[c] + mn[c]) / 2.0;
      scale[c] = (mx[c] - mn[c]) / 2.0;
      if (scale[c] == 0.0) scale[c] = 1.0;
    }
    if (lane < k) {
      const double hx = (yx - shift[0]) / scale[0], hy = (yy - shift[1]) / scale[1], hz = (yz - shift[2]) / scale[2];
      for (int m = 0; m < npoly; ++m) P[lane * 10 + m] = rbf_monomial(m, hx, hy, hz);
    }
    __syncwarp();
    // ---- rows: kernel rows K + sI | P for lanes < k, polynomial rows P^T | 0 for lanes k .. n-1; four columns
    //      per trip enter at the top of the row while the rest moves down
    double a[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) a[j] = 0.0;
#pragma unroll 1
    for (int j0 = 0; j0 < NR; j0 += 4) {
      double nv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        double val = 0.0;
        if (j < k) {  // uniform
          if (lane < k) {
            const double dx = yx - Y[j * 3 + 0], dy = yy - Y[j * 3 + 1], dz = yz - Y[j * 3 + 2];
            val = rbf_phi_call(kern, (dx * dx + dy * dy) + dz * dz) + (j == lane ? p.smoothing : 0.0);
          } else if (lane < n) {
            val = P[j * 10 + (lane - k)];
          }
        } else if (j < n) {
          if (lane < k) val = P[lane * 10 + (j - k)];
        }
        nv[u] = val;
      }
#pragma unroll
      for (int i = 0; i + 4 < NR; ++i) a[i] = a[i + 4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[NR - 4 + u] = nv[u];
    }
    // ---- Gauss-Jordan with partial pivoting on the rotating rows
    bool used = lane >= n;  // lanes beyond the system never pivot
    int my_col = -1;
    double my_piv = 1.0;
    bool singular = false;
#pragma unroll 1
    for (int c = 0; c < n; ++c) {
      // pivot: largest |a[c]| among the unused rows, ties to the lowest row -- two 32-bit warp maxima on the
      // halves of the non-negative double (its bit pattern orders like the value), then the lowest candidate
      const double av = fabs(a[0]);
      const int hi = used ? -1 : __double2hiint(av);
      const int mh = __reduce_max_sync(full, hi);
      const bool c1 = !used && hi == mh;
      const unsigned lo = c1 ? (unsigned)__double2loint(av) : 0u;
      const unsigned ml = __reduce_max_sync(full, lo);
      const unsigned cand = __ballot_sync(full, c1 && lo == ml);
      const int best = __ffs((int)cand) - 1;
      const double piv = __shfl_sync(full, a[0], best < 0 ? 0 : best);
      if (best < 0 || !(fabs(piv) > 0.0)) {  // exact zero (or NaN) pivot column: dsysv info > 0
        singular = true;
        break;
      }
      // multiplier = a[0] * (1 / piv) like LAPACK's dgetf2 (dscal by the reciprocal); the reciprocal comes from
      // MUFU.RCP64H and two Newton steps: a short dependent chain instead of the IEEE division sequence
      double rp;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rp) : "d"(piv));
      rp = fma(fma(-piv, rp, 1.0), rp, rp);
      rp = fma(fma(-piv, rp, 1.0), rp, rp);
      double f = a[0] * rp;
      if (lane == best) { used = true; my_col = c; my_piv = piv; f = 0.0; }
      if (2 * c < n || NR < 16) {
#pragma unroll
        for (int j = 1; j < NR; ++j) {
          const double pj = __shfl_sync(full, a[j], best);
          a[j - 1] = fma(-f, pj, a[j]);
        }
        a[NR - 1] = 0.0;
      } else {  // fewer than n / 2 <= NR / 2 columns are still alive
#pragma unroll
        for (int j = 1; j <= NR / 2; ++j) {
          const double pj = __shfl_sync(full, a[j], best);
          a[j - 1] = fma(-f, pj, a[j]);
        }
      }
      const double p0 = __shfl_sync(full, b0, best), p1 = __shfl_sync(full, b1, best), p2 = __shfl_sync(full, b2, best);
      b0 = fma(-f, p0, b0); b1 = fma(-f, p1, b1); b2 = fma(-f, p2, b2);
    }
    if (singular) {
      if (lane == 0) atomicExch(p.err_flag, 1);
      __syncwarp();
      continue;
    }
    // ---- evaluate vec(x) . coeffs (_rbfinterp_xp.py:213-266): this row pivots unknown my_col, coefficient b / pivot
    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
    {
      const int src = my_col >= 0 && my_col < k ? my_col : lane;
      const double d2c = __shfl_sync(full, d2q, src);  // squared distance voxel -- neighbour my_col
      if (my_col >= 0) {
        double vj;
        if (my_col < k) vj = rbf_phi_call(kern, d2c);
        else vj = rbf_monomial(my_col - k, (vx - shift[0]) / scale[0], (vy - shift[1]) / scale[1], (vz - shift[2]) / scale[2]);
        const double w = vj / my_piv;
        e0 = w * b0; e1 = w * b1; e2 = w * b2;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e0 += __shfl_xor_sync(full, e0, o);
      e1 += __shfl_xor_sync(full, e1, o);
      e2 += __shfl_xor_sync(full, e2, o);
    }
    if (lane == v) { ru = e0; rv = e1; rw = e2; }
    __syncwarp();
  }
  su = ru; sv = rv; sw = rw;
}

template <int T, int TX, int TY, int TZ, typename OutT, int kRbf>  // kRbf: 0 = no RBF, else matrix rows per lane
__device__ __forceinline__ void heap_tile(const KnnParams& p, const int tile, unsigned char* smem_raw) {
  static_assert(TX * TY * TZ == T, "tile shape");
  constexpr int NW = T / 32;
  const int k = p.k;
  double* hkey_all = reinterpret_cast<double*>(smem_raw);               // [k][T]
  ParticleRec* stage = reinterpret_cast<ParticleRec*>(hkey_all + (size_t)k * T);  // [kStageCap]
  double* red = reinterpret_cast<double*>(stage + kStageCap);           // [6][NW]
  int* hidx_all = reinterpret_cast<int*>(red + 6 * NW);                 // [k][T]
  int* seg_start = hidx_all + (size_t)k * T;                            // [T]
  int* seg_off = seg_start + T;                                         // [T+1]
  int* warp_tot = seg_off + T + 1;                                      // [NW]
  double* rbf_scratch = reinterpret_cast<double*>(
      (reinterpret_cast<uintptr_t>(warp_tot + NW) + 15) & ~(uintptr_t)15);  // [NW][rbf_scratch_doubles], RBF only

  const int t = threadIdx.x;
  double* hk = hkey_all + t;
  int* hi = hidx_all + t;

  const HashGrid& g = p.g;
  bool valid, active;
  int64_t vox = 0;
  double qx = 0.0, qy = 0.0, qz = 0.0;
  if (p.qrec != nullptr) {  // point-query mode: T consecutive cell-sorted query records
    const int64_t qi = (int64_t)tile * T + t;
    valid = qi < p.nq;
    if (valid) {
      const ParticleRec r = p.qrec[qi];
      qx = r.x; qy = r.y; qz = r.z;
      vox = r.idx;
    }
    active = valid;
  } else {
    const int tx = tile % p.tiles_x;
    const int ty = (tile / p.tiles_x) % p.tiles_y;
    const int tz = tile / (p.tiles_x * p.tiles_y);
    const int ix = tx * TX + (t % TX);
    const int iy = ty * TY + ((t / TX) % TY);
    const int iz = tz * TZ + (t / (TX * TY));
    valid = ix < p.nx && iy < p.ny && iz < p.nz;
    vox = valid ? ((int64_t)iz * p.ny + iy) * p.nx + ix : 0;
    active = valid && (p.mask == nullptr || p.mask[vox] != 0);
    if (valid) { qx = p.ax[ix]; qy = p.ay[iy]; qz = p.az[iz]; }
  }

  if (!__syncthreads_or(active ? 1 : 0)) {  // tile entirely solid / outside: zero fill
    if (valid) {
      store_out<OutT>(p.u, vox, 0.0);
      store_out<OutT>(p.v, vox, 0.0);
      store_out<OutT>(p.w, vox, 0.0);
      if (p.knn_idx) {
        for (int j = 0; j < k; ++j) {
          p.knn_idx[vox * k + j] = -1;
          p.knn_dist[vox * k + j] = nan("");
        }
      }
    }
    return;
  }

  TileGeom tg;
  tile_geometry<T>(g, active, qx, qy, qz, red, tg);
  const ScanSmem sm{stage, nullptr, nullptr, seg_start, seg_off, warp_tot};

  int count = 0;
  double thr = INFINITY;  // k-th best d2 once the heap is full
  int root_idx = 0x7fffffff;

  // first radius from the local particle density, then geometric growth until every voxel's k-th
  // neighbour is provably inside the scanned region
  const double r_est = estimate_radius<T>(g, tg, p.r0, k, 16, warp_tot);
  // a first, smaller shell fills the heaps with near particles so later candidates are mostly
  // rejected by the threshold compare instead of being sifted in
  double R = r_est > 0.0 ? 0.8 * r_est : 2.0 * g.cell;
  RoundRegion prev = make_region(g, tg, 0.0);
  bool have_prev = false, had_second = false;
  for (;;) {
    const bool last = R >= tg.rmax;
    if (last) R = tg.rmax;
    const RoundRegion rg = make_region(g, tg, R);
    scan_shell<T, kStageCap, false, true, 0, false>(g, tg, rg, prev, have_prev, sm, 0.0, 0.0, 0.0, [&](int m) {
      if (active) {
#pragma unroll 2
        for (int j = 0; j < m; ++j) {
          const double2 xy = *reinterpret_cast<const double2*>(&stage[j].x);
          const double zz = stage[j].z;
          const int pidx = stage[j].idx;
          const double dx = qx - xy.x, dy = qy - xy.y, dz = qz - zz;
          const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
          if (d2 < thr || (d2 == thr && pidx < root_idx)) {
            if (count < k) {
              hk[count * T] = d2;
              hi[count * T] = pidx;
              ++count;
              if (count == k) {
                for (int h = k / 2 - 1; h >= 0; --h) sift_down<T>(hk, hi, k, h, hk[h * T], hi[h * T]);
                thr = hk[0];
                root_idx = hi[0];
              }
            } else {
              sift_down<T>(hk, hi, k, 0, d2, pidx);
              thr = hk[0];
              root_idx = hi[0];
            }
          }
        }
      }
    });
    // exact termination: the k-th distance must be inside the scanned region (margin absorbs
    // rounding in the particle -> cell assignment)
    const double reff = R - 1e-6 * g.cell;
    const bool done = !active || (count >= k && thr < reff * reff);
    if (__syncthreads_and(done ? 1 : 0) || last) break;
    prev = rg;
    have_prev = true;
    R *= (!had_second && r_est > 0.0) ? 1.15 / 0.8 : 1.25;
    had_second = true;
  }

  const int kk = count;  // == k whenever Np >= k (checked on the host)
  double su = 0.0, sv = 0.0, sw = 0.0;
  // slab hash: certified only if the k-th distance stays inside the binned z-range (see knn_duo.cu)
  if (active && p.clip_count != nullptr && (g.clip_lo > -INFINITY || g.clip_hi < INFINITY)) {
    const double dmin = fmin(qz - g.clip_lo, g.clip_hi - qz);
    if (!(count >= k && dmin > 0.0 && thr <= dmin * dmin)) atomicAdd(p.clip_count, 1);
  }

  if (kRbf) {
    // every lane of the warp helps to solve each voxel's local system, so no early exits here
    double* ws = rbf_scratch + (size_t)(t >> 5) * rbf_scratch_doubles(kRbf > 0 ? kRbf : 1);
    if (kRbf >= 3) {  // register-resident rows: 24 (kRbf 3) or 32 (kRbf 4) columns, one kernel each (register budget)
      rbf_reg_epilogue<T, (kRbf == 3 ? 24 : 32)>(p, ws, hkey_all, hidx_all, qx, qy, qz, active && kk == k, su, sv, sw);
    } else {
      rbf_tps_epilogue<T, (kRbf == 2 ? 2 : 1)>(p, ws, hkey_all, hidx_all, qx, qy, qz, active && kk == k, su, sv, sw);
    }
  }

  if (!valid) return;
  if (!active) {
    store_out<OutT>(p.u, vox, 0.0);
    store_out<OutT>(p.v, vox, 0.0);
    store_out<OutT>(p.w, vox, 0.0);
    if (p.knn_idx) {
      for (int j = 0; j < k; ++j) {
        p.knn_idx[vox * k + j] = -1;
        p.knn_dist[vox * k + j] = nan("");
      }
    }
    return;
  }

  if (kk < k) {
    // heap property was never established; order for the outputs below is fixed by the sort
    for (int h = kk / 2 - 1; h >= 0; --h) sift_down<T>(hk, hi, kk, h, hk[h * T], hi[h * T]);
  }

  if (p.knn_idx) {
    // in-place heap sort -> ascending (d2, index); only taken by the parity tests
    for (int n = kk; n > 1; --n) {
      const double lk = hk[(n - 1) * T];
      const int li = hi[(n - 1) * T];
      hk[(n - 1) * T] = hk[0];
      hi[(n - 1) * T] = hi[0];
      sift_down<T>(hk, hi, n - 1, 0, lk, li);
    }
    for (int j = 0; j < k; ++j) {
      p.knn_idx[vox * k + j] = j < kk ? (int64_t)hi[j * T] : (int64_t)g.n;
      p.knn_dist[vox * k + j] = j < kk ? sqrt(hk[j * T]) : INFINITY;
    }
  }

  if (!kRbf && p.method == PTV_METHOD_MADFILTER) {
    // filtering.py:26-48 on the k+1 self-query: drop the first neighbour of the canonical order (the
    // particle itself), median and MAD of the remaining speeds, z-score against the particle's own
    // speed.  The heap columns are reused as scratch for the sorted speeds.
    auto speed_of = [&](int i) {
      const Value4 val = g.vals[i];
      return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(val.u, val.u), __dmul_rn(val.v, val.v)), __dmul_rn(val.w, val.w)));
    };
    int first = 0;
    double kth = 0.0;
    for (int j = 0; j < kk; ++j) {
      if (key_greater(hk[first * T], hi[first * T], hk[j * T], hi[j * T])) first = j;
      kth = fmax(kth, hk[j * T]);
    }
    const int m = kk - 1;
    int n = 0;
    for (int j = 0; j < kk; ++j) {  // insertion sort of the neighbour speeds into hk[0..m)
      if (j == first) continue;
      const double sp = speed_of(hi[j * T]);
      int q = n++;
      while (q > 0 && hk[(q - 1) * T] > sp) {
        hk[q * T] = hk[(q - 1) * T];
        --q;
      }
      hk[q * T] = sp;
    }
    const double med = (m & 1) ? hk[(m / 2) * T] : 0.5 * (hk[(m / 2 - 1) * T] + hk[(m / 2) * T]);
    // absolute deviations, sorted again (they are not monotone in the sorted speeds)
    for (int j = 0; j < m; ++j) {
      const double dv = fabs(hk[j * T] - med);
      hk[j * T] = dv;
    }
    for (int j = 1; j < m; ++j) {
      const double dv = hk[j * T];
      int q = j;
      while (q > 0 && hk[(q - 1) * T] > dv) {
        hk[q * T] = hk[(q - 1) * T];
        --q;
      }
      hk[q * T] = dv;
    }
    const double mad = (m & 1) ? hk[(m / 2) * T] : 0.5 * (hk[(m / 2 - 1) * T] + hk[(m / 2) * T]);
    const double z = fabs(speed_of((int)vox) - med) / (mad + 1e-6);
    p.keep[vox] = z <= p.mad_threshold ? 1 : 0;
    p.kth_dist[vox] = sqrt(kth);
    return;
  }

  if (kRbf) {
    // computed above
  } else if (p.method == PTV_METHOD_NEAREST || (p.method == PTV_METHOD_IDW && kk == 1)) {
    // k = 1: the weight cancels; copy the value (griddata 'nearest', interpolator.py:197)
    int best = 0;
    for (int j = 1; j < kk; ++j)
      if (key_greater(hk[best * T], hi[best * T], hk[j * T], hi[j * T])) best = j;
    const Value4 val = g.vals[hi[best * T]];
    su = val.u; sv = val.v; sw = val.w;
  } else if (p.method == PTV_METHOD_IDW) {
    // interpolator.py:141-153: w = 1/(d**p + 1e-10), out = sum(w*val)/sum(w)
    const double eps = 1e-10;
    const bool p2 = (p.power == 2.0);
    double wsum = 0.0;
    for (int j = 0; j < kk; ++j) {
      const double d2 = hk[j * T];
      const double dp = p2 ? d2 : pow(sqrt(d2), p.power);
      const double wgt = 1.0 / (dp + eps);
      const Value4 val = g.vals[hi[j * T]];
      wsum += wgt;
      su += wgt * val.u;
      sv += wgt * val.v;
      sw += wgt * val.w;
    }
    su /= wsum; sv /= wsum; sw /= wsum;
  } else {  // PTV_METHOD_SIBSON, interpolator.py:102-122
    const double eps = 1e-10;
    double dsum = 0.0, asum = 0.0;
    for (int j = 0; j < kk; ++j) {
      const double d = sqrt(hk[j * T]);
      hk[j * T] = d;  // keep the distance for the next passes
      dsum += d;
      asum += 1.0 / (d + eps);
    }
    const double mean = dsum / kk;
    double var = 0.0;
    for (int j = 0; j < kk; ++j) {
      const double e = hk[j * T] - mean;
      var += e * e;
    }
    const double sd = sqrt(var / kk);  // population std, ddof = 0 (interpolator.py:113)
    const double inv_s = 1.0 / (sd + eps);
    double wsum = 0.0;
    for (int j = 0; j < kk; ++j) {
      const double d = hk[j * T];
      const double wgt = ((1.0 / (d + eps)) / asum) * exp(-d * inv_s);
      const Value4 val = g.vals[hi[j * T]];
      wsum += wgt;
      su += wgt * val.u;
      sv += wgt * val.v;
      sw += wgt * val.w;
    }
    su /= wsum; sv /= wsum; sw /= wsum;
  }
  // main.py:195-199 nan_to_num: NaN -> 0 (cannot arise from finite inputs; kept for parity)
  if (su != su) su = 0.0;
  if (sv != sv) sv = 0.0;
  if (sw != sw) sw = 0.0;
  store_out<OutT>(p.u, vox, su);
  store_out<OutT>(p.v, vox, sv);
  store_out<OutT>(p.w, vox, sw);
}

// One CTA per tile, or -- when the streaming kernel handed over a fail list -- a fixed grid of CTAs
// striding over the listed tiles.
template <int T, int TX, int TY, int TZ, typename OutT, int kRbf>
__global__ void __launch_bounds__(T, (kRbf == 3 ? 3 : 2)) knn_interp_kernel(const KnnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (p.tile_list == nullptr) {
    heap_tile<T, TX, TY, TZ, OutT, kRbf>(p, (int)blockIdx.x, smem_raw);
  } else {
    const int n = *p.tile_count;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
      heap_tile<T, TX, TY, TZ, OutT, kRbf>(p, p.tile_list[i], smem_raw);
      __syncthreads();
    }
  }
}


size_t knn_heap_smem_bytes(int T, int k, int method) {
  const int NW = T / 32;
  size_t b = (size_t)k * T * sizeof(double) + (size_t)kStageCap * sizeof(ParticleRec) +
             (size_t)6 * NW * sizeof(double) + (size_t)k * T * sizeof(int) +
             (size_t)(2 * T + 1 + NW) * sizeof(int);
  b = (b + 15) & ~(size_t)15;
  if (method == PTV_METHOD_RBF) b += 16 + (size_t)NW * rbf_scratch_doubles(k + 10 <= 32 ? (tuning().rbf_regs != 0 ? 3 : 1) : 2) * sizeof(double);  // sized for the largest tail
  return (b + 15) & ~(size_t)15;
}

template <int T, int TX, int TY, int TZ, typename OutT, int kRbf = 0>
static int launch_knn(KnnParams& p, cudaStream_t stream) {
  p.tiles_x = (p.nx + TX - 1) / TX;
  p.tiles_y = (p.ny + TY - 1) / TY;
  p.tiles_z = (p.nz + TZ - 1) / TZ;
  if (p.qrec != nullptr) {
    p.tiles_x = (int)((p.nq + T - 1) / T);
    p.tiles_y = p.tiles_z = 1;
  }
  const int64_t ntiles = (int64_t)p.tiles_x * p.tiles_y * p.tiles_z;
  if (ntiles > 2147483647LL) { set_error("ptv_knn_interp: grid too large for one launch"); return PTV_ERR_INVALID; }
  const size_t smem = knn_heap_smem_bytes(T, p.k, p.method);
  auto kern = knn_interp_kernel<T, TX, TY, TZ, OutT, kRbf>;
  PTV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // fail-list mode: a few CTAs per SM stride over the listed tiles (the count lives on the device)
  const int64_t grid = p.tile_list != nullptr ? (ntiles < 148 * 4 ? ntiles : 148 * 4) : ntiles;
  kern<<<(unsigned)grid, T, smem, stream>>>(p);
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

int launch_knn_heap(KnnParams& p, int T, bool f32, cudaStream_t stream) {
  if (p.method == PTV_METHOD_RBF) {
    if (p.k + p.rbf_npoly <= 24 && tuning().rbf_regs != 0)  // register-resident solve
      return f32 ? launch_knn<128, 8, 4, 4, float, 3>(p, stream) : launch_knn<128, 8, 4, 4, double, 3>(p, stream);
    if (p.k + p.rbf_npoly <= 32 && tuning().rbf_regs != 0)
      return f32 ? launch_knn<128, 8, 4, 4, float, 4>(p, stream) : launch_knn<128, 8, 4, 4, double, 4>(p, stream);
    if (p.k + p.rbf_npoly <= 32)
      return f32 ? launch_knn<128, 8, 4, 4, float, 1>(p, stream) : launch_knn<128, 8, 4, 4, double, 1>(p, stream);
    // up to 60 neighbours: two matrix rows per lane, 64-voxel tiles so the 64x65 systems fit in shared memory
    return f32 ? launch_knn<64, 4, 4, 4, float, 2>(p, stream) : launch_knn<64, 4, 4, 4, double, 2>(p, stream);
  }
  switch (T) {
    case 128: return f32 ? launch_knn<128, 8, 4, 4, float>(p, stream) : launch_knn<128, 8, 4, 4, double>(p, stream);
    case 64: return f32 ? launch_knn<64, 4, 4, 4, float>(p, stream) : launch_knn<64, 4, 4, 4, double>(p, stream);
    default: return f32 ? launch_knn<32, 4, 4, 2, float>(p, stream) : launch_knn<32, 4, 4, 2, double>(p, stream);
  }
}

}  // namespace ptv
