// method='linear' (interpolator.py:197 -> scipy griddata -> LinearNDInterpolator: Qhull Delaunay +
// find_simplex + barycentric weights, fill_value 0 outside the convex hull) without building a
// triangulation.
//
// For points in general position the Delaunay triangulation is unique, and the tetrahedron holding a
// query q is the optimum of a 4-variable linear programme: of all spheres with no particle strictly
// inside, the one that holds q deepest (largest r^2 - |q-c|^2) is the circumsphere of that
// tetrahedron (lifting map: the lower-hull facet under q).  One warp solves the programme for one
// voxel by dual-simplex pivoting: start from a huge tetrahedron of four virtual points around q,
// bring in a particle that lies inside the current circumsphere (all 32 lanes scan candidates; the one
// with the largest violation per squared distance from q enters), drop the vertex picked by the ratio
// test that keeps q inside.  Every pivot raises the objective and keeps q inside, whatever set the
// entering particle was drawn from, so the search may look anywhere; the answer is accepted only when
// all four vertices are real and the cells its circumsphere touches were scanned without a violator --
// then no particle anywhere is inside it and the tetrahedron is THE Delaunay tetrahedron of q, however
// it was found.
//
//   * a warp owns an 8x4 row block of a tile: the candidates within R of the block are gathered once
//     into shared memory (structure of arrays, lane j reads element j: conflict-free) and serve all 32
//     voxels; a voxel that falls inside the previous voxel's tetrahedron re-uses it without a search,
//     otherwise that tetrahedron's vertices seed the programme (lp_seed);
//   * voxels whose sphere leaves the gathered region (next to grains, near the hull) are finished inline
//     by finish_voxel: scan the cells the sphere touches, pivot, repeat until a scan finds nothing;
//   * outside the hull the programme is unbounded: the virtual vertices never leave.  To decide that
//     without scanning all particles the programme is re-run on cache U H, H = a superset of the cloud's
//     extreme points built once per hash (ensure_hull_list: cell-level then particle-level octant
//     dominance; ~2000 records for 10M particles, in chunks of 32 with bounding boxes);
//   * q is nudged by 2^-36 of its distance towards an interior point of the cloud before the search, so
//     a voxel lying exactly on a face of the triangulation (lattice wall particles, main.py:173-178) or
//     on the hull is not a degenerate programme; the weights are computed for the unmoved q.
//
// Oracle: oracle/delaunay_lp.py restates the programme by brute force; tests compare simplex vertex
// sets with scipy.spatial.Delaunay.find_simplex and values with golden vectors of the reference.
#include <limits.h>

#include "knn_common.cuh"

namespace ptv {
namespace {

constexpr int kCap = 416;         // cached candidates per warp (32 B each, 53 KB per CTA)
constexpr int kMaxPivots = 600;
constexpr double kEta = 1.0 / 68719476736.0;  // 2^-36
constexpr unsigned kFull = 0xffffffffu;

struct WarpCache {
  double x[kCap], y[kCap], z[kCap];
  int idx[kCap];
  float rd[kCap];  // 1 / |p - q|^2 for the voxel being solved (set_query)
};

// per-voxel part of the cache: reciprocal squared distances from the (nudged) query, used to scale violations
__device__ __forceinline__ void set_query(WarpCache& wc, int n, double qx, double qy, double qz) {
  __syncwarp();
  for (int j = (threadIdx.x & 31); j < n; j += 32) {
    const double ex = wc.x[j] - qx, ey = wc.y[j] - qy, ez = wc.z[j] - qz;
    wc.rd[j] = __frcp_rn((float)(ex * ex + ey * ey + ez * ez));
  }
  __syncwarp();
}

struct Tet {
  double x[4], y[4], z[4];
  int id[4];  // original particle row, or -1..-4 for the virtual vertices
};

struct Geo {  // rows of the inverse edge matrix, circumcentre relative to vertex 0, squared radius
  double r1x, r1y, r1z, r2x, r2y, r2z, r3x, r3y, r3z;
  double cx, cy, cz, cc;
};

__device__ __forceinline__ void tet_geo(const Tet& t, Geo& g) {
  const double ax = t.x[1] - t.x[0], ay = t.y[1] - t.y[0], az = t.z[1] - t.z[0];
  const double bx = t.x[2] - t.x[0], by = t.y[2] - t.y[0], bz = t.z[2] - t.z[0];
  const double dx = t.x[3] - t.x[0], dy = t.y[3] - t.y[0], dz = t.z[3] - t.z[0];
  const double n1x = by * dz - bz * dy, n1y = bz * dx - bx * dz, n1z = bx * dy - by * dx;  // e2 x e3
  const double n2x = dy * az - dz * ay, n2y = dz * ax - dx * az, n2z = dx * ay - dy * ax;  // e3 x e1
  const double n3x = ay * bz - az * by, n3y = az * bx - ax * bz, n3z = ax * by - ay * bx;  // e1 x e2
  const double inv = 1.0 / (ax * n1x + ay * n1y + az * n1z);
  g.r1x = n1x * inv; g.r1y = n1y * inv; g.r1z = n1z * inv;
  g.r2x = n2x * inv; g.r2y = n2y * inv; g.r2z = n2z * inv;
  g.r3x = n3x * inv; g.r3y = n3y * inv; g.r3z = n3z * inv;
  const double h1 = 0.5 * (ax * ax + ay * ay + az * az);
  const double h2 = 0.5 * (bx * bx + by * by + bz * bz);
  const double h3 = 0.5 * (dx * dx + dy * dy + dz * dz);
  g.cx = h1 * g.r1x + h2 * g.r2x + h3 * g.r3x;
  g.cy = h1 * g.r1y + h2 * g.r2y + h3 * g.r3y;
  g.cz = h1 * g.r1z + h2 * g.r2z + h3 * g.r3z;
  g.cc = g.cx * g.cx + g.cy * g.cy + g.cz * g.cz;
}

__device__ __forceinline__ void bary(const Tet& t, const Geo& g, double x, double y, double z, double b[4]) {
  const double dx = x - t.x[0], dy = y - t.y[0], dz = z - t.z[0];
  b[1] = g.r1x * dx + g.r1y * dy + g.r1z * dz;
  b[2] = g.r2x * dx + g.r2y * dy + g.r2z * dz;
  b[3] = g.r3x * dx + g.r3y * dy + g.r3z * dz;
  b[0] = 1.0 - b[1] - b[2] - b[3];
}

__device__ __forceinline__ void init_virtual(Tet& t, double qx, double qy, double qz, double M) {
  // any tetrahedron around q will do; skewed so that lattice data meets no exact symmetry
  const double d[4][3] = {{1.0, 1.1, 0.9}, {1.05, -1.0, -0.95}, {-1.0, 0.93, -1.07}, {-0.97, -1.02, 1.01}};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t.x[i] = qx + M * d[i][0];
    t.y[i] = qy + M * d[i][1];
    t.z[i] = qz + M * d[i][2];
    t.id[i] = -1 - i;
  }
}

__device__ __forceinline__ bool all_real(const Tet& t) { return (t.id[0] | t.id[1] | t.id[2] | t.id[3]) >= 0; }

// distance from point c to the box [lo, hi]
__device__ __forceinline__ double box_dist(const TileGeom& tg, double cx, double cy, double cz) {
  const double dx = fmax(0.0, fmax(tg.lo[0] - cx, cx - tg.hi[0]));
  const double dy = fmax(0.0, fmax(tg.lo[1] - cy, cy - tg.hi[1]));
  const double dz = fmax(0.0, fmax(tg.lo[2] - cz, cz - tg.hi[2]));
  return sqrt(dx * dx + dy * dy + dz * dz);
}

__device__ __forceinline__ void set_rmax(const HashGrid& g, TileGeom& tg) {
  const double glo[3] = {g.ox, g.oy, g.oz};
  const double ghi[3] = {g.ox + g.cnx * g.cell, g.oy + g.cny * g.cell, g.oz + g.cnz * g.cell};
  double r2 = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double m = fmax(fmax(tg.hi[c] - glo[c], ghi[c] - tg.lo[c]), 0.0);
    r2 += m * m;
  }
  tg.rmax = sqrt(r2) * (1.0 + 1e-9) + 1e-3 * g.cell;
}

// The candidate set of one programme: the cells of region(rg) around box tg -- cached in shared memory
// (n >= 0) or read from global memory (n < 0) -- plus, optionally, the hull-candidate list.
struct CandSet {
  int n;
  TileGeom tg;
  RoundRegion rg;
  bool with_hull;
};

// Visit the records of region(rg) with all 32 lanes busy: rows are resolved 32 at a time (one per lane),
// their record counts prefix-summed, and record j of the batch goes to lane j % 32, which finds its row
// by a binary search over the lanes' offsets (shuffles).  f(batch_position, global_record_index).
// before(batch_total) may stop the walk (returns false) -- used by the gather when the cache is full.
template <typename B, typename F>
__device__ __forceinline__ bool for_each_region_record(const HashGrid& g, const TileGeom& tg, const RoundRegion& rg,
                                                       B&& before, F&& f) {
  const int lane = threadIdx.x & 31;
  const int nrows = region_slots(rg, false);
  for (int base = 0; base < nrows; base += 32) {
    int start, cnt;
    resolve_slot(g, tg, rg, rg, false, base + lane, nrows, start, cnt);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += v;
    }
    const int tot = __shfl_sync(kFull, inc, 31);
    if (!before(tot)) return false;
    const int excl = inc - cnt;
    for (int j0 = 0; j0 < tot; j0 += 32) {
      const int j = j0 + lane;
      int lo = 0;  // last lane whose exclusive offset is <= j
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int cand = lo + step;
        const int off = __shfl_sync(kFull, excl, cand & 31);
        if (cand < 32 && off <= j) lo = cand;
      }
      const int rstart = __shfl_sync(kFull, start, lo);
      const int rexcl = __shfl_sync(kFull, excl, lo);
      if (j < tot) f(j, rstart + (j - rexcl));
    }
  }
  return true;
}

// Gather the records of region(rg) into the warp's cache.  Returns the count, or -1 if they do not fit.
__device__ __forceinline__ int warp_gather(const HashGrid& g, const TileGeom& tg, const RoundRegion& rg,
                                           WarpCache& wc) {
  int total = 0, next = 0;
  __syncwarp();
  const bool ok = for_each_region_record(
      g, tg, rg,
      [&](int tot) {
        total = next;
        next += tot;
        return next <= kCap;
      },
      [&](int j, int r) {
        const int4* src = reinterpret_cast<const int4*>(g.rec + r);
        const int4 a = __ldg(src), c = __ldg(src + 1);
        const int off = total + j;
        wc.x[off] = __hiloint2double(a.y, a.x);
        wc.y[off] = __hiloint2double(a.w, a.z);
        wc.z[off] = __hiloint2double(c.y, c.x);
        wc.idx[off] = c.z;
      });
  __syncwarp();
  return ok ? next : -1;
}

// The entering particle: among those strictly inside the circumsphere of t (violation 2 c.d - d.d =
// r^2 - |p - centre|^2 > 1e-12 r^2, d = p - v0) the one with the largest violation per squared distance
// from the query.  Any violator is a valid pivot; the deepest one is a poor choice while virtual vertices
// make the "sphere" a half-space (it is the FARTHEST particle on that side), the scaled rule prefers
// particles near q and reaches the local scale in fewer pivots.  Ties go to the smaller particle row; all
// lanes return the same answer.
__device__ __forceinline__ bool find_violator(const KnnParams& p, const CandSet& cs, const WarpCache& wc,
                                              const Tet& t, const Geo& geo, double qx, double qy, double qz,
                                              double& px, double& py, double& pz, int& pid) {
  const int lane = threadIdx.x & 31;
  const HashGrid& g = p.g;
  const double tolv = 1e-12 * fmax(geo.cc, 1e-300);
  int bid = INT_MAX;
  double bx = 0.0, by = 0.0, bz = 0.0;
  const double tx = t.x[0], ty = t.y[0], tz = t.z[0];
  const double c2x = 2.0 * geo.cx, c2y = 2.0 * geo.cy, c2z = 2.0 * geo.cz;
  float best = 0.0f;
  auto consider = [&](double x, double y, double z, int id, float rdist) {
    const double dx = x - tx, dy = y - ty, dz = z - tz;
    const double viol = (c2x * dx + c2y * dy + c2z * dz) - (dx * dx + dy * dy + dz * dz);
    // (the tetrahedron's own vertices need no test: they lie ON the sphere, violation 0 up to rounding of
    // ~1e-15 r^2, far below the threshold)
    if (viol > tolv) {
      const float score = fmaxf((float)viol * rdist, 1e-37f);  // float32 is plenty for a preference
      if (score > best) { best = score; bid = id; bx = x; by = y; bz = z; }
    }
  };
  auto rdist_of = [&](double x, double y, double z) {
    const double ex = x - qx, ey = y - qy, ez = z - qz;
    return __frcp_rn((float)(ex * ex + ey * ey + ez * ez));
  };
  if (cs.n >= 0) {
    for (int j = lane; j < cs.n; j += 32) consider(wc.x[j], wc.y[j], wc.z[j], wc.idx[j], wc.rd[j]);
  } else {
    for_each_region_record(
        g, cs.tg, cs.rg, [](int) { return true; },
        [&](int, int r) {
          const int4* src = reinterpret_cast<const int4*>(g.rec + r);
          const int4 a = __ldg(src), c = __ldg(src + 1);
          const double x = __hiloint2double(a.y, a.x), y = __hiloint2double(a.w, a.z), z = __hiloint2double(c.y, c.x);
          consider(x, y, z, c.z, rdist_of(x, y, z));
        });
  }
  if (cs.with_hull) {
    // the list comes in chunks of 32 records with a bounding box each: the violation is a separable concave
    // quadratic, so its maximum over a box is exact per axis -- a chunk that cannot hold a violator is skipped
    const int nchunks = p.hull_n >> 5;
    const double cut = 0.9 * tolv;
    for (int base = 0; base < nchunks; base += 32) {
      const int ch = base + lane;
      bool flag = false;
      if (ch < nchunks) {
        const double* bb = p.hull_box + (size_t)ch * 6;
        double ub = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double ta = a == 0 ? tx : (a == 1 ? ty : tz);
          const double ca = a == 0 ? geo.cx : (a == 1 ? geo.cy : geo.cz);
          const double ds = fmin(fmax(ca, bb[a] - ta), bb[3 + a] - ta);
          ub += 2.0 * ca * ds - ds * ds;
        }
        flag = ub > cut;
      }
      unsigned todo = __ballot_sync(kFull, flag);
      while (todo != 0) {
        const int c = __ffs(todo) - 1;
        todo &= todo - 1;
        const int4* src = reinterpret_cast<const int4*>(p.hull_rec + ((size_t)(base + c) << 5) + lane);
        const int4 a = __ldg(src), c4 = __ldg(src + 1);
        const double x = __hiloint2double(a.y, a.x), y = __hiloint2double(a.w, a.z), z = __hiloint2double(c4.y, c4.x);
        consider(x, y, z, c4.z, rdist_of(x, y, z));
      }
    }
  }
  float rb = best;
  int ri = bid;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(kFull, rb, o);
    const int oi = __shfl_xor_sync(kFull, ri, o);
    if (ob > rb || (ob == rb && oi < ri)) { rb = ob; ri = oi; }
  }
  if (ri == INT_MAX) return false;
  const unsigned who = __ballot_sync(kFull, bid == ri && best == rb);
  const int src = __ffs(who) - 1;
  px = __shfl_sync(kFull, bx, src);
  py = __shfl_sync(kFull, by, src);
  pz = __shfl_sync(kFull, bz, src);
  pid = ri;
  return true;
}

// One dual-simplex pivot: particle (px,py,pz,pid) enters, the ratio test picks the vertex that leaves.
__device__ __forceinline__ bool pivot(Tet& t, const Geo& geo, double qx, double qy, double qz, double px, double py,
                                      double pz, int pid) {
  double lam[4], mu[4];
  bary(t, geo, qx, qy, qz, lam);
  bary(t, geo, px, py, pz, mu);
  // ratio test: the smallest lam_i / mu_i over mu_i > 0, compared by cross-multiplication (no divisions)
  int out = -1;
  double bl = 0.0, bm = 1.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double li = fmax(lam[i], 0.0);
    if (mu[i] > 1e-14 && (out < 0 || li * bm < bl * mu[i])) { bl = li; bm = mu[i]; out = i; }
  }
  if (out < 0) return false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i == out) { t.x[i] = px; t.y[i] = py; t.z[i] = pz; t.id[i] = pid; }
  }
  if (t.id[0] < 0) {  // keep a real vertex as the reference point: small magnitudes in the violation
#pragma unroll
    for (int i = 3; i >= 1; --i) {
      if (t.id[i] >= 0 && t.id[0] < 0) {
        double s;
        s = t.x[0]; t.x[0] = t.x[i]; t.x[i] = s;
        s = t.y[0]; t.y[0] = t.y[i]; t.y[i] = s;
        s = t.z[0]; t.z[0] = t.z[i]; t.z[i] = s;
        const int k = t.id[0]; t.id[0] = t.id[i]; t.id[i] = k;
      }
    }
  }
  return true;
}

// Dual-simplex pivots until no candidate lies inside the circumsphere.  (qx,qy,qz) is the nudged query.
// Returns 0 when converged, 1 on the pivot limit / a degenerate step.
__device__ __forceinline__ int lp_run(const KnnParams& p, const CandSet& cs, const WarpCache& wc, Tet& t, double qx,
                                      double qy, double qz, int& pivots) {
  for (int it = 0; it < kMaxPivots; ++it) {
    Geo geo;
    tet_geo(t, geo);
    double px, py, pz;
    int pid;
    if (!find_violator(p, cs, wc, t, geo, qx, qy, qz, px, py, pz, pid)) return 0;
    if (!pivot(t, geo, qx, qy, qz, px, py, pz, pid)) return 1;
    ++pivots;
  }
  return 1;
}

// The same programme over just the four vertices of `seed` (the neighbouring voxel's tetrahedron), every
// lane for itself: takes the tetrahedron from the huge virtual one down to the local scale without a
// candidate scan.  Any outcome is a valid (dual-feasible) start for lp_run.
__device__ __forceinline__ void lp_seed(const Tet& seed, Tet& t, double qx, double qy, double qz) {
  for (int it = 0; it < 8; ++it) {
    Geo geo;
    tet_geo(t, geo);
    const double c2x = 2.0 * geo.cx, c2y = 2.0 * geo.cy, c2z = 2.0 * geo.cz;
    double best = 1e-12 * fmax(geo.cc, 1e-300);
    int bi = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double dx = seed.x[i] - t.x[0], dy = seed.y[i] - t.y[0], dz = seed.z[i] - t.z[0];
      const double viol = (c2x * dx + c2y * dy + c2z * dz) - (dx * dx + dy * dy + dz * dz);
      const int id = seed.id[i];
      if (viol > best && id != t.id[0] && id != t.id[1] && id != t.id[2] && id != t.id[3]) { best = viol; bi = i; }
    }
    if (bi < 0) return;
    double px = 0.0, py = 0.0, pz = 0.0;
    int pid = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i == bi) { px = seed.x[i]; py = seed.y[i]; pz = seed.z[i]; pid = seed.id[i]; }
    }
    if (!pivot(t, geo, qx, qy, qz, px, py, pz, pid)) return;
  }
}

struct VoxelOut {
  double u, v, w;
  int id[4];
  double b[4];
};

__device__ __forceinline__ void emit(const KnnParams& p, const Tet& t, double qx, double qy, double qz, VoxelOut& o) {
  Geo geo;
  tet_geo(t, geo);
  double b[4];
  bary(t, geo, qx, qy, qz, b);
  double su = 0.0, sv = 0.0, sw = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const Value4 val = p.g.vals[t.id[i]];
    su += b[i] * val.u;
    sv += b[i] * val.v;
    sw += b[i] * val.w;
    o.id[i] = t.id[i];
    o.b[i] = b[i];
  }
  o.u = su; o.v = sv; o.w = sw;
}

// stats slots: 0 shared-pass voxels, 1 re-used tetrahedra, 2 general-path voxels, 3 global-memory
// candidate sets, 4 voxels outside the hull, 5 unresolved (pivot limit), 6 pivots
__device__ __forceinline__ void stat_add(const KnnParams& p, int slot, unsigned long long v) {
  if (p.stats != nullptr && (threadIdx.x & 31) == 0) atomicAdd(p.stats + slot, v);
}

// Finish a voxel whose tetrahedron t has converged on the candidate set cs (cached in wc).  Every pivot
// raises the objective of the programme and keeps q inside, whatever set the entering particle was drawn
// from, so the search is free to look where it pays:
//   * virtual vertices left: q is outside the hull of its neighbourhood -> add the hull-candidate list;
//     still virtual -> outside the hull;
//   * verification by the sphere itself: scan the cells the circumsphere touches (from shared memory if
//     they fit, else straight from global memory, one pivot per scan so that the next scan follows the
//     smaller sphere); a violator shrinks the sphere, no violator proves it empty.
// r_cov: the set of cs covers B(q, r_cov) -- a sphere inside it needs no scan.  `clobbered` is set when the
// warp cache was overwritten.  Returns kVerified, kOutside, kFailed or kVirtual (no hull list: the caller
// must widen the region).
enum { kVerified = 0, kOutside = 1, kFailed = 2, kVirtual = 3 };
__device__ int finish_voxel(const KnnParams& p, WarpCache& wc, CandSet& cs, double qx, double qy, double qz,
                            double qpx, double qpy, double qpz, double r_cov, Tet& t, int& pivots,
                            bool& clobbered) {
  const HashGrid& g = p.g;
  const double margin = 1e-6 * g.cell;
  if (!all_real(t)) {
    if (p.hull_rec == nullptr) return kVirtual;
    cs.with_hull = true;
    const int rc = lp_run(p, cs, wc, t, qpx, qpy, qpz, pivots);
    cs.with_hull = false;
    if (rc != 0) return kFailed;
    if (!all_real(t)) return kOutside;  // the programme over a superset of the hull vertices is unbounded
    // q is inside the hull, but the tetrahedron just found hangs on far-away extreme points: its sphere
    // spans the cloud and verifying it would read everything.  Start again from the virtual tetrahedron and
    // widen the ball around q (x1.6 per step) until the virtual vertices are gone -- they must go, q is
    // inside.  (Restarting costs a few pivots on rare voxels; keeping a copy of the local tetrahedron alive
    // across the hull phase costs every voxel registers.)
    const double r_loc = r_cov > 0.0 ? r_cov : 2.0 * g.cell;
    {
      Geo geo;
      tet_geo(t, geo);
      if (geo.cc <= 36.0 * r_loc * r_loc) goto verify;  // a sphere of the local scale: keep it
    }
    TileGeom qg;
    qg.lo[0] = qg.hi[0] = qx;
    qg.lo[1] = qg.hi[1] = qy;
    qg.lo[2] = qg.hi[2] = qz;
    set_rmax(g, qg);
    init_virtual(t, qx, qy, qz, 1e4 * qg.rmax);
    double rb = r_loc;
    while (!all_real(t)) {
      if (rb >= qg.rmax) return kFailed;  // cannot happen: the hull phase found a real tetrahedron
      rb = fmin(1.6 * rb, qg.rmax);
      CandSet bs;
      bs.with_hull = false;
      bs.tg = qg;
      bs.rg = make_region(g, bs.tg, rb);
      bs.n = warp_gather(g, bs.tg, bs.rg, wc);
      clobbered = true;
      if (bs.n >= 0) set_query(wc, bs.n, qpx, qpy, qpz);
      else stat_add(p, 3, 1);
      if (lp_run(p, bs, wc, t, qpx, qpy, qpz, pivots) != 0) return kFailed;
    }
    r_cov = rb;  // converged on a set covering B(q, rb)
  }
verify:
  bool first = true;
  for (;;) {
    Geo geo;
    tet_geo(t, geo);
    const double ccx = t.x[0] + geo.cx, ccy = t.y[0] + geo.cy, ccz = t.z[0] + geo.cz, rad = sqrt(geo.cc);
    if (first) {
      const double ddx = ccx - qx, ddy = ccy - qy, ddz = ccz - qz;
      if (sqrt(ddx * ddx + ddy * ddy + ddz * ddz) + rad <= r_cov - margin) return kVerified;
      first = false;
    }
    CandSet sp;
    sp.tg.lo[0] = sp.tg.hi[0] = ccx;
    sp.tg.lo[1] = sp.tg.hi[1] = ccy;
    sp.tg.lo[2] = sp.tg.hi[2] = ccz;
    set_rmax(g, sp.tg);
    sp.rg = make_region(g, sp.tg, fmin(rad * (1.0 + 1e-9) + 2.0 * margin, sp.tg.rmax));
    sp.with_hull = false;
    sp.n = warp_gather(g, sp.tg, sp.rg, wc);
    clobbered = true;
    const int before = pivots;
    if (sp.n >= 0) {
      set_query(wc, sp.n, qpx, qpy, qpz);
      if (lp_run(p, sp, wc, t, qpx, qpy, qpz, pivots) != 0) return kFailed;
    } else {
      stat_add(p, 3, 1);
      double px, py, pz;
      int pid;
      if (find_violator(p, sp, wc, t, geo, qpx, qpy, qpz, px, py, pz, pid)) {
        if (!pivot(t, geo, qpx, qpy, qpz, px, py, pz, pid) || ++pivots > 4 * kMaxPivots) return kFailed;
      }
    }
    if (pivots == before) return kVerified;  // nothing inside the sphere
  }
}

// One voxel from scratch (the warp's shared candidate set did not fit the cache, or there is no hull list
// and the neighbourhood must grow until it covers the cloud).
__device__ int solve_general(const KnnParams& p, WarpCache& wc, double qx, double qy, double qz, double qpx,
                             double qpy, double qpz, double r_first, double M, Tet& t, int& pivots) {
  const HashGrid& g = p.g;
  CandSet cs;
  cs.tg.lo[0] = cs.tg.hi[0] = qx;
  cs.tg.lo[1] = cs.tg.hi[1] = qy;
  cs.tg.lo[2] = cs.tg.hi[2] = qz;
  set_rmax(g, cs.tg);
  double R = fmin(r_first, cs.tg.rmax);
  init_virtual(t, qx, qy, qz, M);
  for (;;) {
    const bool covers_all = R >= cs.tg.rmax;
    cs.rg = make_region(g, cs.tg, R);
    cs.with_hull = false;
    cs.n = warp_gather(g, cs.tg, cs.rg, wc);
    if (cs.n < 0) stat_add(p, 3, 1);
    else set_query(wc, cs.n, qpx, qpy, qpz);
    if (lp_run(p, cs, wc, t, qpx, qpy, qpz, pivots) != 0) return kFailed;
    bool clobbered = false;
    const int rc = finish_voxel(p, wc, cs, qx, qy, qz, qpx, qpy, qpz, R, t, pivots, clobbered);
    if (rc != kVirtual) return rc;
    if (covers_all) return kOutside;
    R = fmin(1.6 * R, cs.tg.rmax);
  }
}

template <typename OutT, int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks) delaunay_linear_kernel(const KnnParams p) {
  constexpr int T = 128, TX = 8, TY = 4, TZ = 4, NW = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpCache* cache = reinterpret_cast<WarpCache*>(smem_raw);  // [NW]
  __shared__ double red[6 * NW];
  __shared__ int warp_tot[NW];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int tile = blockIdx.x;
  const HashGrid& g = p.g;
  bool valid, active;
  int64_t vox = 0;
  double qx = 0.0, qy = 0.0, qz = 0.0;
  if (p.qrec != nullptr) {
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
  VoxelOut mine;
  mine.u = mine.v = mine.w = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { mine.id[i] = -1; mine.b[i] = nan(""); }

  if (__syncthreads_or(active ? 1 : 0)) {
    TileGeom tg;
    tile_geometry<T>(g, active, qx, qy, qz, red, tg);
    // radius expected to hold 64 particles = 2.5 mean spacings: Delaunay circumspheres of a voxel
    // (centre offset + radius) rarely reach further
    const double r_est = estimate_radius<T>(g, tg, p.r0, p.k, 16, warp_tot);
    const double r_first = r_est > 0.0 ? r_est : 2.0 * g.cell;
    // the virtual vertices: far enough that, over the extent of the cloud, a sphere through one of them is
    // a plane to 1e-4 of that extent
    const double Mv = 1e4 * tg.rmax;
    // an interior point of the cloud (mean of four spread-out particles) to nudge queries towards
    double ctr[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const ParticleRec r = g.rec[(g.n - 1) * i / 3];
      ctr[0] += 0.25 * r.x; ctr[1] += 0.25 * r.y; ctr[2] += 0.25 * r.z;
    }
    // ---- from here on the warps work independently
    const unsigned act = __ballot_sync(kFull, active);
    if (act != 0) {
      WarpCache& wc = cache[wid];
      CandSet cs;
      double v6[6] = {active ? qx : INFINITY, active ? qy : INFINITY, active ? qz : INFINITY,
                      active ? -qx : INFINITY, active ? -qy : INFINITY, active ? -qz : INFINITY};
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 6; ++c) v6[c] = fmin(v6[c], __shfl_xor_sync(kFull, v6[c], o));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { cs.tg.lo[c] = v6[c]; cs.tg.hi[c] = -v6[c + 3]; }
      set_rmax(g, cs.tg);
      const double R = fmin(r_first, cs.tg.rmax);
      const double margin = 1e-6 * g.cell;
      cs.rg = make_region(g, cs.tg, R);
      cs.with_hull = false;
      cs.n = warp_gather(g, cs.tg, cs.rg, wc);
      unsigned todo = act;
      int pivots = 0, reused = 0, solved = 0, general = 0, outside = 0, failed = 0;
      if (cs.n >= 0) {
        Tet tet;
        bool have = false, have_seed = false;
        for (int step = 0; step < 32; ++step) {
          // boustrophedon through the 8x4 block: consecutive voxels are always neighbours, which is what the
          // tetrahedron re-use and the seeds live on
          const int i = (step & 8) ? (step ^ 7) : step;
          if (!((act >> i) & 1u)) continue;
          const double x = __shfl_sync(kFull, qx, i), y = __shfl_sync(kFull, qy, i), z = __shfl_sync(kFull, qz, i);
          const double xp = x + kEta * (ctr[0] - x), yp = y + kEta * (ctr[1] - y), zp = z + kEta * (ctr[2] - z);
          int rc = kFailed;
          if (have) {  // still inside the previous voxel's (verified) tetrahedron?
            Geo geo;
            tet_geo(tet, geo);
            double b[4];
            bary(tet, geo, xp, yp, zp, b);
            if (b[0] >= 0.0 && b[1] >= 0.0 && b[2] >= 0.0 && b[3] >= 0.0) {
              rc = kVerified;
              ++reused;
            }
          }
          if (rc != kVerified) {
            const Tet seed = tet;
            init_virtual(tet, x, y, z, Mv);
            if (have_seed) lp_seed(seed, tet, xp, yp, zp);
            set_query(wc, cs.n, xp, yp, zp);
            if (lp_run(p, cs, wc, tet, xp, yp, zp, pivots) == 0) {
              have_seed = all_real(tet);  // real vertices from this warp's cache: a good start for the next voxel
              bool inside = false;
              if (have_seed) {
                Geo geo;
                tet_geo(tet, geo);
                const double d = box_dist(cs.tg, tet.x[0] + geo.cx, tet.y[0] + geo.cy, tet.z[0] + geo.cz);
                inside = d + sqrt(geo.cc) <= R - margin;  // the sphere lies inside the gathered region: verified
              }
              if (inside) {
                rc = kVerified;
                ++solved;
              } else {
                bool clobbered = false;
                ++general;
                rc = finish_voxel(p, wc, cs, x, y, z, xp, yp, zp, R, tet, pivots, clobbered);
                if (clobbered) cs.n = warp_gather(g, cs.tg, cs.rg, wc);
              }
            } else {
              have_seed = false;
            }
          }
          have = rc == kVerified;
          if (rc == kVirtual) continue;  // no hull list: solved from scratch below
          todo &= ~(1u << i);
          if (rc == kVerified) {
            VoxelOut o;
            emit(p, tet, x, y, z, o);
            if (lane == i) mine = o;
          } else if (rc == kOutside) {
            ++outside;
          } else {
            ++failed;
          }
        }
      } else {
        stat_add(p, 3, 1);
      }
      while (todo != 0) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        const double x = __shfl_sync(kFull, qx, i), y = __shfl_sync(kFull, qy, i), z = __shfl_sync(kFull, qz, i);
        const double xp = x + kEta * (ctr[0] - x), yp = y + kEta * (ctr[1] - y), zp = z + kEta * (ctr[2] - z);
        Tet tet;
        ++general;
        const int rc = solve_general(p, wc, x, y, z, xp, yp, zp, r_first, Mv, tet, pivots);
        if (rc == kVerified) {
          VoxelOut o;
          emit(p, tet, x, y, z, o);
          if (lane == i) mine = o;
        } else if (rc == kOutside) {
          ++outside;
        } else {
          ++failed;
        }
      }
      stat_add(p, 0, solved);
      stat_add(p, 1, reused);
      stat_add(p, 2, general);
      stat_add(p, 4, outside);
      stat_add(p, 5, failed);
      stat_add(p, 6, pivots);
      if (failed > 0 && lane == 0) atomicAdd(p.err_flag, failed);  // surfaced by the host as an error
    }
  }

  if (!valid) return;
  if (mine.u != mine.u) mine.u = 0.0;  // main.py:195-199 nan_to_num
  if (mine.v != mine.v) mine.v = 0.0;
  if (mine.w != mine.w) mine.w = 0.0;
  store_out<OutT>(p.u, vox, mine.u);
  store_out<OutT>(p.v, vox, mine.v);
  store_out<OutT>(p.w, vox, mine.w);
  if (p.knn_idx) {
    // simplex vertices in ascending row order with their barycentric weights (parity tests)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int b = 0; b < 3 - a; ++b) {
        if (mine.id[b] > mine.id[b + 1]) {
          const int ti = mine.id[b]; mine.id[b] = mine.id[b + 1]; mine.id[b + 1] = ti;
          const double tb = mine.b[b]; mine.b[b] = mine.b[b + 1]; mine.b[b + 1] = tb;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      p.knn_idx[vox * 4 + j] = mine.id[j];
      p.knn_dist[vox * 4 + j] = mine.b[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Hull-candidate list.  A particle p is no hull vertex if every open octant around it holds another
// particle (a separating direction n would have n.(p' - p) > 0 for the p' in the octant of sign(n)).
// Cells make that cheap: the particles of cell (cx,cy,cz) are dominated in octant (+,sy,sz) if some
// occupied cell has cx' > cx, cy' beyond cy in direction sy and cz' beyond cz in direction sz.  With
// rowmax(cy,cz) = the largest occupied cx of a cell row, that is  T(cy,cz) = max over rows strictly
// beyond (cy,cz) of rowmax  >  cx -- a 2-D exclusive running maximum (and minimum for the -x octants).
// H = particles in cells not dominated in all eight octants; it contains every hull vertex.
__global__ void hull_row_extent_kernel(const int32_t* __restrict__ cell_start, int cnx, int nrows,
                                       int* __restrict__ rowmin, int* __restrict__ rowmax) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int64_t base = (int64_t)r * cnx;
  int lo = INT_MAX, hi = -1;
  if (cell_start[base] != cell_start[base + cnx]) {
    for (int c = 0; c < cnx; ++c)
      if (cell_start[base + c + 1] != cell_start[base + c]) { lo = c; break; }
    for (int c = cnx - 1; c >= 0; --c)
      if (cell_start[base + c + 1] != cell_start[base + c]) { hi = c; break; }
  }
  rowmin[r] = lo;
  rowmax[r] = hi;
}

// one CTA; tab[k][cz*cny+cy]: k = combo (bit 0: +y, bit 1: +z) for the maxima, 4 + combo for the minima
__global__ void hull_dominance_kernel(int cny, int cnz, const int* __restrict__ rowmin,
                                      const int* __restrict__ rowmax, int* __restrict__ tmp,
                                      int* __restrict__ tab) {
  const int nrows = cny * cnz;
  int* tmax = tmp;
  int* tmin = tmp + nrows;
  for (int combo = 0; combo < 4; ++combo) {
    const bool yp = combo & 1, zp = combo & 2;
    for (int cz = threadIdx.x; cz < cnz; cz += blockDim.x) {
      int amax = -1, amin = INT_MAX;
      for (int s = 0; s < cny; ++s) {
        const int cy = yp ? cny - 1 - s : s;  // from the far end inwards: exclusive of the row itself
        const int r = cz * cny + cy;
        tmax[r] = amax;
        tmin[r] = amin;
        amax = max(amax, rowmax[r]);
        amin = min(amin, rowmin[r]);
      }
    }
    __syncthreads();
    for (int cy = threadIdx.x; cy < cny; cy += blockDim.x) {
      int amax = -1, amin = INT_MAX;
      for (int s = 0; s < cnz; ++s) {
        const int cz = zp ? cnz - 1 - s : s;
        const int r = cz * cny + cy;
        tab[combo * nrows + r] = amax;
        tab[(4 + combo) * nrows + r] = amin;
        amax = max(amax, tmax[r]);
        amin = min(amin, tmin[r]);
      }
    }
    __syncthreads();
  }
}

// One warp walks a run of kHullRun consecutive (cell-sorted) records and keeps the undominated ones in
// order; the run's output is padded to a multiple of 32 with copies of its last record, so every chunk of
// 32 output records comes from one run and is spatially compact.
constexpr int kHullRun = 4096;
// flags != nullptr: keep record i iff flags[i] (stage 2); else the cell-dominance test (stage 1)
__global__ void __launch_bounds__(128) hull_compact_kernel(const ParticleRec* __restrict__ rec,
                                                           const int32_t* __restrict__ cid, int64_t n, int cnx,
                                                           int nrows, const int* __restrict__ tab,
                                                           const uint8_t* __restrict__ flags,
                                                           ParticleRec* __restrict__ out, int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t run = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int64_t i0 = run * kHullRun;
  if (i0 >= n) return;
  const int64_t i1 = i0 + kHullRun < n ? i0 + kHullRun : n;
  auto keep = [&](const ParticleRec& r, int64_t i) {
    if (flags != nullptr) return flags[i] != 0;
    const int c = cid[r.idx];
    const int cx = c % cnx, row = c / cnx;
    bool dominated = true;
#pragma unroll
    for (int combo = 0; combo < 4; ++combo)
      dominated = dominated && tab[combo * nrows + row] > cx && tab[(4 + combo) * nrows + row] < cx;
    return !dominated;
  };
  int total = 0;
  for (int64_t i = i0 + lane; i < i0 + kHullRun; i += 32) {
    const bool k = i < i1 && keep(rec[i], i);
    total += __popc(__ballot_sync(kFull, k));
  }
  if (total == 0) return;
  const int padded = (total + 31) & ~31;
  int base = 0;
  if (lane == 0) base = atomicAdd(count, padded);
  base = __shfl_sync(kFull, base, 0);
  int pos = 0;
  ParticleRec last = rec[i0];
  for (int64_t i = i0 + lane; i < i0 + kHullRun; i += 32) {
    ParticleRec r = last;
    bool k = false;
    if (i < i1) { r = rec[i]; k = keep(r, i); }
    const unsigned b = __ballot_sync(kFull, k);
    if (k) out[base + pos + __popc(b & ((1u << lane) - 1u))] = r;
    pos += __popc(b);
    if (b != 0) {  // remember the run's latest kept record (all lanes)
      const int src = 31 - __clz(b);
      int4 a = *reinterpret_cast<const int4*>(&r), c = *(reinterpret_cast<const int4*>(&r) + 1);
      a.x = __shfl_sync(kFull, a.x, src); a.y = __shfl_sync(kFull, a.y, src);
      a.z = __shfl_sync(kFull, a.z, src); a.w = __shfl_sync(kFull, a.w, src);
      c.x = __shfl_sync(kFull, c.x, src); c.y = __shfl_sync(kFull, c.y, src);
      c.z = __shfl_sync(kFull, c.z, src); c.w = __shfl_sync(kFull, c.w, src);
      *reinterpret_cast<int4*>(&last) = a;
      *(reinterpret_cast<int4*>(&last) + 1) = c;
    }
  }
  if (total + lane < padded) out[base + total + lane] = last;
}

// Stage 2, per particle of the stage-1 list: p is an extreme point of the cloud only if some CLOSED octant
// around it holds no other particle (if every closed octant {s_i (x_i' - x_i) >= 0} held one, any direction
// n would have n.(p' - p) >= 0 for the p' in the octant of sign(n): p is not strictly separable from the
// rest, i.e. p lies in their hull).  Closed octants also retire the interior points of flat faces (wall
// particles on a lattice plane).  The search walks cell rows outwards from p's cell and stops at the
// first hit; an octant that costs more than kRefineBudget row visits keeps p (conservative).
constexpr int kRefineBudget = 1 << 16;
__global__ void __launch_bounds__(128) hull_refine_kernel(const ParticleRec* __restrict__ list, int n1,
                                                          const ParticleRec* __restrict__ rec,
                                                          const int32_t* __restrict__ cell_start,
                                                          const int32_t* __restrict__ cid, int cnx, int cny, int cnz,
                                                          const int* __restrict__ rowmin,
                                                          const int* __restrict__ rowmax, uint8_t* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n1) return;
  const ParticleRec r = list[i];
  const int c = cid[r.idx];
  const int cx = c % cnx, cy = (c / cnx) % cny, cz = c / (cnx * cny);
  bool extreme = false;
  for (int oct = 0; oct < 8 && !extreme; ++oct) {
    const int sx = (oct & 1) ? 1 : -1, sy = (oct & 2) ? 1 : -1, sz = (oct & 4) ? 1 : -1;
    bool found = false;
    int budget = kRefineBudget;
    for (int z2 = cz; z2 >= 0 && z2 < cnz && !found && budget > 0; z2 += sz) {
      for (int y2 = cy; y2 >= 0 && y2 < cny && !found && budget > 0; y2 += sy) {
        --budget;
        const int row = z2 * cny + y2;
        if (sx > 0 ? rowmax[row] < cx : rowmin[row] > cx) continue;  // no occupied cell on that side
        const int64_t rowbase = (int64_t)row * cnx;
        const int a = sx > 0 ? cell_start[rowbase + cx] : cell_start[rowbase];
        const int b = sx > 0 ? cell_start[rowbase + cnx] : cell_start[rowbase + cx + 1];
        // nearest cells first: ascending records for +x, descending for -x
        for (int j = 0; j < b - a && !found; ++j) {
          const ParticleRec q = rec[sx > 0 ? a + j : b - 1 - j];
          const double ex = q.x - r.x, ey = q.y - r.y, ez = q.z - r.z;
          found = sx * ex >= 0.0 && sy * ey >= 0.0 && sz * ez >= 0.0 && (ex != 0.0 || ey != 0.0 || ez != 0.0);
        }
      }
    }
    if (!found) extreme = true;  // empty closed octant, or the budget ran out
  }
  keep[i] = extreme ? 1 : 0;
}

// bounding box (lo xyz, hi xyz) of every chunk of 32 list records
__global__ void hull_box_kernel(const ParticleRec* __restrict__ rec, int nchunks, double* __restrict__ box) {
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= nchunks) return;
  const ParticleRec r = rec[((size_t)ch << 5) + lane];
  double v6[6] = {r.x, r.y, r.z, -r.x, -r.y, -r.z};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 6; ++c) v6[c] = fmin(v6[c], __shfl_xor_sync(kFull, v6[c], o));
  }
  if (lane < 6) box[(size_t)ch * 6 + lane] = lane < 3 ? v6[lane] : -v6[lane];
}

}  // namespace

int ensure_hull_list(ptv_hash* h, cudaStream_t stream) {
  if (h->hull_valid) return PTV_OK;
  const int cnx = h->dims[0], cny = h->dims[1], cnz = h->dims[2];
  const int nrows = cny * cnz;
  if (h->hull_cap_rows < nrows) {
    cudaFree(h->hull_tab);
    h->hull_tab = nullptr;
    h->hull_cap_rows = 0;
    PTV_CUDA(cudaMalloc(&h->hull_tab, (size_t)(12 * (int64_t)nrows + 4) * sizeof(int)));
    h->hull_cap_rows = nrows;
  }
  const int64_t nruns = (h->n + kHullRun - 1) / kHullRun;
  const int64_t need = h->n + 32 * nruns;  // every run may pad its output up to a multiple of 32
  if (h->hull_cap < need) {
    cudaFree(h->hull_rec);
    cudaFree(h->hull_box);
    h->hull_rec = nullptr;
    h->hull_box = nullptr;
    h->hull_cap = 0;
    PTV_CUDA(cudaMalloc(&h->hull_rec, (size_t)need * sizeof(ParticleRec)));
    PTV_CUDA(cudaMalloc(&h->hull_box, (size_t)(need / 32 + 1) * 6 * sizeof(double)));
    h->hull_cap = need;
  }
  int* rowmin = h->hull_tab;
  int* rowmax = rowmin + nrows;
  int* tmp = rowmax + nrows;
  int* tab = tmp + 2 * (int64_t)nrows;
  int* count = tab + 8 * (int64_t)nrows;
  PTV_CUDA(cudaMemsetAsync(count, 0, sizeof(int), stream));
  hull_row_extent_kernel<<<(nrows + 127) / 128, 128, 0, stream>>>(h->cell_start, cnx, nrows, rowmin, rowmax);
  hull_dominance_kernel<<<1, 512, 0, stream>>>(cny, cnz, rowmin, rowmax, tmp, tab);
  hull_compact_kernel<<<(unsigned)((nruns + 3) / 4), 128, 0, stream>>>(h->rec, h->cid, h->n, cnx, nrows, tab, nullptr,
                                                                      h->hull_rec, count);
  count_launches(3);
  PTV_CUDA(cudaGetLastError());
  int host = 0;
  PTV_CUDA(cudaMemcpyAsync(&host, count, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PTV_CUDA(cudaStreamSynchronize(stream));
  h->hull_n = host;
  h->hull_list = h->hull_rec;
  if (host > 0 && tuning().hull >= 2) {
    // stage 2: particle-level test on the survivors, compacted again (same chunk structure)
    const int64_t nruns2 = ((int64_t)host + kHullRun - 1) / kHullRun;
    const int64_t need2 = (int64_t)host + 32 * nruns2;
    if (h->hull_cap2 < need2) {
      cudaFree(h->hull_rec2);
      cudaFree(h->hull_keep);
      h->hull_rec2 = nullptr;
      h->hull_keep = nullptr;
      h->hull_cap2 = 0;
      PTV_CUDA(cudaMalloc(&h->hull_rec2, (size_t)need2 * sizeof(ParticleRec)));
      PTV_CUDA(cudaMalloc(&h->hull_keep, (size_t)need2));
      h->hull_cap2 = need2;
    }
    PTV_CUDA(cudaMemsetAsync(count, 0, sizeof(int), stream));
    hull_refine_kernel<<<(host + 127) / 128, 128, 0, stream>>>(h->hull_rec, host, h->rec, h->cell_start, h->cid, cnx, cny,
                                                             cnz, rowmin, rowmax, h->hull_keep);
    hull_compact_kernel<<<(unsigned)((nruns2 + 3) / 4), 128, 0, stream>>>(h->hull_rec, h->cid, host, cnx, nrows, tab,
                                                                         h->hull_keep, h->hull_rec2, count);
    count_launches(2);
    PTV_CUDA(cudaGetLastError());
    PTV_CUDA(cudaMemcpyAsync(&host, count, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PTV_CUDA(cudaStreamSynchronize(stream));
    h->hull_n = host;
    h->hull_list = h->hull_rec2;
  }
  if (host > 0) {
    hull_box_kernel<<<(host / 32 + 3) / 4, 128, 0, stream>>>(h->hull_list, host / 32, h->hull_box);
    count_launches(1);
    PTV_CUDA(cudaGetLastError());
  }
  h->hull_valid = true;
  return PTV_OK;
}

int launch_delaunay_linear(KnnParams& p, bool f32, cudaStream_t stream) {
  constexpr int T = 128;
  if (p.qrec != nullptr) {
    p.tiles_x = (int)((p.nq + T - 1) / T);
    p.tiles_y = p.tiles_z = 1;
  } else {
    p.tiles_x = (p.nx + 7) / 8;
    p.tiles_y = (p.ny + 3) / 4;
    p.tiles_z = (p.nz + 3) / 4;
  }
  const int64_t ntiles = (int64_t)p.tiles_x * p.tiles_y * p.tiles_z;
  if (ntiles > 2147483647LL) { set_error("ptv_knn_interp: grid too large for one launch"); return PTV_ERR_INVALID; }
  const size_t smem = 4 * sizeof(WarpCache);
  void (*kern)(const KnnParams);
  if (tuning().linear_occ >= 4) kern = f32 ? delaunay_linear_kernel<float, 4> : delaunay_linear_kernel<double, 4>;
  else kern = f32 ? delaunay_linear_kernel<float, 3> : delaunay_linear_kernel<double, 3>;
  PTV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)ntiles, T, smem, stream>>>(p);
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

}  // namespace ptv
