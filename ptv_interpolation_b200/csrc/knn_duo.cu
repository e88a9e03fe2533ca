// Warp-private streaming kNN + IDW / sibson kernel ("duo": two voxels per lane) -- the production
// path for k >= 8 (replaces cKDTree.query + the NumPy weights of interpolator.py:97-122,139-153).
//
// Same selection idea as knn_stream.cu (float32 histogram of squared distances -> two float64
// thresholds E_lo < E_hi per voxel -> exact classification, never a per-voxel k-best list), but the
// unit of work is a WARP, not a CTA:
//
//   * a CTA owns an 8x8x32 region of voxels and compacts its pore voxels (4x4x4-block-major order);
//     warps take chunks of 64 consecutive pore voxels from that list (dynamic), each lane owns two;
//   * every warp scans the cell list around the bounding box of ITS 64 voxels through its own
//     4 KB staging buffer -- no block barrier after the compaction, and the scanned volume is that of
//     a 4x4x4 .. 8x8x4 box (+) R instead of an 8x8x16 one;
//   * each staged candidate is one broadcast LDS.128 that feeds both voxels of the lane; the float32
//     squared distance uses the expanded form |c|^2 - 2 q.c (+ |q|^2 folded into the threshold / bin
//     offset): 3-4 FFMA per voxel-candidate pair;
//   * histogram counters are 16 bit, laid out [bin][lane][voxel]: one 32-bit word per lane per bin, so
//     every read-modify-write is bank-conflict free and the two voxels' updates never alias;
//   * the four warps of a CTA run decoupled, so everything executed once per chunk (cell-list
//     iterator, radius estimate, thresholds, list selection) lives in small non-inlined functions that
//     keep their state in shared memory: the code the warps touch stays inside the instruction cache
//     and the hot loops keep their registers.
//
// Exactness is unchanged: the selected SET is the canonical k nearest (ties on the particle row)
// whenever n_in <= k <= n_in + n_list and every key below E_hi was scanned; this is verified per voxel,
// and a voxel for which it cannot be established is handed to the exact heap kernel (knn_interp.cu)
// through the fail list (tile granularity, de-duplicated with a bitmap).
#include "knn_common.cuh"

namespace ptv {

static constexpr int kDNB = 64;    // histogram bins over [0, Tmax)
static constexpr int kDList = 16;  // crossing-bin list capacity per voxel
static constexpr int kDCH = 32;    // records per staged chunk; two chunks are resident (one in flight)
static constexpr int kDW = 4;      // warps per CTA
static constexpr int kDT = kDW * 32;
static constexpr int kDVPT = 16;   // voxels of the region per thread (region = 8 x 8 x 32)
static constexpr int kDRZ = 32;    // z-extent of a CTA's region: ~14 chunks for 4 warps keeps the CTA's tail short
static constexpr int kDMinEstimate = 16;
static constexpr unsigned kFull = 0xffffffffu;
static constexpr int kModeIdw = 0, kModeSibson = 2;

// ---- one warp's shared memory -------------------------------------------------------------------
struct __align__(16) WarpScan {  // iterator over the cell list around the warp's voxels (warp-uniform)
  double lo[3], hi[3];           // bounding box of the warp's voxels
  double cx, cy, cz;             // its centre: staged float32 coordinates are relative to it
  double rmax;                   // radius at which the whole cell grid is covered
  double R, R2, pR2;             // current / previous scan radius
  int y0, y1, z0, z1;            // rows of the current region (cell coordinates)
  int py0, py1, pz0, pz1;        // rows of the previous one (shell scans)
  int have_prev, nslots, sb, c0, total;
  int pend_m, pend_half, cur_off;  // records / half of the chunk in flight; offset (0 / kDCH) of the current chunk
  int pad0, pad1;
};
static constexpr size_t kColBytes = (size_t)kDNB * 32 * 2 * sizeof(uint16_t);  // 8 KB: [bin][lane][voxel]
static_assert(kColBytes >= (size_t)2 * kDList * 32 * 8, "lists must fit under the histograms");
static_assert(kDList <= 16, "duo_select packs the list slot into four bits");
static constexpr size_t kOffS32 = kColBytes;
static constexpr size_t kOffS64 = kOffS32 + 2 * kDCH * sizeof(float4);
static constexpr size_t kOffVal = kOffS64 + 2 * kDCH * sizeof(ParticleRec);
template <typename OutT> struct DuoVal;
template <> struct DuoVal<float> {
  using type = float4;
  __device__ static double u(const float4& v) { return (double)v.x; }
  __device__ static double v(const float4& v) { return (double)v.y; }
  __device__ static double w(const float4& v) { return (double)v.z; }
  __device__ static float4 load(const HashGrid& g, int spos) { return __ldg(g.vals_s32 + spos); }
};
template <> struct DuoVal<double> {
  using type = Value4;
  __device__ static double u(const Value4& v) { return v.u; }
  __device__ static double v(const Value4& v) { return v.v; }
  __device__ static double w(const Value4& v) { return v.w; }
  __device__ static Value4 load(const HashGrid& g, int spos) { return g.vals_s64[spos]; }
};
template <typename OutT> struct WarpLayout {
  using ValT = typename DuoVal<OutT>::type;
  static constexpr size_t kOffSeg = kOffVal + 2 * kDCH * sizeof(ValT);
  static constexpr size_t kOffScan = kOffSeg + (72 + 2 * kDCH) * sizeof(int);
  static constexpr size_t kBytes = kOffScan + sizeof(WarpScan);
  __device__ static float4* s32(unsigned char* wb) { return reinterpret_cast<float4*>(wb + kOffS32); }
  __device__ static ParticleRec* s64(unsigned char* wb) { return reinterpret_cast<ParticleRec*>(wb + kOffS64); }
  __device__ static ValT* sval(unsigned char* wb) { return reinterpret_cast<ValT*>(wb + kOffVal); }
  __device__ static int* seg_start(unsigned char* wb) { return reinterpret_cast<int*>(wb + kOffSeg); }
  __device__ static int* seg_off(unsigned char* wb) { return reinterpret_cast<int*>(wb + kOffSeg) + 32; }
  __device__ static int* spos(unsigned char* wb) { return reinterpret_cast<int*>(wb + kOffSeg) + 72; }  // [2 * kDCH]
  __device__ static WarpScan* scan(unsigned char* wb) { return reinterpret_cast<WarpScan*>(wb + kOffScan); }
};

// estimate_radius (knn_common.cuh) for one warp; the box comes from the warp's scan state
__device__ __noinline__ double duo_estimate_radius(const HashGrid* gs, const WarpScan* sc, int r0, int k) {
  const HashGrid& g = *gs;
  const int lane = threadIdx.x & 31;
  int r = max(r0, 0);
  const int c0x = cell_of(sc->lo[0], g.ox, g.inv_cell, g.cnx), c1x = cell_of(sc->hi[0], g.ox, g.inv_cell, g.cnx);
  const int c0y = cell_of(sc->lo[1], g.oy, g.inv_cell, g.cny), c1y = cell_of(sc->hi[1], g.oy, g.inv_cell, g.cny);
  const int c0z = cell_of(sc->lo[2], g.oz, g.inv_cell, g.cnz), c1z = cell_of(sc->hi[2], g.oz, g.inv_cell, g.cnz);
  for (int attempt = 0;; ++attempt) {
    const int x0 = max(c0x - r, 0), x1 = min(c1x + r, g.cnx - 1);
    const int y0 = max(c0y - r, 0), y1 = min(c1y + r, g.cny - 1);
    const int z0 = max(c0z - r, 0), z1 = min(c1z + r, g.cnz - 1);
    const int nry = y1 - y0 + 1, nrows = nry * (z1 - z0 + 1);
    int mine = 0;
    for (int s = lane; s < nrows; s += 32) {
      const int64_t rowbase = ((int64_t)(z0 + s / nry) * g.cny + (y0 + s % nry)) * g.cnx;
      mine += g.cell_start[rowbase + x1 + 1] - g.cell_start[rowbase + x0];
    }
    const int n1 = __reduce_add_sync(kFull, mine);
    const bool whole = x0 == 0 && y0 == 0 && z0 == 0 && x1 == g.cnx - 1 && y1 == g.cny - 1 && z1 == g.cnz - 1;
    if (n1 >= kDMinEstimate) {
      const double vol = (double)(x1 - x0 + 1) * nry * (z1 - z0 + 1) * g.cell * g.cell * g.cell;
      // (3k / (4 pi rho))^(1/3); the estimate only sets the scale of the radius schedule: float32
      return (double)cbrtf((float)(0.238732414637843 * k * vol / n1));
    }
    if (attempt >= 10 || whole) return -1.0;
    r += attempt < 3 ? 1 : (r + 1) / 2;
  }
}

// Start a scan of the region of radius R around the warp's box: the whole region, or (shell) only what
// the previous region did not cover.  Lane 0 writes the iterator state; the other lanes read it back.
__device__ __noinline__ void duo_begin_scan(const HashGrid* gs, WarpScan* sc, double R, int shell) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    const HashGrid& g = *gs;
    if (shell) {
      sc->pR2 = sc->R2;
      sc->py0 = sc->y0; sc->py1 = sc->y1; sc->pz0 = sc->z0; sc->pz1 = sc->z1;
    }
    sc->have_prev = shell;
    sc->R = R;
    sc->R2 = R * R;
    sc->y0 = cell_of(sc->lo[1] - R, g.oy, g.inv_cell, g.cny);
    sc->y1 = cell_of(sc->hi[1] + R, g.oy, g.inv_cell, g.cny);
    sc->z0 = cell_of(sc->lo[2] - R, g.oz, g.inv_cell, g.cnz);
    sc->z1 = cell_of(sc->hi[2] + R, g.oz, g.inv_cell, g.cnz);
    const int nrows = (sc->y1 - sc->y0 + 1) * (sc->z1 - sc->z0 + 1);
    sc->nslots = shell ? 2 * nrows : nrows;
    sc->sb = 0;
    sc->c0 = 0;
    sc->total = 0;
    sc->pend_m = 0;
  }
  __syncwarp();
}

// Scan the current region again from the start (classification / sibson passes: whole region).
__device__ __forceinline__ void duo_restart_scan(WarpScan* sc) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    sc->have_prev = 0;
    sc->nslots = (sc->y1 - sc->y0 + 1) * (sc->z1 - sc->z0 + 1);
    sc->sb = 0;
    sc->c0 = 0;
    sc->total = 0;
    sc->pend_m = 0;
  }
  __syncwarp();
}

// Cells [xa, xb] of row (cy, cz) within distance sqrt(R2) of the box; false if the row is farther.
// Like row_interval (knn_common.cuh) with the half-width rounded UP from a float32 square root: the
// interval may only grow, and it is recomputed identically when the row is subtracted from a later shell.
__device__ __forceinline__ bool duo_row_interval(const HashGrid& g, const WarpScan* sc, double R2, int cy, int cz,
                                                 int& xa, int& xb) {
  const double ylo = g.oy + cy * g.cell, zlo = g.oz + cz * g.cell;
  const double dy = fmax(0.0, fmax(ylo - sc->hi[1], sc->lo[1] - (ylo + g.cell)));
  const double dz = fmax(0.0, fmax(zlo - sc->hi[2], sc->lo[2] - (zlo + g.cell)));
  const double rem = R2 - (dy * dy + dz * dz);
  if (rem < 0.0) return false;
  const double hx = (double)(sqrtf((float)rem) * 1.000001f);
  xa = cell_of(sc->lo[0] - hx, g.ox, g.inv_cell, g.cnx);
  xb = cell_of(sc->hi[0] + hx, g.ox, g.inv_cell, g.cnx);
  return true;
}

// Record range [start, start + cnt) of slot s of the current scan (region, or shell region \ previous).
__device__ __forceinline__ void duo_resolve_slot(const HashGrid& g, const WarpScan* sc, int s, int nslots, int& start,
                                                 int& cnt) {
  start = 0;
  cnt = 0;
  if (s >= nslots) return;
  const bool have_prev = sc->have_prev != 0;
  const int nrows_y = sc->y1 - sc->y0 + 1;
  const int row = have_prev ? (s >> 1) : s;
  const int which = have_prev ? (s & 1) : 0;
  const int cy = sc->y0 + row % nrows_y;
  const int cz = sc->z0 + row / nrows_y;
  int xa, xb;
  if (!duo_row_interval(g, sc, sc->R2, cy, cz, xa, xb)) return;
  if (have_prev) {
    int pa, pb;
    const bool in_prev = cy >= sc->py0 && cy <= sc->py1 && cz >= sc->pz0 && cz <= sc->pz1 &&
                         duo_row_interval(g, sc, sc->pR2, cy, cz, pa, pb);
    if (in_prev) {
      if (which == 0) xb = pa - 1; else xa = pb + 1;
    } else if (which == 1) {
      return;
    }
  }
  if (xa <= xb) {
    const int64_t rowbase = ((int64_t)cz * g.cny + cy) * g.cnx;
    start = g.cell_start[rowbase + xa];
    cnt = g.cell_start[rowbase + xb + 1] - start;
  }
}

__device__ __forceinline__ void duo_cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}

// Start the asynchronous copies (LDGSTS, no registers) of the records [c0, c0 + m) of the current slot batch
// into buffer half `half`: 32-byte records, their values (with_exact), and the sorted position of each.
template <typename OutT>
__device__ __forceinline__ void duo_issue_chunk(const HashGrid& g, unsigned char* wb, int c0, int m, int half,
                                                int with_exact) {
  using L = WarpLayout<OutT>;
  const int lane = threadIdx.x & 31;
  if (lane < m) {
    const int* seg_start = L::seg_start(wb);
    const int* seg_off = L::seg_off(wb);
    const int gpos = c0 + lane;
    int lo = 0, hi2 = 31;
#pragma unroll
    for (int it = 0; it < 5; ++it) {  // record gpos lives in the last segment whose offset <= gpos
      const int mid = (lo + hi2 + 1) >> 1;
      if (seg_off[mid] <= gpos) lo = mid; else hi2 = mid - 1;
    }
    const int spos = seg_start[lo] + (gpos - seg_off[lo]);
    const int j = half * kDCH + lane;
    const char* src = reinterpret_cast<const char*>(g.rec + spos);
    char* dst = reinterpret_cast<char*>(L::s64(wb) + j);
    duo_cp_async16(dst, src);
    duo_cp_async16(dst + 16, src + 16);
    if (with_exact) {
      if (sizeof(OutT) == 4) {
        duo_cp_async16(L::sval(wb) + j, g.vals_s32 + spos);
      } else {
        const char* vs = reinterpret_cast<const char*>(g.vals_s64 + spos);
        char* vd = reinterpret_cast<char*>(L::sval(wb) + j);
        duo_cp_async16(vd, vs);
        duo_cp_async16(vd + 16, vs + 16);
      }
    }
    L::spos(wb)[j] = spos;
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}

// Hand the next chunk of up to kDCH records of the scan to the caller and return its size (0 = the scan is
// over); the chunk sits at offset sc->cur_off of the staging arrays.  float32 entries are centre-relative
// x, y, z and |c|^2, padded with far-away sentinels up to kDCH; with_exact also stages the values.  While the
// caller works on this chunk the next one of the slot batch is already travelling global -> shared
// (cp.async into the other buffer half), so the staging latency is paid once per batch, not per chunk.
// Only __syncwarp(): the warps of a CTA never wait for each other.
template <typename OutT>
__device__ __noinline__ int duo_next_chunk(const HashGrid* gs, unsigned char* wb, int with_exact) {
  using L = WarpLayout<OutT>;
  const HashGrid& g = *gs;
  WarpScan* sc = L::scan(wb);
  const int lane = threadIdx.x & 31;
  __syncwarp();  // the previous chunk has been consumed by every lane
  for (;;) {
    const int pend_m = sc->pend_m, half = sc->pend_half, c0 = sc->c0, total = sc->total;
    if (pend_m > 0) {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");  // this lane's copies have landed
      const int j = half * kDCH + lane;
      float4* s32 = L::s32(wb);
      if (lane < pend_m) {
        int4* recp = reinterpret_cast<int4*>(L::s64(wb) + j);
        const int4 a = recp[0];
        const int4 c = recp[1];
        const double px = __hiloint2double(a.y, a.x), py = __hiloint2double(a.w, a.z);
        const double pz = __hiloint2double(c.y, c.x);
        const float fx = (float)(px - sc->cx), fy = (float)(py - sc->cy), fz = (float)(pz - sc->cz);
        s32[j] = make_float4(fx, fy, fz, fmaf(fz, fz, fmaf(fy, fy, fx * fx)));
        reinterpret_cast<int*>(recp)[7] = L::spos(wb)[j];  // pad word: the record's cell-sorted position
      } else {
        s32[j] = make_float4(0.0f, 0.0f, 0.0f, INFINITY);  // never accepted, lands in the last bin
      }
      const int mnext = min(kDCH, total - c0);
      if (mnext > 0) duo_issue_chunk<OutT>(g, wb, c0, mnext, half ^ 1, with_exact);
      __syncwarp();
      if (lane == 0) {
        sc->cur_off = half * kDCH;
        sc->pend_m = mnext > 0 ? mnext : 0;
        sc->pend_half = half ^ 1;
        sc->c0 = c0 + (mnext > 0 ? kDCH : 0);
      }
      __syncwarp();
      return pend_m;
    }
    const int sb = sc->sb, nslots = sc->nslots;
    if (sb >= nslots) return 0;
    // next batch of 32 row slots -> record ranges, prefix offsets; its first chunk starts travelling at once
    int start, cnt;
    duo_resolve_slot(g, sc, sb + lane, nslots, start, cnt);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += x;
    }
    const int tot = __shfl_sync(kFull, inc, 31);
    L::seg_start(wb)[lane] = start;
    L::seg_off(wb)[lane] = inc - cnt;
    if (lane == 31) L::seg_off(wb)[32] = tot;
    __syncwarp();
    const int m0 = min(kDCH, tot);
    const int h0 = (sc->cur_off / kDCH) ^ 1;  // not the half the caller may still be reading
    if (m0 > 0) duo_issue_chunk<OutT>(g, wb, 0, m0, h0, with_exact);
    __syncwarp();
    if (lane == 0) {
      sc->sb = sb + 32;
      sc->total = tot;
      sc->c0 = m0 > 0 ? kDCH : 0;
      sc->pend_m = m0;
      sc->pend_half = h0;
    }
    __syncwarp();
  }
}

// Particles of one voxel in the first `nbins` histogram bins (h points at the lane's column + voxel).
__device__ __noinline__ int duo_hist_cum(const uint16_t* h, int nbins) {
  int cum = 0;
#pragma unroll 4
  for (int b = 0; b < nbins; ++b) cum += h[b * 64];
  return cum;
}

struct DuoThr {
  double e_lo, e_hi;
  int c_below;
  int flags;  // 1 crowded crossing bin, 2 no usable crossing bin
};
// Crossing bin of one voxel -> thresholds.  rin2 < 0: the whole cell grid was scanned.
__device__ __noinline__ DuoThr duo_thresholds(const uint16_t* h, int k, double binw, double rin2) {
  DuoThr r;
  int cum = 0, bstar = -1, hb = 0;
  for (int b = 0; b < kDNB - 1; ++b) {
    hb = h[b * 64];
    if (cum + hb >= k) { bstar = b; break; }
    cum += hb;
  }
  r.c_below = cum;
  r.e_lo = bstar * binw;
  r.e_hi = (bstar + 1) * binw;
  r.flags = (bstar >= 0 && hb > kDList) ? 1 : 0;
  // the crossing bin must lie inside the scanned radius unless the whole grid was scanned
  if (bstar < 0 || (rin2 >= 0.0 && r.e_hi > rin2)) r.flags |= 2;
  return r;
}
// After the sub-bin pass over a crowded crossing bin: narrower thresholds (one sub-bin of slack on each
// side absorbs float32 fuzz at this resolution); false = still crowded (ties / coincident particles).
__device__ __noinline__ bool duo_refine(const uint16_t* h, int k, double binw, DuoThr* t) {
  const double w2 = binw / kDNB;
  int cum = t->c_below, b2 = -1, hb = 0;
  for (int b = 0; b < kDNB; ++b) {
    hb = h[b * 64];
    if (cum + hb >= k) { b2 = b; break; }
    cum += hb;
  }
  if (b2 < 0 || hb > kDList) return false;
  const double lo2 = t->e_lo + (double)max(b2 - 1, 0) * w2;
  const double hi2 = fmin(t->e_lo + (double)(b2 + 2) * w2, t->e_hi);
  t->e_lo = lo2;
  t->e_hi = hi2;
  return true;
}

// A voxel the optimistic path cannot finish: its heap tile goes to the fail list (once).
__device__ __noinline__ void duo_push_fail(const KnnParams* p, int ix, int iy, int iz, int reason) {
  const int htx = (p->nx + p->fail_tx - 1) / p->fail_tx, hty = (p->ny + p->fail_ty - 1) / p->fail_ty;
  const int tile = ((iz / p->fail_tz) * hty + iy / p->fail_ty) * htx + ix / p->fail_tx;
  const unsigned bit = 1u << (tile & 31);
  if ((atomicOr(&p->fail_flags[tile >> 5], bit) & bit) == 0u) p->fail_list[atomicAdd(p->fail_count, 1)] = tile;
  if (p->stats != nullptr) atomicAdd(&p->stats[reason], 1ULL);
}

__device__ __noinline__ void duo_count(unsigned long long* slot, unsigned long long v) { atomicAdd(slot, v); }

__device__ __noinline__ void duo_mark_solid(int64_t* idx, int k) {
  for (int j = 0; j < k; ++j) idx[j] = -1;
}

__device__ __noinline__ double duo_pow_weight(double d2, double power) {
  return 1.0 / (pow(sqrt(d2), power) + 1e-10);
}

__device__ __forceinline__ double exact_from_rows(const HashGrid& g, double qx, double qy, double qz, int row) {
  const double* pr = g.pts + (int64_t)row * 3;
  const double ex = qx - pr[0], ey = qy - pr[1], ez = qz - pr[2];
  return __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
}
// The same key from the cell-sorted record (the lists hold sorted positions: those lines were just staged,
// so the gather hits L2 / L1 instead of the original-order arrays); also returns the particle row.
__device__ __forceinline__ double exact_from_rec(const HashGrid& g, double qx, double qy, double qz, int spos, int& row) {
  const int4* src = reinterpret_cast<const int4*>(g.rec + spos);
  const int4 a = __ldg(src);
  const int4 c = __ldg(src + 1);
  row = c.z;
  const double ex = qx - __hiloint2double(a.y, a.x), ey = qy - __hiloint2double(a.w, a.z);
  const double ez = qz - __hiloint2double(c.y, c.x);
  return __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
}
__device__ __forceinline__ Value4 sorted_value(const HashGrid& g, int spos, int f32vals) {
  if (f32vals) {
    const float4 v = __ldg(g.vals_s32 + spos);
    Value4 r;
    r.u = (double)v.x; r.v = (double)v.y; r.w = (double)v.z; r.pad = 0.0;
    return r;
  }
  return g.vals_s64[spos];
}

// The `nd` smallest (d2, row) entries of one voxel's crossing-bin list, moved to the front of the list.
// Keys are float32 offsets from E_lo (monotone in the exact key); entries whose offsets coincide across the
// cut (exact ties, or keys closer than the packed float32 resolves) are ordered by the exact key computed
// from the cell-sorted records.
__device__ __noinline__ void duo_select(const HashGrid* gs, float* lk, int* li, int nl, int nd, double qx, double qy,
                                        double qz) {
  // selection on packed integers: the offsets are non-negative floats, so their bit patterns order like the
  // values; the four lowest mantissa bits give way to the list slot, which makes every round one
  // load + integer-min per entry (keys that differ only in those bits count as tied, see below)
  auto packed = [&](int j) { return (int)((__float_as_uint(lk[j * 32]) & ~15u) | (unsigned)j); };
  for (int i = 0; i < nd; ++i) {
    int m = packed(i);
    for (int j = i + 1; j < nl; ++j) m = min(m, packed(j));
    const int best = m & 15;
    if (best != i) {
      const float bk = lk[best * 32];
      const int bi = li[best * 32];
      lk[best * 32] = lk[i * 32];
      li[best * 32] = li[i * 32];
      lk[i * 32] = bk;
      li[i * 32] = bi;
    }
  }
  if (nd <= 0 || nd >= nl) return;
  const unsigned cut = __float_as_uint(lk[(nd - 1) * 32]) & ~15u;
  unsigned rest = 0xffffffffu;
  for (int j = nd; j < nl; ++j) rest = min(rest, __float_as_uint(lk[j * 32]) & ~15u);
  if (rest != cut) return;
  auto tied = [&](int j) { return (__float_as_uint(lk[j * 32]) & ~15u) == cut; };
  int first = nd - 1;
  while (first > 0 && tied(first - 1)) --first;
  // entries [first, nl) tied with the cut compete for the slots [first, nd): exact (d2, row) order
  for (int i = first; i < nd; ++i) {
    int best = -1, bi = 0;
    double bk = 0.0;
    for (int j = i; j < nl; ++j) {
      if (!tied(j)) continue;
      int ij;
      const double kj = exact_from_rec(*gs, qx, qy, qz, li[j * 32], ij);
      if (best < 0 || key_greater(bk, bi, kj, ij)) { best = j; bk = kj; bi = ij; }
    }
    if (best != i) {
      const float tk = lk[i * 32];
      const int ti = li[i * 32];
      lk[i * 32] = lk[best * 32];
      li[i * 32] = li[best * 32];
      lk[best * 32] = tk;
      li[best * 32] = ti;
    }
  }
}

struct DuoAcc {
  double a, b, c, d;
};
// IDW: weights of the first nd list entries -> (sum w, sum w u, sum w v, sum w w).
__device__ __noinline__ DuoAcc duo_list_idw(const HashGrid* gs, float* lk, int* li, int nl, int nd, double qx, double qy,
                                            double qz, double power, double e_lo, int f32vals, int64_t* dbg) {
  DuoAcc r = {0.0, 0.0, 0.0, 0.0};
  duo_select(gs, lk, li, nl, nd, qx, qy, qz);
  if (f32vals && dbg == nullptr && power == 2.0) {
    // float32 output: E_lo + offset is the key to ~3e-9 relative, far inside what the float32 weight keeps,
    // so only the value is gathered (one 16-byte line the stager touched a moment ago)
#pragma unroll 2
    for (int i = 0; i < nd; ++i) {
      const float4 val = __ldg(gs->vals_s32 + li[i * 32]);
      float wf;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(wf) : "f"((float)(e_lo + (double)lk[i * 32]) + 1e-10f));
      const double wgt = (double)wf;
      r.a += wgt;
      r.b += wgt * (double)val.x;
      r.c += wgt * (double)val.y;
      r.d += wgt * (double)val.z;
    }
    return r;
  }
  for (int i = 0; i < nd; ++i) {
    const int spos = li[i * 32];
    int row;
    const double d2 = exact_from_rec(*gs, qx, qy, qz, spos, row);
    if (dbg != nullptr) dbg[i] = row;
    const double wgt = power == 2.0 ? 1.0 / (d2 + 1e-10) : duo_pow_weight(d2, power);
    const Value4 val = sorted_value(*gs, spos, f32vals);
    r.a += wgt;
    r.b += wgt * val.u;
    r.c += wgt * val.v;
    r.d += wgt * val.w;
  }
  return r;
}
// sibson, first pass: moments of the distances about dshift -> (sum dd, sum dd^2).
__device__ __noinline__ DuoAcc duo_list_moments(const HashGrid* gs, float* lk, int* li, int nl, int nd, double qx,
                                                double qy, double qz, double dshift, int64_t* dbg) {
  DuoAcc r = {0.0, 0.0, 0.0, 0.0};
  duo_select(gs, lk, li, nl, nd, qx, qy, qz);
  for (int i = 0; i < nd; ++i) {
    int row;
    const double dd = sqrt(exact_from_rec(*gs, qx, qy, qz, li[i * 32], row)) - dshift;
    if (dbg != nullptr) dbg[i] = row;
    r.a += dd;
    r.b += dd * dd;
  }
  return r;
}
// sibson, second pass: w = (1/(d+eps)) * exp(-d / (std(d) + eps))  (interpolator.py:102-122)
__device__ __noinline__ DuoAcc duo_list_sibson(const HashGrid* gs, const int* li, int nd, double qx, double qy,
                                               double qz, double inv_s, int f32vals) {
  DuoAcc r = {0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < nd; ++i) {
    const int spos = li[i * 32];
    int row;
    const double d = sqrt(exact_from_rec(*gs, qx, qy, qz, spos, row));
    const double wgt = (1.0 / (d + 1e-10)) * exp(-d * inv_s);
    const Value4 val = sorted_value(*gs, spos, f32vals);
    r.a += wgt;
    r.b += wgt * val.u;
    r.c += wgt * val.v;
    r.d += wgt * val.w;
  }
  return r;
}

// Sub-bin histogram of the crossing bins of the crowded voxels (rare: voxels far from every particle,
// whose neighbours all sit at nearly the same distance).
template <typename OutT>
__device__ __noinline__ void duo_subbin_pass(const HashGrid* gs, unsigned char* wb, uint16_t* hist, float ax0, float ay0,
                                             float az0, float b0, float ax1, float ay1, float az1, float b1,
                                             float inv_w, float lo0, float lo1, int cr0, int cr1) {
  using L = WarpLayout<OutT>;
  const float4* s32 = L::s32(wb);
  duo_restart_scan(L::scan(wb));
  for (;;) {
    const int m = duo_next_chunk<OutT>(gs, wb, 0);
    if (m == 0) break;
    const int off = L::scan(wb)->cur_off;
    for (int j = off; j < off + kDCH; ++j) {
      const float4 c = s32[j];
      const float f0 = fmaf(ax0, c.x, fmaf(ay0, c.y, fmaf(az0, c.z, fmaf(c.w, inv_w, b0))));
      const float f1 = fmaf(ax1, c.x, fmaf(ay1, c.y, fmaf(az1, c.z, fmaf(c.w, inv_w, b1))));
      const float r0 = (f0 - lo0) * (float)kDNB, r1 = (f1 - lo1) * (float)kDNB;
      if (cr0 && r0 >= 0.0f && r0 < (float)kDNB) hist[__float2int_rz(r0) * 64] += 1;
      if (cr1 && r1 >= 0.0f && r1 < (float)kDNB) hist[__float2int_rz(r1) * 64 + 1] += 1;
    }
  }
}

// kDiag: work counters ("stats" tuning) and neighbour-row output for the parity tests; the production
// instantiation carries neither.
template <typename OutT, int kMode, bool kDiag>
__global__ void __launch_bounds__(kDT, sizeof(OutT) == 4 ? 4 : 3) knn_duo_kernel(const KnnParams p) {
  using ValT = typename DuoVal<OutT>::type;
  using L = WarpLayout<OutT>;
  constexpr bool kF32 = sizeof(OutT) == 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  unsigned char* wb = smem_raw + (size_t)wid * L::kBytes;
  uint16_t* hist = reinterpret_cast<uint16_t*>(wb) + lane * 2;  // + b * 64 + voxel
  float* lkey = reinterpret_cast<float*>(wb) + lane;            // + (voxel * kDList + i) * 32
  int* lidx = reinterpret_cast<int*>(wb + (size_t)2 * kDList * 32 * 4) + lane;
  const float4* s32 = L::s32(wb);
  const ParticleRec* s64 = L::s64(wb);
  const ValT* sval = L::sval(wb);
  WarpScan* sc = L::scan(wb);
  unsigned char* cbase = smem_raw + (size_t)kDW * L::kBytes;
  KnnParams* pp = reinterpret_cast<KnnParams*>(cbase);  // the kernel parameters, for the non-inlined helpers
  const HashGrid* gs = &pp->g;
  uint16_t* vlist = reinterpret_cast<uint16_t*>(cbase + ((sizeof(KnnParams) + 15) & ~(size_t)15));  // [kDVPT * kDT]
  int* warp_tot = reinterpret_cast<int*>(vlist + kDVPT * kDT);                        // [kDW + 1]
  int* next_chunk = warp_tot + kDW + 1;
  unsigned long long* wcnt = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(next_chunk + 1) + 7) & ~(uintptr_t)7);  // [8] stats mode

  const int k = p.k;
  const bool kStats = kDiag && p.stats != nullptr;
  const bool dbg = kDiag && p.knn_idx != nullptr;
  if (kStats && t < 8) wcnt[t] = 0ULL;
  if (t == 0) {
    *pp = p;
    *next_chunk = 0;
  }

  // ---- region of 8 x 8 x 32 voxels (4x4x4 blocks, block-major) -> compact list of its pore voxels.
  //      A thread owns two x-rows of four voxels of one block.
  constexpr int RX = 8, RY = 8, RZ = kDRZ, NBX = RX / 4, NBY = RY / 4;
  static_assert(RX * RY * RZ == kDVPT * kDT && kDVPT % 4 == 0 && kDVPT <= 32, "region holds kDVPT voxels per thread");
  const int region = blockIdx.x;
  const int rx = region % p.tiles_x;
  const int ry = (region / p.tiles_x) % p.tiles_y;
  const int rz = region / (p.tiles_x * p.tiles_y);
  auto decode = [&](int i, int& ix, int& iy, int& iz) {
    const int b = i >> 6, l = i & 63;
    ix = rx * RX + (b % NBX) * 4 + (l & 3);
    iy = ry * RY + ((b / NBX) % NBY) * 4 + ((l >> 2) & 3);
    iz = rz * RZ + (b / (NBX * NBY)) * 4 + (l >> 4);
  };
  int nact;
  {
    int mine = 0;
    unsigned flags = 0;
    // rows of four voxels are 4-byte (mask) / 16-byte (float32 output) aligned
    const bool vec_ok = (p.nx & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mask) & 3) == 0 &&
                        ((reinterpret_cast<uintptr_t>(p.u) | reinterpret_cast<uintptr_t>(p.v) |
                          reinterpret_cast<uintptr_t>(p.w)) & 15) == 0;
#pragma unroll
    for (int r = 0; r < kDVPT / 4; ++r) {
      int ix, iy, iz;
      decode(kDVPT * t + 4 * r, ix, iy, iz);
      if (iy >= p.ny || iz >= p.nz || ix >= p.nx) continue;
      const int64_t vox0 = ((int64_t)iz * p.ny + iy) * p.nx + ix;
      unsigned m4;  // one byte per voxel of the row, non-zero = pore
      if (p.mask == nullptr) {
        m4 = 0x01010101u;
      } else if (vec_ok) {
        m4 = *reinterpret_cast<const unsigned*>(p.mask + vox0);
      } else {
        m4 = 0u;
        for (int q = 0; q < 4; ++q)
          if (ix + q < p.nx) m4 |= (unsigned)p.mask[vox0 + q] << (8 * q);
      }
      const int nvalid = min(4, p.nx - ix);
      if (kF32 && vec_ok && m4 == 0u && !dbg) {  // a solid row: main.py:202-207 writes zero
        const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.u) + vox0) = z4;
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.v) + vox0) = z4;
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.w) + vox0) = z4;
        continue;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q >= nvalid) continue;
        if ((m4 >> (8 * q)) & 0xffu) {
          flags |= 1u << (4 * r + q);
          ++mine;
        } else {
          const int64_t vox = vox0 + q;
          store_out<OutT>(p.u, vox, 0.0);
          store_out<OutT>(p.v, vox, 0.0);
          store_out<OutT>(p.w, vox, 0.0);
          if (dbg) duo_mark_solid(p.knn_idx + vox * k, k);
        }
      }
    }
    const int off = block_scan_excl<kDT>(mine, warp_tot, &nact);
    int o = off;
#pragma unroll
    for (int q = 0; q < kDVPT; ++q)
      if (flags & (1u << q)) vlist[o++] = (uint16_t)(kDVPT * t + q);
  }
  if (nact == 0) return;
  __syncthreads();  // the last block barrier: from here on every warp works alone
  const int nchunks = (nact + 63) >> 6;
  const HashGrid& g = *gs;

  for (;;) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(next_chunk, 1);
    chunk = __shfl_sync(kFull, chunk, 0);
    if (chunk >= nchunks) break;
    const int cbeg = chunk * 64, ccnt = min(64, nact - cbeg);

    // ---- the lane's two voxels
    bool ok[2];  // active and still on the optimistic path
    int vcode[2];
    double qx[2], qy[2], qz[2];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      ok[v] = lane + 32 * v < ccnt;
      vcode[v] = ok[v] ? vlist[cbeg + lane + 32 * v] : 0;
      int ix, iy, iz;
      decode(vcode[v], ix, iy, iz);
      qx[v] = ok[v] ? p.ax[ix] : 0.0;
      qy[v] = ok[v] ? p.ay[iy] : 0.0;
      qz[v] = ok[v] ? p.az[iz] : 0.0;
    }
    auto voxel_index = [&](int v) {
      int ix, iy, iz;
      decode(vcode[v], ix, iy, iz);
      return ((int64_t)iz * p.ny + iy) * p.nx + ix;
    };
    // a voxel the optimistic path cannot finish -> its heap tile goes to the fail list (once)
    auto fail_voxel = [&](int v, int reason) {
      if (!ok[v]) return;
      ok[v] = false;
      int ix, iy, iz;
      decode(vcode[v], ix, iy, iz);
      duo_push_fail(pp, ix, iy, iz, reason);
    };
    // ---- bounding box of the warp's voxels, radius that covers the whole cell grid
    {
      // float32 bounds rounded outwards (the box may only grow), reduced as order-preserving integer keys
      auto key_of = [](float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; };
      auto float_of = [](int kk) { return __int_as_float(kk >= 0 ? kk : kk ^ 0x7fffffff); };
      int klo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, khi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
#pragma unroll
      for (int v = 0; v < 2; ++v)
        if (ok[v]) {
          klo[0] = min(klo[0], key_of(__double2float_rd(qx[v]))); khi[0] = max(khi[0], key_of(__double2float_ru(qx[v])));
          klo[1] = min(klo[1], key_of(__double2float_rd(qy[v]))); khi[1] = max(khi[1], key_of(__double2float_ru(qy[v])));
          klo[2] = min(klo[2], key_of(__double2float_rd(qz[v]))); khi[2] = max(khi[2], key_of(__double2float_ru(qz[v])));
        }
      double v6[6];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        v6[c] = (double)float_of(__reduce_min_sync(kFull, klo[c]));
        v6[c + 3] = -(double)float_of(__reduce_max_sync(kFull, khi[c]));
      }
      __syncwarp();
      if (lane == 0) {
        const double glo[3] = {g.ox, g.oy, g.oz};
        const double ghi[3] = {g.ox + g.cnx * g.cell, g.oy + g.cny * g.cell, g.oz + g.cnz * g.cell};
        double r2 = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          sc->lo[c] = v6[c];
          sc->hi[c] = -v6[c + 3];
          const double mm = fmax(fmax(-v6[c + 3] - glo[c], ghi[c] - v6[c]), 0.0);
          r2 += mm * mm;
        }
        sc->cx = 0.5 * (v6[0] - v6[3]);
        sc->cy = 0.5 * (v6[1] - v6[4]);
        sc->cz = 0.5 * (v6[2] - v6[5]);
        sc->rmax = sqrt(r2) * (1.0 + 1e-9) + 1e-3 * g.cell;
        sc->R2 = 0.0;
        sc->cur_off = 0;
        sc->pend_m = 0;
      }
      __syncwarp();
    }

    // ---- local density -> radius schedule and histogram scale
    const double r_est = duo_estimate_radius(gs, sc, p.r0, k);
    if (!(r_est > 0.0)) {  // nothing to estimate a scale from (deep void / tiny cloud): exact kernel
      fail_voxel(0, 1);
      fail_voxel(1, 1);
      continue;
    }
    // three scan radii whose squares sit just above histogram bin edges 24, 40 and 64 (= Tmax)
    constexpr int kEdges[3] = {24, 40, kDNB};
    const double binw = p.rscale * r_est * r_est / kEdges[0];
    const float inv_w = (float)(1.0 / binw);
    float qf[2][3];  // centre-relative query
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      qf[v][0] = (float)(qx[v] - sc->cx);
      qf[v][1] = (float)(qy[v] - sc->cy);
      qf[v][2] = (float)(qz[v] - sc->cz);
    }

    // ---- phase A: grow the scanned region, float32 histogram of squared distances
    {
      uint32_t* hw = reinterpret_cast<uint32_t*>(hist);
#pragma unroll 8
      for (int b = 0; b < kDNB; ++b) hw[b * 32] = 0u;
    }
    bool finished = false;
    DuoThr th[2];  // thresholds of the last stage scanned
    {
      float qa[2][3], qb[2];  // -2 q / binw and |q|^2 / binw: bin index = qa.c + |c|^2 / binw + qb
#pragma unroll
      for (int v = 0; v < 2; ++v) {
#pragma unroll
        for (int c = 0; c < 3; ++c) qa[v][c] = -2.0f * qf[v][c] * inv_w;
        qb[v] = (qf[v][0] * qf[v][0] + qf[v][1] * qf[v][1] + qf[v][2] * qf[v][2]) * inv_w;
      }
#pragma unroll 1
      for (int stage = 0; stage < 3; ++stage) {
        double R = sqrt((kEdges[stage] + 0.02) * binw) + 1e-6 * g.cell;
        const bool last = R >= sc->rmax;
        if (last) R = sc->rmax;
        duo_begin_scan(gs, sc, R, stage > 0 ? 1 : 0);
        int staged = 0;
        for (;;) {
          const int m = duo_next_chunk<OutT>(gs, wb, 0);
          if (m == 0) break;
          staged += m;
          const int off = sc->cur_off;
#pragma unroll 1
          for (int j0 = off; j0 < off + kDCH; j0 += 4) {
            // all loads and bin indices of four candidates first: the counter updates below are the only
            // dependent chain left (shared-memory stores may alias the staged candidates for the compiler).
            // Four, not eight: the smaller body is worth 2 % (253.8 vs 259.1 ms at config 4; two: 257.9) --
            // the decoupled warps of an SM partition share a 6 KB L0 instruction cache
            int b0[4], b1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 c = s32[j0 + i];
              const float f0 = fmaf(qa[0][0], c.x, fmaf(qa[0][1], c.y, fmaf(qa[0][2], c.z, fmaf(c.w, inv_w, qb[0]))));
              const float f1 = fmaf(qa[1][0], c.x, fmaf(qa[1][1], c.y, fmaf(qa[1][2], c.z, fmaf(c.w, inv_w, qb[1]))));
              b0[i] = min(kDNB - 1, __float2int_rz(f0)) * 64;
              b1[i] = min(kDNB - 1, __float2int_rz(f1)) * 64 + 1;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint16_t x0 = hist[b0[i]], x1 = hist[b1[i]];  // the two voxels' counters never share bytes
              hist[b0[i]] = x0 + 1;
              hist[b1[i]] = x1 + 1;
            }
          }
        }
        if (kStats && lane == 0) duo_count(&wcnt[0], (unsigned long long)staged * ccnt);
        // stop test: >= k particles in bins that lie entirely inside the scanned radius -- the same scan that
        // finds the crossing bin (its upper edge inside the scanned radius <=> that many bins are complete)
        bool done = true;
        {
          const double rin = sc->R - 1e-6 * g.cell;
          const double rin2 = last ? -1.0 : rin * rin;
#pragma unroll
          for (int v = 0; v < 2; ++v)
            if (ok[v]) {
              th[v] = duo_thresholds(hist + v, k, binw, rin2);
              done = done && !(th[v].flags & 2);
            }
        }
        if (__all_sync(kFull, done) || last) {
          finished = true;
          if (kStats && lane == 0 && stage > 0) duo_count(&p.stats[4 + stage], 1ULL);
          break;
        }
      }
    }
    // ---- thresholds from the crossing bin (found by the last stage's scan).  A voxel whose crossing bin is
    //      not inside the scanned radius -- only possible when the three stages were not enough: the k-th
    //      neighbour is beyond the histogram range -- goes to the exact kernel; the others are classified on
    //      what was scanned.
    double e_lo[2] = {0.0, 0.0}, e_hi[2] = {0.0, 0.0};
    {
      bool crowded[2] = {false, false};
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        if (ok[v]) {
          if (th[v].flags & 2) fail_voxel(v, finished ? 3 : 2);
          crowded[v] = ok[v] && (th[v].flags & 1);
        }
      }
      if (!__any_sync(kFull, ok[0] || ok[1])) continue;
      // ---- phase A2 (rare): a crossing bin that holds more particles than the short list is
      //      re-histogrammed with kDNB sub-bins; phase B still verifies the result exactly
      if (__any_sync(kFull, crowded[0] || crowded[1])) {
        __syncwarp();
        uint32_t* hw = reinterpret_cast<uint32_t*>(hist);
        for (int b = 0; b < kDNB; ++b) hw[b * 32] = 0u;
        const float q20 = (qf[0][0] * qf[0][0] + qf[0][1] * qf[0][1] + qf[0][2] * qf[0][2]) * inv_w;
        const float q21 = (qf[1][0] * qf[1][0] + qf[1][1] * qf[1][1] + qf[1][2] * qf[1][2]) * inv_w;
        duo_subbin_pass<OutT>(gs, wb, hist, -2.0f * qf[0][0] * inv_w, -2.0f * qf[0][1] * inv_w, -2.0f * qf[0][2] * inv_w,
                              q20, -2.0f * qf[1][0] * inv_w, -2.0f * qf[1][1] * inv_w, -2.0f * qf[1][2] * inv_w, q21,
                              inv_w, crowded[0] ? (float)(th[0].e_lo / binw) : 0.0f,
                              crowded[1] ? (float)(th[1].e_lo / binw) : 0.0f, crowded[0] ? 1 : 0, crowded[1] ? 1 : 0);
#pragma unroll
        for (int v = 0; v < 2; ++v)
          if (crowded[v] && !duo_refine(hist + v, k, binw, &th[v])) fail_voxel(v, 3);
      }
#pragma unroll
      for (int v = 0; v < 2; ++v)
        if (ok[v]) { e_lo[v] = th[v].e_lo; e_hi[v] = th[v].e_hi; }
    }
    if (!__any_sync(kFull, ok[0] || ok[1])) continue;
    // slab hash: the neighbours are certified only if every key below E_hi belongs to a binned particle,
    // i.e. the k-th distance stays inside the binned z-range; otherwise the caller redoes the frame on the
    // full hash (pipeline.hot_path_step)
    if (g.clip_lo > -INFINITY || g.clip_hi < INFINITY) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        if (ok[v]) {
          const double dmin = fmin(qz[v] - g.clip_lo, g.clip_hi - qz[v]);
          if (!(dmin > 0.0 && e_hi[v] <= dmin * dmin)) atomicAdd(p.clip_count, 1);
        }
      }
    }
    __syncwarp();  // the histograms are dead: the columns now hold the lists

    // ---- phase B: exact classification.  Every key below E_hi has to be seen, nothing more: the region shrinks
    //      to the box (+) the largest sqrt(E_hi) of the warp's voxels (never larger than what phase A scanned)
    {
      float rb = 0.0f;
#pragma unroll
      for (int v = 0; v < 2; ++v)
        if (ok[v]) rb = fmaxf(rb, __double2float_ru(sqrt(e_hi[v])));
      rb = __int_as_float(__reduce_max_sync(kFull, __float_as_int(rb)));  // non-negative floats order like ints
      const double RB = fmin((double)rb * (1.0 + 1e-6) + 1e-6 * g.cell, sc->R);
      duo_begin_scan(gs, sc, RB, 0);
    }
    // float32 error of the expanded squared distance: coordinates are below `half` in magnitude
    const float half = (float)(sc->R + fmax(sc->hi[0] - sc->lo[0], fmax(sc->hi[1] - sc->lo[1], sc->hi[2] - sc->lo[2])) +
                               g.cell) * 1.000001f;
    const float ec = half * 2.4e-7f;         // 2 ulp of the largest coordinate
    const float en = 4.0e-6f * half * half;  // rounding of |q|^2, |c|^2 and q.c (each <= 3 half^2)
    // float32 limit that no exact key below e can exceed (all margins rounded generously upwards)
    auto lim32 = [&](double e) {
      const float ef = (float)e;
      return (ef + 16.0f * sqrtf(ef) * ec + 64.0f * ec * ec + en) * (1.0f + 1e-5f) + ef * 3e-7f;
    };
    float qc[2][3], thr[2];  // -2 q (centre-relative) and limit - |q|^2
    int n_in[2] = {0, 0}, n_l[2] = {0, 0};
    double wsum[2] = {0.0, 0.0}, su[2] = {0.0, 0.0}, sv[2] = {0.0, 0.0}, sw[2] = {0.0, 0.0};
    // sibson moments about a shift close to the distances themselves (the crossing-bin edge)
    double dshift[2] = {0.0, 0.0};
    int64_t vox[2] = {0, 0};
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const float q2 = qf[v][0] * qf[v][0] + qf[v][1] * qf[v][1] + qf[v][2] * qf[v][2];
#pragma unroll
      for (int c = 0; c < 3; ++c) qc[v][c] = -2.0f * qf[v][c];
      thr[v] = ok[v] ? lim32(e_hi[v]) - q2 : -INFINITY;
      if (kMode == kModeSibson) dshift[v] = sqrt(e_lo[v]);
      if (dbg) vox[v] = voxel_index(v);
    }
    const bool p2 = p.power == 2.0;
    unsigned nexact = 0u;
    int staged_b = 0;
    for (;;) {
      const int m = duo_next_chunk<OutT>(gs, wb, 1);
      if (m == 0) break;
      staged_b += m;
      const int base = sc->cur_off;
      {
        unsigned mk[2] = {0u, 0u};
#pragma unroll 1
        for (int j8 = 0; j8 < 32; j8 += 4) {  // small body: the decoupled warps share the instruction cache
          unsigned m0 = 0u, m1 = 0u;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float4 c = s32[base + j8 + jj];
            const float t0 = fmaf(qc[0][0], c.x, fmaf(qc[0][1], c.y, fmaf(qc[0][2], c.z, c.w)));
            const float t1 = fmaf(qc[1][0], c.x, fmaf(qc[1][1], c.y, fmaf(qc[1][2], c.z, c.w)));
            m0 |= (t0 <= thr[0] ? 1u : 0u) << jj;
            m1 |= (t1 <= thr[1] ? 1u : 0u) << jj;
          }
          mk[0] |= m0 << j8;
          mk[1] |= m1 << j8;
        }
        if (kStats) nexact += __popc(mk[0]) + __popc(mk[1]);
        // float32 output: the weights of one group of candidates are summed in float32 (<= 32 terms), the
        // groups in float64
        float gw[2] = {0.0f, 0.0f}, gu[2] = {0.0f, 0.0f}, gv[2] = {0.0f, 0.0f}, gq[2] = {0.0f, 0.0f};
        // each lane walks only ITS accepted candidates, one of each voxel per trip; straight-line code so
        // that the two independent evaluations overlap (a voxel that has run out re-reads slot `base` with
        // every effect masked)
        while ((mk[0] | mk[1]) != 0u) {
          bool on[2], in[2];
          int j[2], sp[2];
          double d2[2];
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            on[v] = mk[v] != 0u;
            j[v] = base + (on[v] ? __ffs((int)mk[v]) - 1 : 0);
            mk[v] &= mk[v] - 1u;  // 0 stays 0
          }
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const double2 xy = *reinterpret_cast<const double2*>(&s64[j[v]].x);
            const int4 zi = *reinterpret_cast<const int4*>(&s64[j[v]].z);  // z, particle row, sorted position
            sp[v] = zi.w;
            const double ex = qx[v] - xy.x, ey = qy[v] - xy.y, ez = qz[v] - __hiloint2double(zi.y, zi.x);
            d2[v] = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
            in[v] = on[v] && d2[v] < e_lo[v];
            if (dbg && in[v] && n_in[v] < k) p.knn_idx[vox[v] * k + n_in[v]] = zi.z;
            n_in[v] += in[v] ? 1 : 0;
          }
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            if (kMode == kModeSibson) {
              const double dd = in[v] ? sqrt(d2[v]) - dshift[v] : 0.0;
              wsum[v] += dd;  // moments first; the weights need the std of all k distances
              su[v] += dd * dd;
            } else if (kF32) {
              float wgt;
              if (p2) {  // the weight only needs float32 accuracy: MUFU.RCP, 1 ulp
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(wgt) : "f"((float)d2[v] + 1e-10f));
              } else {
                wgt = in[v] ? (float)duo_pow_weight(d2[v], p.power) : 0.0f;
              }
              wgt = in[v] ? wgt : 0.0f;
              const float4 val = *reinterpret_cast<const float4*>(&sval[j[v]]);
              gw[v] += wgt;
              gu[v] = fmaf(wgt, val.x, gu[v]);
              gv[v] = fmaf(wgt, val.y, gv[v]);
              gq[v] = fmaf(wgt, val.z, gq[v]);
            } else {
              double wgt = 0.0;
              if (in[v]) wgt = p2 ? 1.0 / (d2[v] + 1e-10) : duo_pow_weight(d2[v], p.power);
              const ValT val = sval[j[v]];
              wsum[v] += wgt;
              su[v] += wgt * DuoVal<OutT>::u(val);
              sv[v] += wgt * DuoVal<OutT>::v(val);
              sw[v] += wgt * DuoVal<OutT>::w(val);
            }
          }
          // crossing-bin candidates -> the voxel's list (float32 offset from E_lo, sorted position)
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            if (on[v] && !in[v] && d2[v] < e_hi[v]) {
              if (n_l[v] < kDList) {
                lkey[(v * kDList + n_l[v]) * 32] = (float)(d2[v] - e_lo[v]);
                lidx[(v * kDList + n_l[v]) * 32] = sp[v];
              }
              ++n_l[v];  // > kDList: overflow
            }
          }
        }
        if (kF32 && kMode != kModeSibson) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            wsum[v] += (double)gw[v];
            su[v] += (double)gu[v];
            sv[v] += (double)gv[v];
            sw[v] += (double)gq[v];
          }
        }
      }
    }
    if (kStats) {
      nexact = __reduce_add_sync(kFull, nexact);
      if (lane == 0) {
        duo_count(&wcnt[1], (unsigned long long)staged_b * ccnt);
        duo_count(&wcnt[2], (unsigned long long)nexact);
      }
    }

    // ---- the `need` smallest (d2, row) of the list complete the k nearest
    int need[2];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      need[v] = k - n_in[v];
      if (ok[v] && (n_l[v] > kDList || need[v] < 0 || need[v] > n_l[v])) fail_voxel(v, 4);
      if (!ok[v]) continue;
      if (kStats) duo_count(&wcnt[3], (unsigned long long)n_l[v]);
      int64_t* dptr = dbg ? p.knn_idx + vox[v] * k + n_in[v] : nullptr;
      if (kMode == kModeSibson) {
        const DuoAcc a = duo_list_moments(gs, lkey + v * kDList * 32, lidx + v * kDList * 32, n_l[v], need[v], qx[v], qy[v],
                                          qz[v], dshift[v], dptr);
        wsum[v] += a.a;
        su[v] += a.b;
      } else {
        const DuoAcc a = duo_list_idw(gs, lkey + v * kDList * 32, lidx + v * kDList * 32, n_l[v], need[v], qx[v], qy[v],
                                      qz[v], p.power, e_lo[v], kF32 ? 1 : 0, dptr);
        wsum[v] += a.a;
        su[v] += a.b;
        sv[v] += a.c;
        sw[v] += a.d;
      }
    }

    if (kMode == kModeSibson) {
      // interpolator.py:102-122: w = (1/(d+eps)) * exp(-d / (std(d) + eps)), normalised
      const double eps = 1e-10;
      double inv_s[2];
      float thr_lo[2];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const double mean = wsum[v] / k;
        const double var = fmax(su[v] / k - mean * mean, 0.0);
        inv_s[v] = 1.0 / (sqrt(var) + eps);
        const float q2 = qf[v][0] * qf[v][0] + qf[v][1] * qf[v][1] + qf[v][2] * qf[v][2];
        thr_lo[v] = ok[v] ? lim32(e_lo[v]) - q2 : -INFINITY;
        wsum[v] = su[v] = sv[v] = sw[v] = 0.0;
        if (ok[v]) {
          const DuoAcc a = duo_list_sibson(gs, lidx + v * kDList * 32, need[v], qx[v], qy[v], qz[v], inv_s[v], kF32 ? 1 : 0);
          wsum[v] = a.a;
          su[v] = a.b;
          sv[v] = a.c;
          sw[v] = a.d;
        }
      }
      duo_restart_scan(sc);
      int staged_c = 0;
      for (;;) {
        const int m = duo_next_chunk<OutT>(gs, wb, 1);
        if (m == 0) break;
        staged_c += m;
        const int base = sc->cur_off;
        {
          unsigned mk[2] = {0u, 0u};
#pragma unroll 1
          for (int j8 = 0; j8 < 32; j8 += 8) {
            unsigned m0 = 0u, m1 = 0u;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const float4 c = s32[base + j8 + jj];
              const float t0 = fmaf(qc[0][0], c.x, fmaf(qc[0][1], c.y, fmaf(qc[0][2], c.z, c.w)));
              const float t1 = fmaf(qc[1][0], c.x, fmaf(qc[1][1], c.y, fmaf(qc[1][2], c.z, c.w)));
              m0 |= (t0 <= thr_lo[0] ? 1u : 0u) << jj;
              m1 |= (t1 <= thr_lo[1] ? 1u : 0u) << jj;
            }
            mk[0] |= m0 << j8;
            mk[1] |= m1 << j8;
          }
          while ((mk[0] | mk[1]) != 0u) {
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              if (mk[v] != 0u) {
                const int j = base + __ffs((int)mk[v]) - 1;
                mk[v] &= mk[v] - 1u;
                const double2 xy = *reinterpret_cast<const double2*>(&s64[j].x);
                const double zz = s64[j].z;
                const double ex = qx[v] - xy.x, ey = qy[v] - xy.y, ez = qz[v] - zz;
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
                if (d2 < e_lo[v]) {
                  const ValT val = sval[j];
                  const double d = sqrt(d2);
                  const double wgt = (1.0 / (d + eps)) * exp(-d * inv_s[v]);
                  wsum[v] += wgt;
                  su[v] += wgt * DuoVal<OutT>::u(val);
                  sv[v] += wgt * DuoVal<OutT>::v(val);
                  sw[v] += wgt * DuoVal<OutT>::w(val);
                }
              }
            }
          }
        }
      }
      if (kStats && lane == 0) duo_count(&wcnt[1], (unsigned long long)staged_c * ccnt);
    }

#pragma unroll
    for (int v = 0; v < 2; ++v) {
      if (ok[v]) {
        double ou, ov, ow;
        if (kF32) {  // the result is rounded to float32 anyway: float32 quotient of the float64 sums
          const float rw = 1.0f / (float)wsum[v];
          ou = (double)((float)su[v] * rw); ov = (double)((float)sv[v] * rw); ow = (double)((float)sw[v] * rw);
        } else {
          ou = su[v] / wsum[v]; ov = sv[v] / wsum[v]; ow = sw[v] / wsum[v];
        }
        // main.py:195-199 nan_to_num
        if (ou != ou) ou = 0.0;
        if (ov != ov) ov = 0.0;
        if (ow != ow) ow = 0.0;
        const int64_t vx = voxel_index(v);
        store_out<OutT>(p.u, vx, ou);
        store_out<OutT>(p.v, vx, ov);
        store_out<OutT>(p.w, vx, ow);
      }
    }
    if (kStats) {
      const int nok = __popc(__ballot_sync(kFull, ok[0])) + __popc(__ballot_sync(kFull, ok[1]));
      if (lane == 0) {
        duo_count(&wcnt[4], (unsigned long long)nok);
        duo_count(&wcnt[5], 1ULL);
      }
    }
    __syncwarp();  // the columns and the staging buffers are reused by the next chunk
  }

  if (kStats) {
    __syncthreads();
    if (t == 0) {
      for (int i = 0; i < 6; ++i) duo_count(&p.stats[8 + i], wcnt[i]);
      duo_count(&p.stats[0], 1ULL);
    }
  }
}

// Debug / parity output: the streaming kernel wrote the selected particle rows of every pore voxel in
// arbitrary order; put them into the canonical order (exact d2, row) and fill in the distances, like
// cKDTree.query after canonicalisation (oracle/reference_port.py knn_canonical).  Tiles redone by the
// heap kernel are already sorted; sorting them again changes nothing.
__global__ void knn_sort_lists_kernel(const KnnParams p) {
  const int64_t nvox = (int64_t)p.nx * p.ny * p.nz;
  const int64_t vox = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vox >= nvox) return;
  const int k = p.k;
  int64_t* idx = p.knn_idx + vox * k;
  double* dst = p.knn_dist + vox * k;
  if (idx[0] < 0) {  // solid voxel
    for (int j = 0; j < k; ++j) dst[j] = nan("");
    return;
  }
  const int ix = (int)(vox % p.nx), iy = (int)((vox / p.nx) % p.ny), iz = (int)(vox / ((int64_t)p.nx * p.ny));
  const double qx = p.ax[ix], qy = p.ay[iy], qz = p.az[iz];
  for (int j = 0; j < k; ++j) dst[j] = exact_from_rows(p.g, qx, qy, qz, (int)idx[j]);
  for (int i = 1; i < k; ++i) {  // insertion sort by (d2, row)
    const double kd = dst[i];
    const int64_t ki = idx[i];
    int j = i - 1;
    while (j >= 0 && (dst[j] > kd || (dst[j] == kd && idx[j] > ki))) {
      dst[j + 1] = dst[j];
      idx[j + 1] = idx[j];
      --j;
    }
    dst[j + 1] = kd;
    idx[j + 1] = ki;
  }
  for (int j = 0; j < k; ++j) dst[j] = sqrt(dst[j]);
}

template <typename OutT>
static size_t duo_smem_bytes() {
  const size_t b = (size_t)kDW * WarpLayout<OutT>::kBytes + ((sizeof(KnnParams) + 15) & ~(size_t)15) +
                   (size_t)kDVPT * kDT * sizeof(uint16_t) + (kDW + 2) * sizeof(int) + 8 + 8 * sizeof(unsigned long long);
  return (b + 15) & ~(size_t)15;
}

template <typename OutT, int kMode, bool kDiag>
static int launch_duo_t(KnnParams& p, cudaStream_t stream) {
  p.tiles_x = (p.nx + 7) / 8;
  p.tiles_y = (p.ny + 7) / 8;
  p.tiles_z = (p.nz + kDRZ - 1) / kDRZ;
  const int64_t nreg = (int64_t)p.tiles_x * p.tiles_y * p.tiles_z;
  if (nreg > 2147483647LL) { set_error("ptv_knn_interp: grid too large for one launch"); return PTV_ERR_INVALID; }
  const size_t smem = duo_smem_bytes<OutT>();
  auto kern = knn_duo_kernel<OutT, kMode, kDiag>;
  PTV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)nreg, kDT, smem, stream>>>(p);
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

template <bool kDiag>
static int launch_duo_d(KnnParams& p, bool f32, cudaStream_t stream) {
  if (p.method == PTV_METHOD_SIBSON)
    return f32 ? launch_duo_t<float, kModeSibson, kDiag>(p, stream) : launch_duo_t<double, kModeSibson, kDiag>(p, stream);
  return f32 ? launch_duo_t<float, kModeIdw, kDiag>(p, stream) : launch_duo_t<double, kModeIdw, kDiag>(p, stream);
}

int launch_knn_duo(KnnParams& p, bool f32, cudaStream_t stream) {
  return (p.stats != nullptr || p.knn_idx != nullptr) ? launch_duo_d<true>(p, f32, stream)
                                                      : launch_duo_d<false>(p, f32, stream);
}

int launch_knn_sort_lists(KnnParams& p, cudaStream_t stream) {
  const int64_t nvox = (int64_t)p.nx * p.ny * p.nz;
  knn_sort_lists_kernel<<<(unsigned)((nvox + 127) / 128), 128, 0, stream>>>(p);
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

}  // namespace ptv
