// Declarations shared by the two fused kNN kernels (heap: knn_interp.cu, streaming: knn_stream.cu).
#pragma once
#include <math.h>

#include "ptv_internal.cuh"

namespace ptv {

static constexpr int kStageCap = 256;  // particle records per staging chunk (8 KB)

struct KnnParams {
  HashGrid g;
  const double* ax;
  const double* ay;
  const double* az;
  int nx, ny, nz;
  const uint8_t* mask;
  int method;
  int k;
  double power;
  void* u;
  void* v;
  void* w;
  int64_t* knn_idx;
  double* knn_dist;
  int tiles_x, tiles_y, tiles_z;
  int r0;
  double rscale;  // stream kernel: first scan radius^2 = rscale * r_est^2
  double smoothing;
  int rbf_kernel;  // 0 thin_plate_spline, 1 cubic, 2 linear, 3 quintic
  int rbf_npoly;   // monomials of the polynomial tail: 4 (degree 1), 1 (degree 0), 10 (degree 2)
  int* err_flag;
  int* clip_count;  // slab hash: voxels whose certified radius reaches outside the binned z-range
  // stream kernel -> heap kernel hand-off: tiles the optimistic kernel could not finish
  int* fail_list;        // [capacity]
  int* fail_count;       // [1]
  unsigned* fail_flags;  // duo kernel: one bit per heap tile, so a tile enters the fail list once
  int fail_tx, fail_ty, fail_tz;  // shape of the heap kernel's tiles the fail list is expressed in
  const int* tile_list;  // heap kernel: process only these tiles (NULL = all tiles of the grid)
  const int* tile_count;
  unsigned long long* stats;  // optional counters (tiles, failed tiles, candidates, accepted)
  // point-query mode (heap kernel): queries are the cell-sorted records of a second hash (or of the
  // particle hash itself for self-queries); outputs are indexed by the query's original row
  const ParticleRec* qrec;
  int64_t nq;
  // outlier filter (filtering.py:5-58): keep flag and distance to the k-th neighbour per particle
  uint8_t* keep;
  double* kth_dist;
  double mad_threshold;
  // method='linear' (delaunay_linear.cu): particles that may be convex-hull vertices (NULL = scan all)
  const ParticleRec* hull_rec;  // chunks of 32 records
  const double* hull_box;       // [hull_n / 32][6] bounding box of each chunk
  int hull_n;
};

__device__ __forceinline__ bool key_greater(double ka, int ia, double kb, int ib) {
  return ka > kb || (ka == kb && ia > ib);
}

template <int T>
__device__ __forceinline__ int block_scan_excl(int v, int* warp_tot, int* total) {
  constexpr int NW = T / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int w2 = 0; w2 < NW; ++w2) {
    const int x = warp_tot[w2];
    if (w2 < wid) woff += x;
    tot += x;
  }
  __syncthreads();
  *total = tot;
  return woff + inc - v;
}

__device__ __forceinline__ int cell_of(double p, double o, double inv_cell, int n) {
  // round-down conversion saturates at INT_MIN / INT_MAX (p may be far outside the grid; NaN -> 0)
  return min(max(__double2int_rd((p - o) * inv_cell), 0), n - 1);
}

template <typename OutT>
__device__ __forceinline__ void store_out(void* base, int64_t i, double v) {
  reinterpret_cast<OutT*>(base)[i] = (OutT)v;
}



// ------------------------------------------------------------------------------------------
// Scan geometry shared by both kernels.  A tile scans the cells within distance R of the bounding
// box of its active voxels ("rounded box": exact per cell row in y/z, sqrt-clipped along x), then
// grows R and scans only the new shell.  Every particle within R of ANY point of the bounding box
// lies in a scanned cell, so for a voxel inside the box "k-th distance < R" proves the search is
// complete -- the same exactness argument as a ring walk, at ~half the scanned volume.
struct TileGeom {
  double lo[3], hi[3];  // bounding box of the tile's active voxels
  double rmax;          // radius at which the whole cell grid is covered
};

struct RoundRegion {
  double R, R2;
  int y0, y1, z0, z1;  // bounding rectangle of rows (cell coordinates)
};

__device__ __forceinline__ RoundRegion make_region(const HashGrid& g, const TileGeom& tg, double R) {
  RoundRegion rg;
  rg.R = R;
  rg.R2 = R * R;
  rg.y0 = cell_of(tg.lo[1] - R, g.oy, g.inv_cell, g.cny);
  rg.y1 = cell_of(tg.hi[1] + R, g.oy, g.inv_cell, g.cny);
  rg.z0 = cell_of(tg.lo[2] - R, g.oz, g.inv_cell, g.cnz);
  rg.z1 = cell_of(tg.hi[2] + R, g.oz, g.inv_cell, g.cnz);
  return rg;
}

// Cells [xa, xb] of row (cy, cz) within the region; false if the row is farther than R.
__device__ __forceinline__ bool row_interval(const HashGrid& g, const TileGeom& tg, const RoundRegion& rg, int cy,
                                             int cz, int& xa, int& xb) {
  const double ylo = g.oy + cy * g.cell, zlo = g.oz + cz * g.cell;
  const double dy = fmax(0.0, fmax(ylo - tg.hi[1], tg.lo[1] - (ylo + g.cell)));
  const double dz = fmax(0.0, fmax(zlo - tg.hi[2], tg.lo[2] - (zlo + g.cell)));
  const double rem = rg.R2 - (dy * dy + dz * dz);
  if (rem < 0.0) return false;
  const double hx = sqrt(rem);
  xa = cell_of(tg.lo[0] - hx, g.ox, g.inv_cell, g.cnx);
  xb = cell_of(tg.hi[0] + hx, g.ox, g.inv_cell, g.cnx);
  return true;
}

__device__ __forceinline__ int region_slots(const RoundRegion& rg, bool have_prev) {
  const int nrows = (rg.y1 - rg.y0 + 1) * (rg.z1 - rg.z0 + 1);
  return have_prev ? 2 * nrows : nrows;
}

// Record range [start, start+cnt) of slot s of the shell  region(rg) \ region(prev).
__device__ __forceinline__ void resolve_slot(const HashGrid& g, const TileGeom& tg, const RoundRegion& rg,
                                             const RoundRegion& prev, bool have_prev, int s, int nslots, int& start,
                                             int& cnt) {
  start = 0;
  cnt = 0;
  if (s >= nslots) return;
  const int nrows_y = rg.y1 - rg.y0 + 1;
  const int row = have_prev ? (s >> 1) : s;
  const int which = have_prev ? (s & 1) : 0;
  const int cy = rg.y0 + row % nrows_y;
  const int cz = rg.z0 + row / nrows_y;
  int xa, xb;
  if (!row_interval(g, tg, rg, cy, cz, xa, xb)) return;
  if (have_prev) {
    int pa, pb;
    const bool in_prev = cy >= prev.y0 && cy <= prev.y1 && cz >= prev.z0 && cz <= prev.z1 &&
                         row_interval(g, tg, prev, cy, cz, pa, pb);
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

struct ScanSmem {
  ParticleRec* stage64;  // [kStageCap] exact records (may be NULL if unused)
  float4* stage32;       // [kStageCap] tile-centre-relative float32 x,y,z (may be NULL if unused)
  void* stage_val;       // [kStageCap] float4 (kVal == 1) or Value4 (kVal == 2) values, cell-sorted order
  int* seg_start;        // [T]
  int* seg_off;          // [T+1]
  int* warp_tot;         // [NW]
};

// Stage the shell  region(rg) \ region(prev)  chunk by chunk through shared memory and call body(m)
// on each staged chunk of m records.  Returns the number of records staged.
template <int T, int CAP, bool kWith32, bool kWith64, int kVal, bool kPad32, typename F>
__device__ __forceinline__ int scan_shell(const HashGrid& g, const TileGeom& tg, const RoundRegion& rg,
                                          const RoundRegion& prev, bool have_prev, const ScanSmem& sm, double cx,
                                          double cy, double cz, F&& body) {
  const int t = threadIdx.x;
  const int nslots = region_slots(rg, have_prev);
  int staged = 0;
  for (int slot_base = 0; slot_base < nslots; slot_base += T) {
    int start, cnt;
    resolve_slot(g, tg, rg, prev, have_prev, slot_base + t, nslots, start, cnt);
    int total;
    const int off = block_scan_excl<T>(cnt, sm.warp_tot, &total);
    sm.seg_start[t] = start;
    sm.seg_off[t] = off;
    if (t == 0) sm.seg_off[T] = total;
    __syncthreads();
    staged += total;
    for (int chunk0 = 0; chunk0 < total; chunk0 += CAP) {
      const int m = min(CAP, total - chunk0);
      for (int j = t; j < m; j += T) {
        const int gpos = chunk0 + j;
        int lo = 0, hi2 = T - 1;
        while (lo < hi2) {  // record j of the chunk lives in the last segment whose offset <= gpos
          const int mid = (lo + hi2 + 1) >> 1;
          if (sm.seg_off[mid] <= gpos) lo = mid; else hi2 = mid - 1;
        }
        const int spos = sm.seg_start[lo] + (gpos - sm.seg_off[lo]);
        const ParticleRec* src = g.rec + spos;
        if (kVal == 1) reinterpret_cast<float4*>(sm.stage_val)[j] = __ldg(g.vals_s32 + spos);
        if (kVal == 2) {
          const int4* vs = reinterpret_cast<const int4*>(g.vals_s64 + spos);
          int4* vd = reinterpret_cast<int4*>(reinterpret_cast<Value4*>(sm.stage_val) + j);
          vd[0] = __ldg(vs);
          vd[1] = __ldg(vs + 1);
        }
        const int4 a = __ldg(reinterpret_cast<const int4*>(src));
        const int4 c = __ldg(reinterpret_cast<const int4*>(src) + 1);
        if (kWith64) {
          int4* dst = reinterpret_cast<int4*>(sm.stage64 + j);
          dst[0] = a;
          dst[1] = c;
        }
        if (kWith32) {
          const double px = __hiloint2double(a.y, a.x), py = __hiloint2double(a.w, a.z);
          const double pz = __hiloint2double(c.y, c.x);
          sm.stage32[j] = make_float4((float)(px - cx), (float)(py - cy), (float)(pz - cz), 0.0f);
        }
      }
      if (kPad32) {  // pad to a multiple of 64 with far-away sentinels so consumers can run fixed trip counts
        for (int j = m + t; j < ((m + 63) & ~63); j += T) sm.stage32[j] = make_float4(1e30f, 1e30f, 1e30f, 0.0f);
      }
      __syncthreads();
      body(m);
      __syncthreads();
    }
  }
  return staged;
}

// ---- pipelined variant used by the streaming kernel ---------------------------------------------
// Two staging buffers; the records (and values) of chunk c+1 travel global -> shared with cp.async
// (LDGSTS, no registers) while chunk c is being scanned, so the global-memory latency of staging is
// hidden and each chunk costs one block barrier instead of two.
static constexpr int kPipeCap = 64;  // records per pipelined chunk

struct PipeBuf {
  ParticleRec* stage64;  // [kPipeCap]
  float4* stage32;       // [kPipeCap] tile-centre-relative float32 coordinates (converted on arrival)
  void* stage_val;       // [kPipeCap] float4 / Value4
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <int T, int kVal, typename F>
__device__ __forceinline__ int scan_shell_pipe(const HashGrid& g, const TileGeom& tg, const RoundRegion& rg,
                                               const RoundRegion& prev, bool have_prev, const PipeBuf (&buf)[2],
                                               int* seg_start, int* seg_off, int* warp_tot, double cx, double cy,
                                               double cz, F&& body) {
  const int t = threadIdx.x;
  const int nslots = region_slots(rg, have_prev);
  int staged = 0;
  for (int slot_base = 0; slot_base < nslots; slot_base += T) {
    int start, cnt;
    resolve_slot(g, tg, rg, prev, have_prev, slot_base + t, nslots, start, cnt);
    int total;
    const int off = block_scan_excl<T>(cnt, warp_tot, &total);
    seg_start[t] = start;
    seg_off[t] = off;
    if (t == 0) seg_off[T] = total;
    __syncthreads();
    staged += total;
    const int nchunks = (total + kPipeCap - 1) / kPipeCap;
    auto issue = [&](int c) {  // asynchronous copies of chunk c into buffer c & 1
      const PipeBuf& b = buf[c & 1];
      const int m = min(kPipeCap, total - c * kPipeCap);
      for (int j = t; j < m; j += T) {
        const int gpos = c * kPipeCap + j;
        int lo = 0, hi2 = T - 1;
        while (lo < hi2) {  // record gpos lives in the last segment whose offset <= gpos
          const int mid = (lo + hi2 + 1) >> 1;
          if (seg_off[mid] <= gpos) lo = mid; else hi2 = mid - 1;
        }
        const int spos = seg_start[lo] + (gpos - seg_off[lo]);
        const char* src = reinterpret_cast<const char*>(g.rec + spos);
        char* dst = reinterpret_cast<char*>(b.stage64 + j);
        cp_async16(dst, src);
        cp_async16(dst + 16, src + 16);
        if (kVal == 1) cp_async16(reinterpret_cast<float4*>(b.stage_val) + j, g.vals_s32 + spos);
        if (kVal == 2) {
          const char* vs = reinterpret_cast<const char*>(g.vals_s64 + spos);
          char* vd = reinterpret_cast<char*>(reinterpret_cast<Value4*>(b.stage_val) + j);
          cp_async16(vd, vs);
          cp_async16(vd + 16, vs + 16);
        }
      }
    };
    if (nchunks > 0) issue(0);
    for (int c = 0; c < nchunks; ++c) {
      const PipeBuf& b = buf[c & 1];
      const int m = min(kPipeCap, total - c * kPipeCap);
      cp_async_commit_wait_all();  // this thread's copies of chunk c have landed
      for (int j = t; j < ((m + 63) & ~63); j += T) {
        if (j < m) {
          const ParticleRec r = b.stage64[j];
          b.stage32[j] = make_float4((float)(r.x - cx), (float)(r.y - cy), (float)(r.z - cz), 0.0f);
        } else {
          b.stage32[j] = make_float4(1e30f, 1e30f, 1e30f, 0.0f);  // far-away sentinel padding
        }
      }
      __syncthreads();
      if (c + 1 < nchunks) issue(c + 1);
      body(b, m);
    }
    __syncthreads();  // seg_* and the buffers are rewritten by the next batch
  }
  return staged;
}

// Bounding box of the tile's active voxels (block reduction through `red`[6*NW]) and the radius
// that covers the whole cell grid.
template <int T>
__device__ __forceinline__ void tile_geometry(const HashGrid& g, bool active, double qx, double qy, double qz,
                                              double* red, TileGeom& tg) {
  constexpr int NW = T / 32;
  const int t = threadIdx.x;
  double v6[6];
  v6[0] = active ? qx : INFINITY;
  v6[1] = active ? qy : INFINITY;
  v6[2] = active ? qz : INFINITY;
  v6[3] = active ? -qx : INFINITY;
  v6[4] = active ? -qy : INFINITY;
  v6[5] = active ? -qz : INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 6; ++c) v6[c] = fmin(v6[c], __shfl_xor_sync(0xffffffffu, v6[c], o));
  }
  if ((t & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 6; ++c) red[c * NW + (t >> 5)] = v6[c];
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double a = red[c * NW], b = red[(c + 3) * NW];
#pragma unroll
    for (int w2 = 1; w2 < NW; ++w2) {
      a = fmin(a, red[c * NW + w2]);
      b = fmin(b, red[(c + 3) * NW + w2]);
    }
    tg.lo[c] = a;
    tg.hi[c] = -b;
  }
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

// Particles in the cells overlapping the tile's box widened by r0 cells -> local density -> the
// radius expected to hold k particles.  Returns r_est (or a negative value if fewer than
// `min_count` particles are nearby and no estimate can be made).
template <int T>
__device__ __forceinline__ double estimate_radius(const HashGrid& g, const TileGeom& tg, int r0, int k, int min_count,
                                                  int* warp_tot) {
  const int t = threadIdx.x;
  int r = max(r0, 0);
  for (int attempt = 0;; ++attempt) {
    int b0[3], b1[3];
    b0[0] = max(cell_of(tg.lo[0], g.ox, g.inv_cell, g.cnx) - r, 0);
    b1[0] = min(cell_of(tg.hi[0], g.ox, g.inv_cell, g.cnx) + r, g.cnx - 1);
    b0[1] = max(cell_of(tg.lo[1], g.oy, g.inv_cell, g.cny) - r, 0);
    b1[1] = min(cell_of(tg.hi[1], g.oy, g.inv_cell, g.cny) + r, g.cny - 1);
    b0[2] = max(cell_of(tg.lo[2], g.oz, g.inv_cell, g.cnz) - r, 0);
    b1[2] = min(cell_of(tg.hi[2], g.oz, g.inv_cell, g.cnz) + r, g.cnz - 1);
    const int nry = b1[1] - b0[1] + 1, nrows = nry * (b1[2] - b0[2] + 1);
    int mine = 0;
    for (int s = t; s < nrows; s += T) {
      const int64_t rowbase = ((int64_t)(b0[2] + s / nry) * g.cny + (b0[1] + s % nry)) * g.cnx;
      mine += g.cell_start[rowbase + b1[0] + 1] - g.cell_start[rowbase + b0[0]];
    }
    int n1;
    (void)block_scan_excl<T>(mine, warp_tot, &n1);
    const bool whole = b0[0] == 0 && b0[1] == 0 && b0[2] == 0 && b1[0] == g.cnx - 1 && b1[1] == g.cny - 1 &&
                       b1[2] == g.cnz - 1;
    if (n1 >= min_count) {
      const double vol = (double)(b1[0] - b0[0] + 1) * nry * (b1[2] - b0[2] + 1) * g.cell * g.cell * g.cell;
      return cbrt(0.238732414637843 * k * vol / n1);  // (3k / (4 pi rho))^(1/3)
    }
    if (attempt >= 10 || whole) return -1.0;
    r += attempt < 3 ? 1 : (r + 1) / 2;  // 1, 2, 3, 4, 6, 9, 14, ... cells: voids (e.g. voxels inside grains)
  }
}

// launchers defined next to their kernels
int launch_knn_heap(KnnParams& p, int T, bool f32, cudaStream_t stream);
int launch_knn_stream(KnnParams& p, int T, bool f32, cudaStream_t stream);
int launch_knn_duo(KnnParams& p, bool f32, cudaStream_t stream);
int launch_knn_sort_lists(KnnParams& p, cudaStream_t stream);
size_t knn_heap_smem_bytes(int T, int k, int method);
int launch_delaunay_linear(KnnParams& p, bool f32, cudaStream_t stream);
int ensure_hull_list(ptv_hash* h, cudaStream_t stream);

}  // namespace ptv
