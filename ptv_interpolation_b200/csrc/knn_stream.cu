// Streaming (heap-free) fused kNN + IDW / sibson kernel -- the production path for k >= 8.
//
// Same mapping as the heap kernel (one CTA = one voxel tile, one thread = one voxel, cell-list
// rings staged through shared memory), but the per-voxel k-best list is never materialised:
//
//   phase A  walk the rings; every thread histograms the float32 squared distances of the staged
//            particles into 32 bins (bin width from the tile's local particle density).  Stop when
//            every voxel has >= k particles closer than the scanned box's nearest face (exact
//            criterion, conservative in float32).  The bin where the cumulative count crosses k
//            gives two float64 thresholds E_lo < E_hi per voxel.
//   phase B  rescan the final box.  A float32 pre-test rejects far particles; the rest get the exact
//            float64 key d2 = (dx*dx + dy*dy) + dz*dz.  Keys below E_lo are certainly among the k
//            nearest and are accumulated on the fly; keys in [E_lo, E_hi) go to a short list
//            (<= 20 entries) from which the k - n_in smallest by (d2, index) are taken.
//   phase C  (sibson only) one more rescan to apply weights that need the std of the k distances.
//
// The selected SET is exactly the canonical k nearest whenever n_in <= k <= n_in + n_list; any tile
// where that cannot be established (histogram overflow, list overflow, too few particles near the
// tile) is appended to a fail list and redone by the exact heap kernel (knn_interp.cu), so results
// never depend on the optimistic path succeeding.  Shared memory per thread drops from 12*k bytes
// to 240 bytes (5 CTAs/SM instead of 2 at k = 50) and the divergent heap maintenance disappears.
#include "knn_common.cuh"

namespace ptv {

static constexpr int kNB = 32;        // histogram bins over [0, Tmax)
static constexpr int kListCap = 20;   // capacity of the crossing-bin list
static constexpr int kMinEstimate = 16;

struct Box {
  int x0, x1, y0, y1, z0, z1;
  __device__ bool operator==(const Box& o) const {
    return x0 == o.x0 && x1 == o.x1 && y0 == o.y0 && y1 == o.y1 && z0 == o.z0 && z1 == o.z1;
  }
};

__device__ __forceinline__ Box make_box(const int c0[3], const int c1[3], int r, const HashGrid& g) {
  Box b;
  b.x0 = max(c0[0] - r, 0); b.x1 = min(c1[0] + r, g.cnx - 1);
  b.y0 = max(c0[1] - r, 0); b.y1 = min(c1[1] + r, g.cny - 1);
  b.z0 = max(c0[2] - r, 0); b.z1 = min(c1[2] + r, g.cnz - 1);
  return b;
}

// Record range [start, start+cnt) of slot s of the shell  box \ prev  (prev ignored if !have_prev).
__device__ __forceinline__ void resolve_slot(const HashGrid& g, const Box& b, const Box& pb, bool have_prev, int s,
                                             int nslots, int& start, int& cnt) {
  start = 0;
  cnt = 0;
  if (s >= nslots) return;
  const int nrows_y = b.y1 - b.y0 + 1;
  const int row = have_prev ? (s >> 1) : s;
  const int which = have_prev ? (s & 1) : 0;
  const int cy = b.y0 + row % nrows_y;
  const int cz = b.z0 + row / nrows_y;
  int xa, xb;
  if (!have_prev || cy < pb.y0 || cy > pb.y1 || cz < pb.z0 || cz > pb.z1) {
    xa = which == 0 ? b.x0 : 1;
    xb = which == 0 ? b.x1 : 0;
  } else if (which == 0) {
    xa = b.x0;
    xb = pb.x0 - 1;
  } else {
    xa = pb.x1 + 1;
    xb = b.x1;
  }
  if (xa <= xb) {
    const int64_t rowbase = ((int64_t)cz * g.cny + cy) * g.cnx;
    start = g.cell_start[rowbase + xa];
    cnt = g.cell_start[rowbase + xb + 1] - start;
  }
}

struct StreamSmem {
  float4* stage32;       // [kStageCap] tile-centre-relative float32 x,y,z (+ unused)
  ParticleRec* stage64;  // [kStageCap] exact records
  int* seg_start;        // [T]
  int* seg_off;          // [T+1]
  int* warp_tot;         // [NW]
};

// Stage the shell  box \ prev  chunk by chunk and call body(m) on each staged chunk of m records.
template <int T, bool kWith64, typename F>
__device__ __forceinline__ void scan_shell(const HashGrid& g, const Box& b, const Box& pb, bool have_prev,
                                           const StreamSmem& sm, double cx, double cy, double cz, F&& body) {
  const int t = threadIdx.x;
  const int nrows = (b.y1 - b.y0 + 1) * (b.z1 - b.z0 + 1);
  const int nslots = have_prev ? 2 * nrows : nrows;
  for (int slot_base = 0; slot_base < nslots; slot_base += T) {
    int start, cnt;
    resolve_slot(g, b, pb, have_prev, slot_base + t, nslots, start, cnt);
    int total;
    const int off = block_scan_excl<T>(cnt, sm.warp_tot, &total);
    sm.seg_start[t] = start;
    sm.seg_off[t] = off;
    if (t == 0) sm.seg_off[T] = total;
    __syncthreads();
    for (int chunk0 = 0; chunk0 < total; chunk0 += kStageCap) {
      const int m = min(kStageCap, total - chunk0);
      for (int j = t; j < m; j += T) {
        const int gpos = chunk0 + j;
        int lo = 0, hi2 = T - 1;
        while (lo < hi2) {
          const int mid = (lo + hi2 + 1) >> 1;
          if (sm.seg_off[mid] <= gpos) lo = mid; else hi2 = mid - 1;
        }
        const ParticleRec* src = g.rec + (sm.seg_start[lo] + (gpos - sm.seg_off[lo]));
        const int4 a = __ldg(reinterpret_cast<const int4*>(src));
        const int4 c = __ldg(reinterpret_cast<const int4*>(src) + 1);
        if (kWith64) {
          int4* dst = reinterpret_cast<int4*>(sm.stage64 + j);
          dst[0] = a;
          dst[1] = c;
        }
        const double px = __hiloint2double(a.y, a.x), py = __hiloint2double(a.w, a.z);
        const double pz = __hiloint2double(c.y, c.x);
        sm.stage32[j] = make_float4((float)(px - cx), (float)(py - cy), (float)(pz - cz), 0.0f);
      }
      __syncthreads();
      body(m);
      __syncthreads();
    }
  }
}

template <int T, int TX, int TY, int TZ, typename OutT>
__global__ void __launch_bounds__(T) knn_stream_kernel(const KnnParams p) {
  static_assert(TX * TY * TZ == T, "tile shape");
  constexpr int NW = T / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // per-thread columns: histogram (phase A) aliases the crossing-bin list (phase B)
  double* lkey_all = reinterpret_cast<double*>(smem_raw);                       // [kListCap][T]
  int* lidx_all = reinterpret_cast<int*>(lkey_all + (size_t)kListCap * T);      // [kListCap][T]
  int* hist_all = reinterpret_cast<int*>(smem_raw);                             // [kNB][T] (alias)
  static_assert(kNB * 4 <= kListCap * 8, "histogram must fit under the list keys");
  ParticleRec* stage64 = reinterpret_cast<ParticleRec*>(lidx_all + (size_t)kListCap * T);
  float4* stage32 = reinterpret_cast<float4*>(stage64 + kStageCap);
  double* red = reinterpret_cast<double*>(stage32 + kStageCap);                 // [6][NW]
  int* seg_start = reinterpret_cast<int*>(red + 6 * NW);
  int* seg_off = seg_start + T;
  int* warp_tot = seg_off + T + 1;
  StreamSmem sm{stage32, stage64, seg_start, seg_off, warp_tot};

  const int t = threadIdx.x;
  const int k = p.k;
  const HashGrid& g = p.g;
  const int tile = blockIdx.x;
  const int tx = tile % p.tiles_x;
  const int ty = (tile / p.tiles_x) % p.tiles_y;
  const int tz = tile / (p.tiles_x * p.tiles_y);
  const int ix = tx * TX + (t % TX);
  const int iy = ty * TY + ((t / TX) % TY);
  const int iz = tz * TZ + (t / (TX * TY));
  const bool valid = ix < p.nx && iy < p.ny && iz < p.nz;
  const int64_t vox = valid ? ((int64_t)iz * p.ny + iy) * p.nx + ix : 0;
  const bool active = valid && (p.mask == nullptr || p.mask[vox] != 0);

  if (!__syncthreads_or(active ? 1 : 0)) {  // tile entirely solid / outside: zero fill
    if (valid) {
      store_out<OutT>(p.u, vox, 0.0);
      store_out<OutT>(p.v, vox, 0.0);
      store_out<OutT>(p.w, vox, 0.0);
    }
    return;
  }
  const double qx = valid ? p.ax[ix] : 0.0;
  const double qy = valid ? p.ay[iy] : 0.0;
  const double qz = valid ? p.az[iz] : 0.0;

  // ---- bounding box of the tile's active voxels -> cell range of ring 0, tile centre
  {
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
  }
  __syncthreads();
  double bb[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double a = red[c * NW];
#pragma unroll
    for (int w2 = 1; w2 < NW; ++w2) a = fmin(a, red[c * NW + w2]);
    bb[c] = a;
  }
  int c0[3], c1[3];
  c0[0] = cell_of(bb[0], g.ox, g.inv_cell, g.cnx);
  c0[1] = cell_of(bb[1], g.oy, g.inv_cell, g.cny);
  c0[2] = cell_of(bb[2], g.oz, g.inv_cell, g.cnz);
  c1[0] = cell_of(-bb[3], g.ox, g.inv_cell, g.cnx);
  c1[1] = cell_of(-bb[4], g.oy, g.inv_cell, g.cny);
  c1[2] = cell_of(-bb[5], g.oz, g.inv_cell, g.cnz);
  const double cx = 0.5 * (bb[0] - bb[3]), cy = 0.5 * (bb[1] - bb[4]), cz = 0.5 * (bb[2] - bb[5]);
  const float qfx = (float)(qx - cx), qfy = (float)(qy - cy), qfz = (float)(qz - cz);

  // ---- local density -> histogram scale.  N1 = particles in the first box (cell-start lookups only)
  int r = max(p.r0, 0);
  Box box = make_box(c0, c1, r, g);
  int n1 = 0;
  for (int attempt = 0;; ++attempt) {
    const int nrows = (box.y1 - box.y0 + 1) * (box.z1 - box.z0 + 1);
    int mine = 0;
    for (int s = t; s < nrows; s += T) {
      int start, cnt;
      resolve_slot(g, box, box, false, s, nrows, start, cnt);
      mine += cnt;
    }
    int total;
    (void)block_scan_excl<T>(mine, warp_tot, &total);
    n1 = total;
    if (n1 >= kMinEstimate || attempt >= 3) break;
    const Box nb = make_box(c0, c1, r + 1, g);
    if (nb == box) break;
    box = nb;
    r += 1;
  }
  if (n1 < kMinEstimate) {  // nothing to estimate a scale from (deep void / tiny cloud): exact kernel
    if (t == 0) p.fail_list[atomicAdd(p.fail_count, 1)] = tile;
    return;
  }
  const double vol = (double)(box.x1 - box.x0 + 1) * (box.y1 - box.y0 + 1) * (box.z1 - box.z0 + 1) * g.cell * g.cell *
                     g.cell;
  const double r_est2 = pow(0.238732414637843 * k * vol / n1, 2.0 / 3.0);  // (3k / (4 pi rho))^(2/3)
  const double tmax = 2.5 * r_est2;
  const double binw = tmax / kNB;
  const float inv_w = (float)(1.0 / binw);

  // ---- phase A: ring walk with float32 histogram
  int* hist = hist_all + t;
#pragma unroll
  for (int b = 0; b < kNB; ++b) hist[b * T] = 0;
  Box prev = box;
  bool have_prev = false;
  for (;;) {
    scan_shell<T, false>(g, box, prev, have_prev, sm, cx, cy, cz, [&](int m) {
      if (active) {
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
          const float4 c = stage32[j];
          const float dx = qfx - c.x, dy = qfy - c.y, dz = qfz - c.z;
          const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          const int b = min(kNB - 1, __float2int_rz(d2 * inv_w));
          hist[b * T] += 1;
        }
      }
    });
    // exact stop test, conservative in float32: >= k particles in bins entirely below the nearest
    // unscanned face
    double gap = INFINITY;
    if (box.x0 > 0) gap = fmin(gap, qx - (g.ox + box.x0 * g.cell));
    if (box.x1 < g.cnx - 1) gap = fmin(gap, (g.ox + (box.x1 + 1) * g.cell) - qx);
    if (box.y0 > 0) gap = fmin(gap, qy - (g.oy + box.y0 * g.cell));
    if (box.y1 < g.cny - 1) gap = fmin(gap, (g.oy + (box.y1 + 1) * g.cell) - qy);
    if (box.z0 > 0) gap = fmin(gap, qz - (g.oz + box.z0 * g.cell));
    if (box.z1 < g.cnz - 1) gap = fmin(gap, (g.oz + (box.z1 + 1) * g.cell) - qz);
    gap -= 1e-6 * g.cell;
    bool done = true;
    if (active) {
      done = false;
      if (gap > 0.0) {
        const double gl = gap * gap * (1.0 - 1e-3);
        const int nfull = gl >= tmax ? kNB - 1 : min(kNB - 1, (int)(gl / binw));  // bins [0, nfull) lie below gl
        int cum = 0;
        for (int b = 0; b < nfull; ++b) cum += hist[b * T];
        done = cum >= k;
      }
    }
    if (__syncthreads_and(done ? 1 : 0)) break;
    const Box nb = make_box(c0, c1, r + 1, g);
    if (nb == box) break;  // the whole cell grid has been scanned
    prev = box;
    have_prev = true;
    box = nb;
    r += 1;
  }

  // ---- thresholds from the crossing bin
  bool fail = false;
  double e_lo = 0.0, e_hi = 0.0;
  if (active) {
    int cum = 0, bstar = -1;
    for (int b = 0; b < kNB - 1; ++b) {
      const int h = hist[b * T];
      if (cum + h >= k) {
        bstar = b;
        if (h > kListCap) fail = true;
        break;
      }
      cum += h;
    }
    if (bstar < 0) fail = true;  // k-th neighbour beyond the histogram range
    e_lo = bstar * binw;
    e_hi = (bstar + 1) * binw;
  }
  if (__syncthreads_or(fail ? 1 : 0)) {
    if (t == 0) p.fail_list[atomicAdd(p.fail_count, 1)] = tile;
    return;
  }

  // ---- phase B: exact classification of the final box
  // float32 error of d2 relative to the tile centre: coordinates are below `half` in magnitude
  double half = 0.0;
  {
    const double lo[3] = {g.ox + box.x0 * g.cell, g.oy + box.y0 * g.cell, g.oz + box.z0 * g.cell};
    const double hi[3] = {g.ox + (box.x1 + 1) * g.cell, g.oy + (box.y1 + 1) * g.cell, g.oz + (box.z1 + 1) * g.cell};
    const double cc[3] = {cx, cy, cz};
    const double ext[3] = {-(bb[0] + bb[3]), -(bb[1] + bb[4]), -(bb[2] + bb[5])};  // tile extent per axis
#pragma unroll
    for (int a = 0; a < 3; ++a)
      half = fmax(half, fmax(fabs(lo[a] - cc[a]), fabs(hi[a] - cc[a])) + ext[a]);
  }
  const double ec = half * 2.4e-7;  // 2 ulp of the largest coordinate
  const float hi32 = (float)((e_hi + 16.0 * sqrt(e_hi) * ec + 64.0 * ec * ec) * (1.0 + 1e-5));
  double* lkey = lkey_all + t;
  int* lidx = lidx_all + t;
  int n_in = 0, n_l = 0;
  bool overflow = false;
  const double eps = 1e-10;
  const bool sib = p.method == PTV_METHOD_SIBSON;
  const bool p2 = p.power == 2.0;
  double wsum = 0.0, su = 0.0, sv = 0.0, sw = 0.0;  // idw accumulators
  double dsum = 0.0, ksum = 0.0;                    // sibson moments: sum d, sum d^2
  const Box whole = box;
  auto idw_weight = [&](double d2) { return 1.0 / ((p2 ? d2 : pow(sqrt(d2), p.power)) + eps); };
  scan_shell<T, true>(g, whole, whole, false, sm, cx, cy, cz, [&](int m) {
    if (active) {
#pragma unroll 2
      for (int j = 0; j < m; ++j) {
        const float4 c = stage32[j];
        const float dx = qfx - c.x, dy = qfy - c.y, dz = qfz - c.z;
        const float d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d2f <= hi32) {
          const double2 xy = *reinterpret_cast<const double2*>(&stage64[j].x);
          const double zz = stage64[j].z;
          const int pidx = stage64[j].idx;
          const double ex = qx - xy.x, ey = qy - xy.y, ez = qz - zz;
          const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
          if (d2 < e_lo) {
            ++n_in;
            if (sib) {
              dsum += sqrt(d2);
              ksum += d2;
            } else {
              const double wgt = idw_weight(d2);
              const Value4 val = g.vals[pidx];
              wsum += wgt;
              su += wgt * val.u;
              sv += wgt * val.v;
              sw += wgt * val.w;
            }
          } else if (d2 < e_hi) {
            if (n_l < kListCap) {
              lkey[n_l * T] = d2;
              lidx[n_l * T] = pidx;
              ++n_l;
            } else {
              overflow = true;
            }
          }
        }
      }
    }
  });
  const int need = k - n_in;
  const bool bad = active && (overflow || need < 0 || need > n_l);
  if (__syncthreads_or(bad ? 1 : 0)) {
    if (t == 0) p.fail_list[atomicAdd(p.fail_count, 1)] = tile;
    return;
  }

  // ---- the `need` smallest (d2, index) of the list complete the k nearest
  if (active) {
    for (int i = 0; i < need; ++i) {
      int best = i;
      double bk = lkey[i * T];
      int bi = lidx[i * T];
      for (int j = i + 1; j < n_l; ++j) {
        const double kj = lkey[j * T];
        const int ij = lidx[j * T];
        if (key_greater(bk, bi, kj, ij)) { best = j; bk = kj; bi = ij; }
      }
      if (best != i) {
        lkey[best * T] = lkey[i * T];
        lidx[best * T] = lidx[i * T];
        lkey[i * T] = bk;
        lidx[i * T] = bi;
      }
      if (sib) {
        dsum += sqrt(bk);
        ksum += bk;
      } else {
        const double wgt = idw_weight(bk);
        const Value4 val = g.vals[bi];
        wsum += wgt;
        su += wgt * val.u;
        sv += wgt * val.v;
        sw += wgt * val.w;
      }
    }
  }

  if (sib) {
    // interpolator.py:102-122: w = (1/(d+eps)) * exp(-d / (std(d) + eps)), normalised
    const double mean = dsum / k;
    const double var = fmax(ksum / k - mean * mean, 0.0);
    const double inv_s = 1.0 / (sqrt(var) + eps);
    auto sib_acc = [&](double d2, int pidx) {
      const double d = sqrt(d2);
      const double wgt = (1.0 / (d + eps)) * exp(-d * inv_s);
      const Value4 val = g.vals[pidx];
      wsum += wgt;
      su += wgt * val.u;
      sv += wgt * val.v;
      sw += wgt * val.w;
    };
    if (active)
      for (int i = 0; i < need; ++i) sib_acc(lkey[i * T], lidx[i * T]);
    const float lo32 = (float)((e_lo + 16.0 * sqrt(e_lo) * ec + 64.0 * ec * ec) * (1.0 + 1e-5));
    scan_shell<T, true>(g, whole, whole, false, sm, cx, cy, cz, [&](int m) {
      if (active) {
#pragma unroll 2
        for (int j = 0; j < m; ++j) {
          const float4 c = stage32[j];
          const float dx = qfx - c.x, dy = qfy - c.y, dz = qfz - c.z;
          const float d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          if (d2f <= lo32) {
            const double2 xy = *reinterpret_cast<const double2*>(&stage64[j].x);
            const double zz = stage64[j].z;
            const double ex = qx - xy.x, ey = qy - xy.y, ez = qz - zz;
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
            if (d2 < e_lo) sib_acc(d2, stage64[j].idx);
          }
        }
      }
    });
  }

  if (p.stats != nullptr && t == 0) atomicAdd(&p.stats[0], 1ULL);
  if (!valid) return;
  double ou = 0.0, ov = 0.0, ow = 0.0;
  if (active) {
    ou = su / wsum; ov = sv / wsum; ow = sw / wsum;
    // main.py:195-199 nan_to_num
    if (ou != ou) ou = 0.0;
    if (ov != ov) ov = 0.0;
    if (ow != ow) ow = 0.0;
  }
  store_out<OutT>(p.u, vox, ou);
  store_out<OutT>(p.v, vox, ov);
  store_out<OutT>(p.w, vox, ow);
}

static size_t stream_smem_bytes(int T) {
  const int NW = T / 32;
  size_t b = (size_t)kListCap * T * 12 + (size_t)kStageCap * (sizeof(ParticleRec) + sizeof(float4)) +
             (size_t)6 * NW * sizeof(double) + (size_t)(2 * T + 1 + NW) * sizeof(int);
  return (b + 15) & ~(size_t)15;
}

template <int T, int TX, int TY, int TZ, typename OutT>
static int launch_stream_t(KnnParams& p, cudaStream_t stream) {
  p.tiles_x = (p.nx + TX - 1) / TX;
  p.tiles_y = (p.ny + TY - 1) / TY;
  p.tiles_z = (p.nz + TZ - 1) / TZ;
  const int64_t ntiles = (int64_t)p.tiles_x * p.tiles_y * p.tiles_z;
  if (ntiles > 2147483647LL) { set_error("ptv_knn_interp: grid too large for one launch"); return PTV_ERR_INVALID; }
  const size_t smem = stream_smem_bytes(T);
  auto kern = knn_stream_kernel<T, TX, TY, TZ, OutT>;
  PTV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)ntiles, T, smem, stream>>>(p);
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

int launch_knn_stream(KnnParams& p, int T, bool f32, cudaStream_t stream) {
  switch (T) {
    case 128: return f32 ? launch_stream_t<128, 8, 4, 4, float>(p, stream) : launch_stream_t<128, 8, 4, 4, double>(p, stream);
    case 64: return f32 ? launch_stream_t<64, 4, 4, 4, float>(p, stream) : launch_stream_t<64, 4, 4, 4, double>(p, stream);
    default: return f32 ? launch_stream_t<32, 4, 4, 2, float>(p, stream) : launch_stream_t<32, 4, 4, 2, double>(p, stream);
  }
}

}  // namespace ptv
