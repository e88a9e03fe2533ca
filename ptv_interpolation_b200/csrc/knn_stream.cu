// Streaming (heap-free) fused kNN + IDW / sibson kernel -- the production path for k >= 8.
//
// One CTA owns an 8x8x16 region of voxels, compacts its ACTIVE (pore) voxels and processes them in
// rounds of up to 128 lanes (one thread = one voxel).  Per round the cell list around the round's
// bounding box is scanned through shared memory (knn_common.cuh: rounded regions, cp.async double
// buffering), but the per-voxel k-best list is never materialised:
//
//   phase A  grow the scanned region through three radii; every thread histograms the float32 squared
//            distances of the staged particles into 96 16-bit bins (bin width from the tile's local
//            particle density).  Stop when every voxel has >= k particles in bins that lie wholly inside
//            the scanned radius.  The bin where the cumulative count crosses k gives two float64
//            thresholds E_lo < E_hi per voxel.
//   phase B  rescan the final region.  A float32 pre-test builds per-lane accept masks; accepted
//            particles get the exact float64 key d2 = (dx*dx + dy*dy) + dz*dz.  Keys below E_lo are
//            certainly among the k nearest and are accumulated on the fly; keys in [E_lo, E_hi) go to a
//            short list (<= 16 entries) from which the k - n_in smallest by (d2, index) are taken.
//   phase C  (sibson only) one more rescan to apply weights that need the std of the k distances.
//
// The selected SET is exactly the canonical k nearest whenever n_in <= k <= n_in + n_list and every key
// below E_hi was scanned (E_hi <= scanned radius^2 by construction).  Any region where that cannot be
// established (no local density estimate, k-th neighbour beyond the histogram range, crossing bin
// larger than the list, verification failure) is appended to a fail list and redone by the exact heap
// kernel (knn_interp.cu), so results never depend on the optimistic path succeeding.  Shared memory
// per voxel drops from 12*k bytes to 192 bytes (5 CTAs/SM instead of 2 at k = 50), the divergent heap
// maintenance disappears, and the main loops run in float32.
#include "knn_common.cuh"

namespace ptv {

static constexpr int kNB = 96;        // histogram bins over [0, Tmax), 16-bit counters
static constexpr int kListCap = 16;   // capacity of the crossing-bin list
static constexpr int kMinEstimate = 16;
static constexpr int kVPT = 8;  // voxels of the region examined per thread (region = kVPT * T voxels)
static_assert(kPipeCap <= 128, "the exact passes use two 64-bit accept masks per chunk");

// Values staged next to the candidates: float32 when the output is float32 (rounding 6e-8 relative,
// far inside the 1e-5 bar), float64 when the caller asked for float64 output.
template <typename OutT> struct StageVal;
template <> struct StageVal<float> {
  using type = float4;
  static constexpr int kind = 1;
  __device__ static float u(const float4& v) { return v.x; }
  __device__ static float v(const float4& v) { return v.y; }
  __device__ static float w(const float4& v) { return v.z; }
};
template <> struct StageVal<double> {
  using type = Value4;
  static constexpr int kind = 2;
  __device__ static double u(const Value4& v) { return v.u; }
  __device__ static double v(const Value4& v) { return v.v; }
  __device__ static double w(const Value4& v) { return v.w; }
};

template <int T, int TX, int TY, int TZ, typename OutT>
__global__ void __launch_bounds__(T, T == 128 ? 6 : 8) knn_stream_kernel(const KnnParams p) {
  static_assert(TX * TY * TZ == T, "tile shape");
  constexpr int NW = T / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // per-thread columns: histogram (phase A) aliases the crossing-bin list (phase B)
  double* lkey_all = reinterpret_cast<double*>(smem_raw);                       // [kListCap][T]
  int* lidx_all = reinterpret_cast<int*>(lkey_all + (size_t)kListCap * T);      // [kListCap][T]
  // 16-bit counters [kNB][T] (alias): two lanes share a 32-bit word, which is conflict-free
  uint16_t* hist_all = reinterpret_cast<uint16_t*>(smem_raw);
  static_assert(kNB * 2 <= kListCap * 12 && kNB % 2 == 0, "histogram must fit under the list");
  using ValT = typename StageVal<OutT>::type;  // float4 for float32 output, Value4 for float64 output
  constexpr int kVal = StageVal<OutT>::kind;
  ParticleRec* s64 = reinterpret_cast<ParticleRec*>(lidx_all + (size_t)kListCap * T);  // [2][kPipeCap]
  ValT* sval = reinterpret_cast<ValT*>(s64 + 2 * kPipeCap);                             // [2][kPipeCap]
  float4* s32 = reinterpret_cast<float4*>(sval + 2 * kPipeCap);                         // [2][kPipeCap]
  double* red = reinterpret_cast<double*>(s32 + 2 * kPipeCap);                          // [6][NW]
  int* seg_start = reinterpret_cast<int*>(red + 6 * NW);
  int* seg_off = seg_start + T;
  int* warp_tot = seg_off + T + 1;
  const PipeBuf pbuf[2] = {{s64, s32, sval}, {s64 + kPipeCap, s32 + kPipeCap, sval + kPipeCap}};
  uint16_t* vlist = reinterpret_cast<uint16_t*>(warp_tot + NW + 1);  // [kVPT*T] compacted active voxels
  // stats mode only: work counters of this CTA (voxel-candidate pairs of phases A and B/C, exact keys,
  // crossing-bin list entries), flushed once per CTA
  unsigned long long* wcnt = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(vlist + kVPT * T) + 7) & ~(uintptr_t)7);  // [4]

  const int t = threadIdx.x;
  const int k = p.k;
  const HashGrid& g = p.g;
  const bool kStats = p.stats != nullptr;
  if (kStats && t < 4) wcnt[t] = 0ULL;
  auto flush_stats = [&](int nvox, int nrounds) {
    __syncthreads();
    if (t == 0) {
      for (int i = 0; i < 4; ++i) atomicAdd(&p.stats[8 + i], wcnt[i]);
      atomicAdd(&p.stats[12], (unsigned long long)nvox);
      atomicAdd(&p.stats[13], (unsigned long long)nrounds);
    }
  };
  // ---- region of kVPT*T voxels (4x4x4 blocks, block-major) -> compact list of its ACTIVE voxels, so
  //      that every lane of a round owns a pore voxel even where the tile straddles a grain surface
  constexpr int RX = 8, RY = T >= 64 ? 8 : 4, RZ = (T == 128 ? 8 : 4) * (kVPT / 4);
  constexpr int NBX = RX / 4, NBY = RY / 4;
  static_assert(RX * RY * RZ == kVPT * T && kVPT % 4 == 0, "region holds kVPT voxels per thread");
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
  auto push_fail = [&](int reason) {  // hand every heap tile of this region to the exact kernel
    if (t == 0) {
      const int htx = (p.nx + TX - 1) / TX, hty = (p.ny + TY - 1) / TY, htz = (p.nz + TZ - 1) / TZ;
      for (int cz2 = 0; cz2 < RZ / TZ; ++cz2)
        for (int cy2 = 0; cy2 < RY / TY; ++cy2)
          for (int cx2 = 0; cx2 < RX / TX; ++cx2) {
            const int hx = rx * (RX / TX) + cx2, hy = ry * (RY / TY) + cy2, hz = rz * (RZ / TZ) + cz2;
            if (hx < htx && hy < hty && hz < htz)
              p.fail_list[atomicAdd(p.fail_count, 1)] = (hz * hty + hy) * htx + hx;
          }
      if (p.stats != nullptr) atomicAdd(&p.stats[reason], 1ULL);
    }
  };
  int nact;
  {
    int mine = 0;
    unsigned flags = 0;
#pragma unroll
    for (int q = 0; q < kVPT; ++q) {
      int ix, iy, iz;
      decode(kVPT * t + q, ix, iy, iz);
      if (ix < p.nx && iy < p.ny && iz < p.nz) {
        const int64_t vox = ((int64_t)iz * p.ny + iy) * p.nx + ix;
        if (p.mask == nullptr || p.mask[vox] != 0) {
          flags |= 1u << q;
          ++mine;
        } else {  // solid voxel: main.py:202-207 writes zero
          store_out<OutT>(p.u, vox, 0.0);
          store_out<OutT>(p.v, vox, 0.0);
          store_out<OutT>(p.w, vox, 0.0);
        }
      }
    }
    const int off = block_scan_excl<T>(mine, warp_tot, &nact);
    int o = off;
#pragma unroll
    for (int q = 0; q < kVPT; ++q)
      if (flags & (1u << q)) vlist[o++] = (uint16_t)(kVPT * t + q);
  }
  if (nact == 0) return;
  __syncthreads();
  const int rounds = (nact + T - 1) / T;
  int round_start = 0;
  for (int round = 0; round < rounds; ++round) {
  // greedy rounds: full warps first -- a short last round leaves whole warps idle at the barriers, which
  // costs no issue slots, instead of spreading idle lanes over every warp of every round
  const int round_cnt = min(T, nact - round_start);
  const bool active = t < round_cnt;
  int ix = 0, iy = 0, iz = 0;
  if (active) decode(vlist[round_start + t], ix, iy, iz);
  round_start += round_cnt;
  const int64_t vox = active ? ((int64_t)iz * p.ny + iy) * p.nx + ix : 0;
  const double qx = active ? p.ax[ix] : 0.0;
  const double qy = active ? p.ay[iy] : 0.0;
  const double qz = active ? p.az[iz] : 0.0;

  TileGeom tg;
  tile_geometry<T>(g, active, qx, qy, qz, red, tg);
  const double cx = 0.5 * (tg.lo[0] + tg.hi[0]), cy = 0.5 * (tg.lo[1] + tg.hi[1]), cz = 0.5 * (tg.lo[2] + tg.hi[2]);
  const float qfx = (float)(qx - cx), qfy = (float)(qy - cy), qfz = (float)(qz - cz);

  // ---- local density -> radius schedule and histogram scale
  const double r_est = estimate_radius<T>(g, tg, p.r0, k, kMinEstimate, warp_tot);
  if (!(r_est > 0.0)) {  // nothing to estimate a scale from (deep void / tiny cloud): exact kernel
    push_fail(1);
    return;
  }
  // three scan radii whose squares sit just above histogram bin edges 36, 60 and 96 (= Tmax)
  const double binw = p.rscale * r_est * r_est / 36.0;
  const float inv_w = (float)(1.0 / binw);
  constexpr int kEdges[3] = {36, 60, kNB};

  // ---- phase A: grow the scanned region, float32 histogram of squared distances
  uint16_t* hist = hist_all + t;
#pragma unroll
  for (int b = 0; b < kNB; ++b) hist[b * T] = 0;
  RoundRegion prev = make_region(g, tg, 0.0);
  RoundRegion rg = prev;
  bool have_prev = false, finished = false;
  for (int stage = 0; stage < 3; ++stage) {
    double R = sqrt((kEdges[stage] + 0.02) * binw) + 1e-6 * g.cell;
    const bool last = R >= tg.rmax;
    if (last) R = tg.rmax;
    rg = make_region(g, tg, R);
    scan_shell_pipe<T, 0>(g, tg, rg, prev, have_prev, pbuf, seg_start, seg_off, warp_tot, cx, cy, cz,
                          [&](const PipeBuf& pb, int m) {
      if (kStats && t == 0) wcnt[0] += (unsigned long long)m * round_cnt;
      if (active) {
        const float4* stage32 = pb.stage32;
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
    // stop test: >= k particles in bins that lie entirely inside the scanned radius (bins below
    // kEdges[stage]; float32 binning errors are far smaller than the 0.02-bin margin)
    bool done = true;
    if (active && !last) {
      int cum = 0;
      const int nfull = min(kEdges[stage], kNB - 1);
      for (int b = 0; b < nfull; ++b) cum += hist[b * T];
      done = cum >= k;
    }
    if (__syncthreads_and(done ? 1 : 0) || last) {
      finished = true;
      if (p.stats != nullptr && t == 0 && stage > 0) atomicAdd(&p.stats[4 + stage], 1ULL);  // finished at stage 1 / 2
      break;
    }
    prev = rg;
    have_prev = true;
  }
  if (!finished) {  // the k-th neighbour is beyond the histogram range for some voxel
    push_fail(2);
    return;
  }

  // ---- thresholds from the crossing bin
  bool fail = false, crowded = false;
  double e_lo = 0.0, e_hi = 0.0;
  int c_below = 0;  // particles in the bins below the crossing bin
  if (active) {
    int cum = 0, bstar = -1;
    for (int b = 0; b < kNB - 1; ++b) {
      const int h = hist[b * T];
      if (cum + h >= k) {
        bstar = b;
        crowded = h > kListCap;
        break;
      }
      cum += h;
    }
    if (bstar < 0) fail = true;  // crossing in the open-ended last bin
    c_below = cum;
    e_lo = bstar * binw;
    e_hi = (bstar + 1) * binw;
    // the crossing bin must lie inside the scanned radius unless the whole grid was scanned
    if (rg.R < tg.rmax && e_hi > (rg.R - 1e-6 * g.cell) * (rg.R - 1e-6 * g.cell)) fail = true;
  }
  if (__syncthreads_or(fail ? 1 : 0)) {
    push_fail(3);
    return;
  }
  // ---- phase A2 (rare): a crossing bin that holds more particles than the short list -- typical for
  //      voxels far from every particle, whose neighbours all sit at nearly the same distance -- is
  //      re-histogrammed with kNB sub-bins; phase B still verifies the result exactly.
  if (__syncthreads_or(crowded ? 1 : 0)) {
#pragma unroll
    for (int b = 0; b < kNB; ++b) hist[b * T] = 0;
    const float lo32 = (float)e_lo, inv_w2 = (float)((double)kNB / binw);
    scan_shell_pipe<T, 0>(g, tg, rg, rg, false, pbuf, seg_start, seg_off, warp_tot, cx, cy, cz,
                          [&](const PipeBuf& pb, int m) {
      if (active && crowded) {
        const float4* stage32 = pb.stage32;
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
          const float4 c = stage32[j];
          const float dx = qfx - c.x, dy = qfy - c.y, dz = qfz - c.z;
          const float rel = (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) - lo32) * inv_w2;
          if (rel >= 0.0f && rel < (float)kNB) hist[__float2int_rz(rel) * T] += 1;
        }
      }
    });
    if (active && crowded) {
      const double w2 = binw / kNB;
      int cum = c_below, b2 = -1;
      for (int b = 0; b < kNB; ++b) {
        const int h = hist[b * T];
        if (cum + h >= k) {
          b2 = b;
          if (h > kListCap) fail = true;  // still crowded (ties / coincident particles): exact kernel
          break;
        }
        cum += h;
      }
      if (b2 < 0) {
        fail = true;
      } else {  // one sub-bin of slack on each side absorbs float32 fuzz at this resolution
        const double lo2 = e_lo + (double)max(b2 - 1, 0) * w2, hi2 = fmin(e_lo + (double)(b2 + 2) * w2, e_hi);
        e_lo = lo2;
        e_hi = hi2;
      }
    }
    if (__syncthreads_or(fail ? 1 : 0)) {
      push_fail(3);
      return;
    }
  }

  // ---- phase B: exact classification of the final region
  // float32 error of d2 relative to the tile centre: staged coordinates are below `half` in magnitude
  const double half = rg.R + fmax(tg.hi[0] - tg.lo[0], fmax(tg.hi[1] - tg.lo[1], tg.hi[2] - tg.lo[2])) + g.cell;
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
  // sibson moments about a shift close to the distances themselves (the crossing-bin edge): the
  // one-pass variance E[(d-s)^2] - E[d-s]^2 then cancels only a few digits even when std << mean
  double dsum = 0.0, ksum = 0.0;
  const double dshift = sqrt(e_lo);
  // float32 output: the weight only needs float32 accuracy (relative 6e-8, far inside the 1e-5 bar),
  // so the reciprocal runs on the SFU; float64 output keeps the IEEE division.
  auto idw_weight = [&](double d2) -> double {
    const double den = (p2 ? d2 : pow(sqrt(d2), p.power)) + eps;
    if (sizeof(OutT) == 4 && p2) {  // den in [1e-10, ~1e12]: float32-safe; MUFU.RCP, 1 ulp
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)den));
      return (double)r;
    }
    return 1.0 / den;
  };
  // float32 pre-test of one 64-candidate half chunk -> bit mask of the candidates that need the
  // exact float64 key (the staged chunk is padded with far-away sentinels)
  auto prefilter64 = [&](const float4* stage32, int base, float lim) {
    unsigned long long msk = 0ULL;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      const float4 c = stage32[base + j];
      const float dx = qfx - c.x, dy = qfy - c.y, dz = qfz - c.z;
      const float d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      msk |= (unsigned long long)(d2f <= lim ? 1 : 0) << j;
    }
    return msk;
  };
  auto exact_d2 = [&](const ParticleRec* stage64, int j) {
    const double2 xy = *reinterpret_cast<const double2*>(&stage64[j].x);
    const double zz = stage64[j].z;
    const double ex = qx - xy.x, ey = qy - xy.y, ez = qz - zz;
    return __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
  };
  scan_shell_pipe<T, kVal>(g, tg, rg, rg, false, pbuf, seg_start, seg_off, warp_tot, cx, cy, cz,
                           [&](const PipeBuf& pb, int m) {
    if (kStats && t == 0) wcnt[1] += (unsigned long long)m * round_cnt;
    if (active) {
      const float4* stage32 = pb.stage32;
      const ParticleRec* stage64 = pb.stage64;
      const ValT* stage_val = reinterpret_cast<const ValT*>(pb.stage_val);
      unsigned long long m0 = prefilter64(stage32, 0, hi32);
      unsigned long long m1 = m > 64 ? prefilter64(stage32, 64, hi32) : 0ULL;
      if (kStats) atomicAdd(&wcnt[2], (unsigned long long)(__popcll(m0) + __popcll(m1)));
      // each thread walks only ITS accepted candidates (dense per lane instead of "any lane")
      while ((m0 | m1) != 0ULL) {
        int j;
        if (m0 != 0ULL) { j = __ffsll((long long)m0) - 1; m0 &= m0 - 1ULL; }
        else { j = 64 + __ffsll((long long)m1) - 1; m1 &= m1 - 1ULL; }
        const double d2 = exact_d2(stage64, j);
        if (d2 < e_lo) {
          ++n_in;
          if (sib) {
            const double dd = sqrt(d2) - dshift;
            dsum += dd;
            ksum += dd * dd;
          } else {
            const double wgt = idw_weight(d2);
            const ValT val = stage_val[j];
            wsum += wgt;
            su += wgt * (double)StageVal<OutT>::u(val);
            sv += wgt * (double)StageVal<OutT>::v(val);
            sw += wgt * (double)StageVal<OutT>::w(val);
          }
        } else if (d2 < e_hi) {
          if (n_l < kListCap) {
            lkey[n_l * T] = d2;
            lidx[n_l * T] = stage64[j].idx;
            ++n_l;
          } else {
            overflow = true;
          }
        }
      }
    }
  });
  const int need = k - n_in;
  if (kStats && active) atomicAdd(&wcnt[3], (unsigned long long)n_l);
  const bool bad = active && (overflow || need < 0 || need > n_l);
  if (__syncthreads_or(bad ? 1 : 0)) {
    push_fail(4);
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
        const double dd = sqrt(bk) - dshift;
        dsum += dd;
        ksum += dd * dd;
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
    auto sib_acc = [&](double d2, double vu, double vv, double vw) {
      const double d = sqrt(d2);
      const double wgt = (1.0 / (d + eps)) * exp(-d * inv_s);
      wsum += wgt;
      su += wgt * vu;
      sv += wgt * vv;
      sw += wgt * vw;
    };
    if (active)
      for (int i = 0; i < need; ++i) {
        const Value4 val = g.vals[lidx[i * T]];
        sib_acc(lkey[i * T], val.u, val.v, val.w);
      }
    const float lo32 = (float)((e_lo + 16.0 * sqrt(e_lo) * ec + 64.0 * ec * ec) * (1.0 + 1e-5));
    scan_shell_pipe<T, kVal>(g, tg, rg, rg, false, pbuf, seg_start, seg_off, warp_tot, cx, cy, cz,
                             [&](const PipeBuf& pb, int m) {
      if (kStats && t == 0) wcnt[1] += (unsigned long long)m * round_cnt;
      if (active) {
        const float4* stage32 = pb.stage32;
        const ParticleRec* stage64 = pb.stage64;
        const ValT* stage_val = reinterpret_cast<const ValT*>(pb.stage_val);
        unsigned long long m0 = prefilter64(stage32, 0, lo32);
        unsigned long long m1 = m > 64 ? prefilter64(stage32, 64, lo32) : 0ULL;
        while ((m0 | m1) != 0ULL) {
          int j;
          if (m0 != 0ULL) { j = __ffsll((long long)m0) - 1; m0 &= m0 - 1ULL; }
          else { j = 64 + __ffsll((long long)m1) - 1; m1 &= m1 - 1ULL; }
          const double d2 = exact_d2(stage64, j);
          if (d2 < e_lo) {
            const ValT val = stage_val[j];
            sib_acc(d2, (double)StageVal<OutT>::u(val), (double)StageVal<OutT>::v(val),
                    (double)StageVal<OutT>::w(val));
          }
        }
      }
    });
  }

  if (p.stats != nullptr && t == 0) atomicAdd(&p.stats[0], 1ULL);
  if (active) {
    double ou = su / wsum, ov = sv / wsum, ow = sw / wsum;
    // main.py:195-199 nan_to_num
    if (ou != ou) ou = 0.0;
    if (ov != ov) ov = 0.0;
    if (ow != ow) ow = 0.0;
    store_out<OutT>(p.u, vox, ou);
    store_out<OutT>(p.v, vox, ov);
    store_out<OutT>(p.w, vox, ow);
  }
  __syncthreads();  // shared memory is reused by the next round
  }  // rounds
  if (kStats) flush_stats(nact, rounds);
}

static size_t stream_smem_bytes(int T, bool f32) {
  const int NW = T / 32;
  size_t b = (size_t)kListCap * T * 12 +
             (size_t)2 * kPipeCap * (sizeof(ParticleRec) + sizeof(float4) + (f32 ? sizeof(float4) : sizeof(Value4))) +
             (size_t)6 * NW * sizeof(double) + (size_t)(2 * T + 2 + NW) * sizeof(int) + (size_t)kVPT * T * sizeof(uint16_t) +
             4 * sizeof(unsigned long long) + 8;
  return (b + 15) & ~(size_t)15;
}

template <int T, int TX, int TY, int TZ, typename OutT>
static int launch_stream_t(KnnParams& p, cudaStream_t stream) {
  constexpr int RX = 8, RY = T >= 64 ? 8 : 4, RZ = (T == 128 ? 8 : 4) * (kVPT / 4);  // region = kVPT voxels per thread
  p.tiles_x = (p.nx + RX - 1) / RX;
  p.tiles_y = (p.ny + RY - 1) / RY;
  p.tiles_z = (p.nz + RZ - 1) / RZ;
  const int64_t ntiles = (int64_t)p.tiles_x * p.tiles_y * p.tiles_z;
  if (ntiles > 2147483647LL) { set_error("ptv_knn_interp: grid too large for one launch"); return PTV_ERR_INVALID; }
  const size_t smem = stream_smem_bytes(T, sizeof(OutT) == 4);
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
