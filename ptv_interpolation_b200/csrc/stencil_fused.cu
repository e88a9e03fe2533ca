// One-pass masked divergence + plane fluxes + mean|div| statistics (physics.py:6-53,160-165,174;
// plot_flux.py:6-16).  HBM-bound: every field value is read once from DRAM (u, v, w, mask = 13 B per
// voxel for float32) and the divergence written once (4 B); the +-1 row / plane neighbours come from
// registers (y), L1 (x) and L2 (z).
//
// Mapping: a CTA owns `rows` consecutive y-rows of one z-plane and sweeps them top to bottom; a
// thread owns 4 consecutive x (one 16-byte vector per field per row) and keeps the v / mask rows in
// a sliding register window.  Flux partial sums stay in registers / shared memory for the whole
// sweep, so the global float64 atomics are one per x-column, per row and per CTA.
#include <math.h>

#include "ptv_internal.cuh"
#include "bulk_pipe.cuh"

namespace ptv {

template <typename Tf> struct Vec4 { Tf v[4]; };

template <typename Tf, bool kVec>
__device__ __forceinline__ Vec4<Tf> load4(const Tf* __restrict__ row, int x, int nx) {
  Vec4<Tf> r;
  if (kVec) {
    if (sizeof(Tf) == 4) {
      const float4 t = *reinterpret_cast<const float4*>(row + x);
      r.v[0] = (Tf)t.x; r.v[1] = (Tf)t.y; r.v[2] = (Tf)t.z; r.v[3] = (Tf)t.w;
    } else {
      const double2 a = *reinterpret_cast<const double2*>(row + x);
      const double2 b = *reinterpret_cast<const double2*>(row + x + 2);
      r.v[0] = (Tf)a.x; r.v[1] = (Tf)a.y; r.v[2] = (Tf)b.x; r.v[3] = (Tf)b.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) r.v[j] = x + j < nx ? row[x + j] : (Tf)0;
  }
  return r;
}

template <bool kVec>
__device__ __forceinline__ uchar4 loadm4(const uint8_t* __restrict__ row, int x, int nx) {
  if (kVec) return *reinterpret_cast<const uchar4*>(row + x);
  uchar4 r;
  r.x = x + 0 < nx ? row[x + 0] : 0;
  r.y = x + 1 < nx ? row[x + 1] : 0;
  r.z = x + 2 < nx ? row[x + 2] : 0;
  r.w = x + 3 < nx ? row[x + 3] : 0;
  return r;
}

__device__ __forceinline__ uint8_t mget(const uchar4& m, int j) {
  return j == 0 ? m.x : j == 1 ? m.y : j == 2 ? m.z : m.w;
}

// Closed form of physics.py:26-53 on FACE SUMS.  With S(i+1/2) = m[i+1] ? f[i] + f[i+1] : 0
// (and S = 2 f at the two domain edges) the reference's F+ - F- equals 0.5 * (S(i+1/2) - S(i-1/2)), and because
// scaling by 0.5 commutes with IEEE rounding,
//     div = 0.5 * ((Sx+ - Sx-) / dx + (Sy+ - Sy-) / dy) + (Sz+ - Sz-) / dz)
// is bit-identical to the reference's float64 result while every face sum is shared by its two voxels.
__device__ __forceinline__ double face_sum(double a, double b, bool open) {
  return open ? __dadd_rn(a, b) : 0.0;
}

// t / h with the IEEE result, but without entering the division's slow path for the (very common)
// zero numerators of masked faces: 0 / h is a signed zero.
__device__ __forceinline__ double div_spacing(double t, double h) {
  if (t == 0.0) return h > 0.0 ? t : -t;
  return __ddiv_rn(t, h);
}

static constexpr int kSfThreads = 256;

template <typename Tf, bool kVec, bool kUnit>
__global__ void __launch_bounds__(kSfThreads, 2) div_flux_kernel(
    const Tf* __restrict__ u, const Tf* __restrict__ v, const Tf* __restrict__ w, const uint8_t* __restrict__ mask,
    int nx, int ny, int nz, double dx, double dy, double dz, const Tf* __restrict__ w_below,
    const Tf* __restrict__ w_above, const uint8_t* __restrict__ mask_above, Tf* __restrict__ div,
    double* __restrict__ stats, double* __restrict__ qxy, double* __restrict__ qxz, double* __restrict__ qyz,
    int rows, int chunks_y) {
  // [rows][warps] per-warp partial sums over x of v per row (flux through XZ planes); plain stores --
  // float64 atomics on shared memory compile to a CAS loop
  extern __shared__ double row_part[];
  constexpr int kWarps = kSfThreads / 32;
  __shared__ double red[3][kSfThreads / 32];
  const int z = blockIdx.x / chunks_y;
  const int y0 = (blockIdx.x % chunks_y) * rows;
  const int y1 = min(ny, y0 + rows);
  const int t = threadIdx.x;
  const int64_t plane = (int64_t)nx * ny;
  // z neighbours: inside the slab, from the halo planes, or absent (domain edge -> Neumann)
  const bool z_lo_edge = z == 0 && w_below == nullptr;
  const bool z_hi_edge = z == nz - 1 && w_above == nullptr;
  const bool want_flux = qxy != nullptr;

  for (int r = t; r < rows * kWarps; r += kSfThreads) row_part[r] = 0.0;
  __syncthreads();
  double acc_w = 0.0, acc_abs = 0.0;
  int acc_cnt = 0;

  for (int xc = 0; xc < nx; xc += 4 * kSfThreads) {
    const int x = xc + 4 * t;
    const bool in = x < nx;
    const bool x_first = x == 0, x_last = x + 4 >= nx;
    double col_u[4] = {0.0, 0.0, 0.0, 0.0};
    double vc[4], yface[4];   // current v row (float64) and the face sums S(y-1/2)
    uchar4 m_cur = make_uchar4(0, 0, 0, 0);
    // row pointers of this thread's 4 columns, advanced by nx per row
    const int64_t o0 = (int64_t)z * plane + (int64_t)y0 * nx;
    const Tf* urow = u + o0;
    const Tf* vrow_p = v + o0;
    const Tf* wrow = w + o0;
    const uint8_t* mrow = mask + o0;
    Tf* drow = div + o0;
    const Tf* wbrow = z > 0 ? wrow - plane : (w_below != nullptr ? w_below + (int64_t)y0 * nx : wrow);
    const Tf* warow = z < nz - 1 ? wrow + plane : (w_above != nullptr ? w_above + (int64_t)y0 * nx : wrow);
    const uint8_t* marow = z < nz - 1 ? mrow + plane : (mask_above != nullptr ? mask_above + (int64_t)y0 * nx : mrow);
    if (in) {
      const Vec4<Tf> v0 = load4<Tf, kVec>(vrow_p, x, nx);
      m_cur = loadm4<kVec>(mrow, x, nx);
      const uint8_t mc[4] = {m_cur.x, m_cur.y, m_cur.z, m_cur.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) vc[j] = (double)v0.v[j];
      if (y0 > 0) {
        const Vec4<Tf> vm = load4<Tf, kVec>(vrow_p - nx, x, nx);
#pragma unroll
        for (int j = 0; j < 4; ++j) yface[j] = face_sum((double)vm.v[j], vc[j], mc[j] != 0);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) yface[j] = __dadd_rn(vc[j], vc[j]);  // Neumann edge: F- = v
      }
    }
    for (int y = y0; y < y1; ++y) {
      double vsum = 0.0;
      if (in) {
        const bool y_hi = y == ny - 1;
        // ---- loads of this row (v and mask of the NEXT row feed the y+1/2 faces)
        const Vec4<Tf> u4 = load4<Tf, kVec>(urow, x, nx);
        const Vec4<Tf> w4 = load4<Tf, kVec>(wrow, x, nx);
        Vec4<Tf> vn4, wb4, wa4;
        uchar4 m_next = m_cur, ma4 = m_cur;
        if (!y_hi) {
          vn4 = load4<Tf, kVec>(vrow_p + nx, x, nx);
          m_next = loadm4<kVec>(mrow + nx, x, nx);
        }
        if (!z_lo_edge) wb4 = load4<Tf, kVec>(wbrow, x, nx);
        if (!z_hi_edge) {
          wa4 = load4<Tf, kVec>(warow, x, nx);
          ma4 = loadm4<kVec>(marow, x, nx);
        }
        const Tf u_l = x_first ? (Tf)0 : urow[x - 1];
        const Tf u_r = x_last ? (Tf)0 : urow[x + 4];
        const uint8_t m_r = x_last ? (uint8_t)0 : mrow[x + 4];
        const uint8_t mc[4] = {m_cur.x, m_cur.y, m_cur.z, m_cur.w};
        const uint8_t mn[4] = {m_next.x, m_next.y, m_next.z, m_next.w};
        const uint8_t mab[4] = {ma4.x, ma4.y, ma4.z, ma4.w};
        // ---- x face sums S(j-1/2), j = 0..4 (face j lies left of voxel j)
        double ud[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ud[j] = (double)u4.v[j];
        double xf[5];
        xf[0] = x_first ? __dadd_rn(ud[0], ud[0]) : face_sum((double)u_l, ud[0], mc[0] != 0);
#pragma unroll
        for (int j = 1; j < 4; ++j) xf[j] = face_sum(ud[j - 1], ud[j], mc[j] != 0);
        if (kVec) {
          xf[4] = x_last ? __dadd_rn(ud[3], ud[3]) : face_sum(ud[3], (double)u_r, m_r != 0);
        } else {  // ragged tail: the domain edge may fall inside this thread's 4 columns
#pragma unroll
          for (int j = 1; j < 4; ++j)
            if (x + j == nx) xf[j] = __dadd_rn(ud[j - 1], ud[j - 1]);
          xf[4] = x + 4 == nx ? __dadd_rn(ud[3], ud[3]) : (x + 4 < nx ? face_sum(ud[3], (double)u_r, m_r != 0) : 0.0);
        }
        Vec4<Tf> d4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double wd = (double)w4.v[j];
          // y faces: S(y+1/2) computed here, S(y-1/2) carried from the previous row
          const double yn = y_hi ? __dadd_rn(vc[j], vc[j]) : face_sum(vc[j], (double)vn4.v[j], mn[j] != 0);
          // z faces
          const double zl = z_lo_edge ? __dadd_rn(wd, wd) : face_sum((double)wb4.v[j], wd, mc[j] != 0);
          const double zh = z_hi_edge ? __dadd_rn(wd, wd) : face_sum(wd, (double)wa4.v[j], mab[j] != 0);
          double ax = __dsub_rn(xf[j + 1], xf[j]);
          double ay = __dsub_rn(yn, yface[j]);
          double az = __dsub_rn(zh, zl);
          if (!kUnit) {  // kUnit: dx == dy == dz == 1 and x / 1.0 == x exactly
            ax = div_spacing(ax, dx);
            ay = div_spacing(ay, dy);
            az = div_spacing(az, dz);
          }
          const double d = __dmul_rn(__dadd_rn(__dadd_rn(ax, ay), az), 0.5);
          d4.v[j] = (Tf)d;
          if (kVec || x + j < nx) {
            if (mc[j] != 0) {
              acc_abs += fabs((double)d4.v[j]);
              acc_cnt += 1;
            }
            col_u[j] += ud[j];
            vsum += vc[j];
            acc_w += wd;
          }
          yface[j] = yn;
          if (!y_hi) vc[j] = (double)vn4.v[j];
        }
        if (kVec) {
          if (sizeof(Tf) == 4) {
            *reinterpret_cast<float4*>(drow + x) = make_float4((float)d4.v[0], (float)d4.v[1], (float)d4.v[2], (float)d4.v[3]);
          } else {
            *reinterpret_cast<double2*>(drow + x) = make_double2((double)d4.v[0], (double)d4.v[1]);
            *reinterpret_cast<double2*>(drow + x + 2) = make_double2((double)d4.v[2], (double)d4.v[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (x + j < nx) drow[x + j] = d4.v[j];
        }
        m_cur = m_next;
        urow += nx; vrow_p += nx; wrow += nx; mrow += nx; drow += nx; wbrow += nx; warow += nx; marow += nx;
      }
      if (want_flux) {  // row sum of v over this x-chunk: warp shuffle, per-warp slot in shared memory
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        if ((t & 31) == 0) row_part[(y - y0) * kWarps + (t >> 5)] += vsum;
      }
    }
    if (want_flux && in) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (x + j < nx) atomicAdd(&qyz[x + j], col_u[j]);
    }
  }

  // block totals: sum of w (flux through XY planes), sum |div| and fluid count
  double cnt = (double)acc_cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_w += __shfl_xor_sync(0xffffffffu, acc_w, o);
    acc_abs += __shfl_xor_sync(0xffffffffu, acc_abs, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((t & 31) == 0) {
    red[0][t >> 5] = acc_w;
    red[1][t >> 5] = acc_abs;
    red[2][t >> 5] = cnt;
  }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = 0; i < kSfThreads / 32; ++i) { a += red[0][i]; b += red[1][i]; c += red[2][i]; }
    if (want_flux) atomicAdd(&qxy[z], a);
    if (stats != nullptr && c > 0.0) {
      atomicAdd(&stats[0], b);
      atomicAdd(&stats[1], c);
    }
  }
  if (want_flux)
    for (int r = t; r < y1 - y0; r += kSfThreads) {
      double a = 0.0;
#pragma unroll
      for (int i = 0; i < kWarps; ++i) a += row_part[r * kWarps + i];
      atomicAdd(&qxz[y0 + r], a);
    }
}

// ------------------------------------------------------------------------------------------------------
// Bulk-async row pipeline (sm_100a): the same sweep, but the rows a CTA needs travel through a ring of shared
// memory stages filled by 1-D bulk copies (cp.async.bulk, the TMA engine) that one producer thread issues
// rows ahead of the eight consumer warps.  The per-row load latency that bounded the kernel above (3.7
// long-scoreboard stall cycles per issued instruction at 16 warps/SM) disappears from the consumers: they wait
// on a stage's mbarrier (already complete in steady state), read it with conflict-free LDS.128 and hand it back.
// Stage = {u[y] (+4-column x halos), w[y], v[y+1], w[z-1][y], w[z+1][y], mask[y+1] (+16 B), mask[z+1][y]}.
// Needs nx % 16 == 0 and 16-byte aligned fields (bulk copies move multiples of 16 bytes).

static constexpr int kBfConsumers = 256;              // 8 consumer warps, 4 columns per thread
static constexpr int kBfThreads = kBfConsumers;       // thread 0 doubles as the producer (a ninth warp would cap the kernel at 96 registers)
static constexpr int kBfCW = 4 * kBfConsumers;        // columns per x-chunk
static constexpr int kBfMaxRows = 64;                 // rows per CTA sweep (bounds the flux row slots)

template <typename Tf> struct BfLayout {
  static constexpr int kFRow = (kBfCW + 8) * (int)sizeof(Tf);  // field row with a 4-column halo on both sides
  static constexpr int kMRow = kBfCW + 16;                     // mask row with a 16-byte halo on the right
  static constexpr int kStage = 5 * kFRow + 2 * kMRow;
  static constexpr int kStages = sizeof(Tf) == 4 ? 4 : 2;
  static constexpr int kRing = kStages * kStage;
};

template <typename Tf>
__device__ __forceinline__ Vec4<Tf> lds4(const unsigned char* row, int col) {
  Vec4<Tf> r;
  const Tf* p = reinterpret_cast<const Tf*>(row) + col;
  if (sizeof(Tf) == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = (Tf)t.x; r.v[1] = (Tf)t.y; r.v[2] = (Tf)t.z; r.v[3] = (Tf)t.w;
  } else {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *(reinterpret_cast<const double2*>(p) + 1);
    r.v[0] = (Tf)a.x; r.v[1] = (Tf)a.y; r.v[2] = (Tf)b.x; r.v[3] = (Tf)b.y;
  }
  return r;
}

struct Recip3 { double x, y, z; };

// face sum S = a + b across an open face, 0 across a closed one (`open` = the mask byte, still in place in its word)
__device__ __forceinline__ double face_if(double a, double b, uint32_t open) {
  double s;  // zero, then a predicated add: one instruction less than add + two 32-bit selects
  asm("{\n.reg .pred p;\nsetp.ne.u32 p, %3, 0;\nmov.f64 %0, 0d0000000000000000;\n@p add.rn.f64 %0, %1, %2;\n}"
      : "=d"(s) : "d"(a), "d"(b), "r"(open));
  return s;
}

// |d| as float64 for fluid voxels, 0 for solid ones; float32: the select is one AND on the bits before the conversion
__device__ __forceinline__ double abs_if(float d, uint32_t open) {
  const uint32_t keep = open != 0u ? 0x7fffffffu : 0u;
  return (double)__uint_as_float(__float_as_uint(d) & keep);
}
__device__ __forceinline__ double abs_if(double d, uint32_t open) { return open != 0u ? fabs(d) : 0.0; }

// Sum over the warp's 128 columns of up to 8 rows at once: every row the lanes park their 4-column sums in
// vs[row & 7][lane]; here lane = (row, quarter) adds 8 of them (stride 36 doubles: conflict-free both ways) and
// two shuffle steps finish -- 4 instructions per row instead of a 5-step float64 butterfly per row.
static constexpr int kBfVsStride = 36;
__device__ __forceinline__ void flush_row_sums(const double* vs, double* row_part, int row0, int nb, int warp, int lane) {
  __syncwarp();
  const int row = lane >> 2, q = lane & 3;
  double a = 0.0;
  if (row < nb) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a += vs[row * kBfVsStride + i * 4 + q];
  }
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  if (q == 0 && row < nb) row_part[(row0 + row) * (kBfConsumers / 32) + warp] += a;
  __syncwarp();
}

template <typename Tf, bool kUnit>
__global__ void __launch_bounds__(kBfThreads, 2) div_flux_bulk_kernel(
    const Tf* __restrict__ u, const Tf* __restrict__ v, const Tf* __restrict__ w, const uint8_t* __restrict__ mask,
    int nx, int ny, int nz, double dx, double dy, double dz, const Recip3 rh, const Tf* __restrict__ w_below,
    const Tf* __restrict__ w_above, const uint8_t* __restrict__ mask_above, Tf* __restrict__ div,
    double* __restrict__ stats, double* __restrict__ qxy, double* __restrict__ qxz, double* __restrict__ qyz,
    int rows, int chunks_y, int la) {
  using L = BfLayout<Tf>;
  constexpr int kWarps = kBfConsumers / 32;
  extern __shared__ __align__(128) unsigned char bf_smem[];
  unsigned char* ring = bf_smem;
  double* vs_all = reinterpret_cast<double*>(bf_smem + L::kRing);  // [8 warps][8 rows][36]
  double* row_part = vs_all + kWarps * 8 * kBfVsStride;           // [rows][8 warps]
  __shared__ __align__(8) uint64_t full_bar[L::kStages], empty_bar[L::kStages];
  __shared__ double red[3][kWarps];
  const int z = blockIdx.x / chunks_y;
  const int y0 = (blockIdx.x % chunks_y) * rows;
  const int y1 = min(ny, y0 + rows);
  const int nrow = y1 - y0;
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int64_t plane = (int64_t)nx * ny;
  const bool z_lo_edge = z == 0 && w_below == nullptr;
  const bool z_hi_edge = z == nz - 1 && w_above == nullptr;
  const bool want_flux = qxy != nullptr;
  const int nchunks = (nx + kBfCW - 1) / kBfCW;

  if (t == 0) {
#pragma unroll
    for (int s = 0; s < L::kStages; ++s) {
      mbar_init(smem_addr(&full_bar[s]), 1);
      mbar_init(smem_addr(&empty_bar[s]), kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int r = t; r < rows * kWarps; r += kBfThreads) row_part[r] = 0.0;
  __syncthreads();

  double acc_w = 0.0, acc_abs = 0.0;
  int acc_cnt = 0;
  const int64_t o0 = (int64_t)z * plane + (int64_t)y0 * nx;  // first row of the sweep

  // ---------------- producer duty: row `itp` is requested by lane 0 of warp itp % 8, `la` rows ahead of
  // its use (rotating the duty keeps the warps level: a warp that always produced would trail the others, and
  // they would burn its issue slots polling the barrier it has not armed yet)
  const Tf* wb0 = z > 0 ? w + o0 - plane : (w_below != nullptr ? w_below + (int64_t)y0 * nx : nullptr);
  const Tf* wa0 = z < nz - 1 ? w + o0 + plane : (w_above != nullptr ? w_above + (int64_t)y0 * nx : nullptr);
  const uint8_t* ma0 = z < nz - 1 ? mask + o0 + plane : (mask_above != nullptr ? mask_above + (int64_t)y0 * nx : nullptr);
  const int total_it = nchunks * nrow;
  auto produce = [&](int itp) {
    const int ci = itp / nrow, r = itp - ci * nrow;
    const int xc = ci * kBfCW;
    const int cw = min(kBfCW, nx - xc);
    const int left = xc > 0 ? 4 : 0, right = xc + cw < nx ? 4 : 0;
    const int s = itp % L::kStages;
    if (itp >= L::kStages) mbar_wait(smem_addr(&empty_bar[s]), ((itp / L::kStages) & 1) ^ 1);
    const bool y_hi = y0 + r == ny - 1;
    const int64_t o = (int64_t)r * nx + xc;
    const uint32_t fb = (uint32_t)cw * sizeof(Tf);
    const uint32_t ub = (uint32_t)(cw + left + right) * sizeof(Tf);
    const uint32_t mnb = (uint32_t)cw + (right ? 16u : 0u);
    // a Neumann edge is the open face between a value and itself (a + a): at the domain edges the neighbour slot
    // is filled with the row itself, so the consumers have no edge cases in their data path
    const uint32_t total = ub + 4u * fb + (y_hi ? 0u : mnb) + (z_hi_edge ? 0u : (uint32_t)cw);
    const uint32_t bar = smem_addr(&full_bar[s]);
    const uint32_t sb = smem_addr(ring + (size_t)s * L::kStage);
    mbar_expect_tx(bar, total);
    bulk_g2s(sb + (uint32_t)(4 - left) * sizeof(Tf), u + o0 + o - left, ub, bar);
    bulk_g2s(sb + L::kFRow, w + o0 + o, fb, bar);
    bulk_g2s(sb + 2 * L::kFRow, v + o0 + o + (y_hi ? 0 : nx), fb, bar);
    if (!y_hi) bulk_g2s(sb + 5 * L::kFRow, mask + o0 + o + nx, mnb, bar);
    bulk_g2s(sb + 3 * L::kFRow, z_lo_edge ? w + o0 + o : wb0 + o, fb, bar);
    bulk_g2s(sb + 4 * L::kFRow, z_hi_edge ? w + o0 + o : wa0 + o, fb, bar);
    if (!z_hi_edge) bulk_g2s(sb + 5 * L::kFRow + L::kMRow, ma0 + o, (uint32_t)cw, bar);
  };
  if (t == 0)
    for (int itp = 0; itp < la && itp < total_it; ++itp) produce(itp);

  // ---------------- consumers
  double* vs = vs_all + warp * 8 * kBfVsStride;
  int it = 0;
  for (int ci = 0; ci < nchunks; ++ci) {
    const int xc = ci * kBfCW;
    const int x = xc + 4 * t;
    const bool in = x < nx;
    const bool x_first = x == 0, x_last = x + 4 >= nx;
    double col_u[4] = {0.0, 0.0, 0.0, 0.0};
    double vc[4] = {0.0, 0.0, 0.0, 0.0}, yface[4] = {0.0, 0.0, 0.0, 0.0};
    uint32_t m_cur = 0u, m_r = 0u;  // mask bytes of the thread's 4 voxels (one word) and of the voxel right of them
    Tf* drow = div + o0;
    if (in) {  // first row of the sweep: v, mask (and v of the row above) straight from global memory
      const Vec4<Tf> v0 = load4<Tf, true>(v + o0, x, nx);
      m_cur = *reinterpret_cast<const uint32_t*>(mask + o0 + x);
      if (!x_last) m_r = mask[o0 + x + 4];
#pragma unroll
      for (int j = 0; j < 4; ++j) vc[j] = (double)v0.v[j];
      if (y0 > 0) {
        const Vec4<Tf> vm = load4<Tf, true>(v + o0 - nx, x, nx);
#pragma unroll
        for (int j = 0; j < 4; ++j) yface[j] = face_if((double)vm.v[j], vc[j], m_cur & (0xffu << (8 * j)));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) yface[j] = __dadd_rn(vc[j], vc[j]);  // Neumann edge: F- = v
      }
    }
#pragma unroll 1  // two rows per trip were tried: 40 % slower (the loop no longer fits the instruction cache)
    for (int y = y0; y < y1; ++y, ++it) {
      {
        const int itp = it + la;
        if (lane == 0 && (itp & (kWarps - 1)) == warp && itp < total_it) produce(itp);
      }
      const int s = it % L::kStages;
      const unsigned char* sb = ring + (size_t)s * L::kStage;
      mbar_wait(smem_addr(&full_bar[s]), (it / L::kStages) & 1);
      Vec4<Tf> u4, w4, vn4, wb4, wa4;
      uint32_t m_next = 0u, m_above = 0u, m_r_next = 0u;
      Tf u_l = (Tf)0, u_r = (Tf)0;
      const bool y_hi = y == ny - 1;
      if (in) {
        const int c = 4 * t;
        u4 = lds4<Tf>(sb, 4 + c);
        u_l = reinterpret_cast<const Tf*>(sb)[3 + c];
        u_r = reinterpret_cast<const Tf*>(sb)[8 + c];
        w4 = lds4<Tf>(sb + L::kFRow, c);
        vn4 = lds4<Tf>(sb + 2 * L::kFRow, c);
        wb4 = lds4<Tf>(sb + 3 * L::kFRow, c);
        wa4 = lds4<Tf>(sb + 4 * L::kFRow, c);
        const unsigned char* smn = sb + 5 * L::kFRow;
        m_next = *reinterpret_cast<const uint32_t*>(smn + c);
        m_r_next = smn[c + 4];
        m_above = *reinterpret_cast<const uint32_t*>(smn + L::kMRow + c);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&empty_bar[s]));  // the stage's values are in registers
      double vsum = 0.0;
      if (in) {
        // edges as data (see the producer): only the open/closed words depend on the edge flags
        const uint32_t oy = y_hi ? 0xffffffffu : m_next;
        const uint32_t ozl = z_lo_edge ? 0xffffffffu : m_cur;
        const uint32_t ozh = z_hi_edge ? 0xffffffffu : m_above;
        double ud[4], wd[4], vnd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ud[j] = (double)u4.v[j];
          wd[j] = (double)w4.v[j];
          vnd[j] = (double)vn4.v[j];
        }
        double xf[5];
        xf[0] = face_if((double)(x_first ? u4.v[0] : u_l), ud[0], x_first ? 1u : (m_cur & 0xffu));
#pragma unroll
        for (int j = 1; j < 4; ++j) xf[j] = face_if(ud[j - 1], ud[j], m_cur & (0xffu << (8 * j)));
        xf[4] = face_if(ud[3], (double)(x_last ? u4.v[3] : u_r), x_last ? 1u : m_r);
        Vec4<Tf> d4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t bj = 0xffu << (8 * j);
          const double yn = face_if(vc[j], vnd[j], oy & bj);
          const double zl = face_if((double)wb4.v[j], wd[j], ozl & bj);
          const double zh = face_if(wd[j], (double)wa4.v[j], ozh & bj);
          double ax = __dsub_rn(xf[j + 1], xf[j]);
          double ay = __dsub_rn(yn, yface[j]);
          double az = __dsub_rn(zh, zl);
          if (!kUnit) {  // kUnit: dx == dy == dz == 1 and x / 1.0 == x exactly
            ax = div_by_spacing(ax, dx, rh.x);
            ay = div_by_spacing(ay, dy, rh.y);
            az = div_by_spacing(az, dz, rh.z);
          }
          const double d = __dmul_rn(__dadd_rn(__dadd_rn(ax, ay), az), 0.5);
          d4.v[j] = (Tf)d;
          acc_abs += abs_if(d4.v[j], m_cur & bj);
          col_u[j] += ud[j];
          acc_w += wd[j];
          yface[j] = yn;
        }
        vsum = __dadd_rn(__dadd_rn(vc[0], vc[1]), __dadd_rn(vc[2], vc[3]));
#pragma unroll
        for (int j = 0; j < 4; ++j) vc[j] = vnd[j];
        {  // fluid voxels of the four: nonzero bytes of the mask word
          const uint32_t nzb = (((m_cur & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m_cur) & 0x80808080u;
          acc_cnt += __popc(nzb);
        }
        if (sizeof(Tf) == 4) {
          *reinterpret_cast<float4*>(drow + x) = make_float4((float)d4.v[0], (float)d4.v[1], (float)d4.v[2], (float)d4.v[3]);
        } else {
          *reinterpret_cast<double2*>(drow + x) = make_double2((double)d4.v[0], (double)d4.v[1]);
          *reinterpret_cast<double2*>(drow + x + 2) = make_double2((double)d4.v[2], (double)d4.v[3]);
        }
        m_cur = m_next;
        m_r = m_r_next;
        drow += nx;
      }
      if (want_flux) {
        const int r = y - y0;
        vs[(r & 7) * kBfVsStride + lane] = vsum;
        if ((r & 7) == 7 || y == y1 - 1) flush_row_sums(vs, row_part, r & ~7, (r & 7) + 1, warp, lane);
      }
    }
    if (want_flux && in) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&qyz[x + j], col_u[j]);
    }
  }

  double cnt = (double)acc_cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_w += __shfl_xor_sync(0xffffffffu, acc_w, o);
    acc_abs += __shfl_xor_sync(0xffffffffu, acc_abs, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    red[0][warp] = acc_w;
    red[1][warp] = acc_abs;
    red[2][warp] = cnt;
  }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = 0; i < kWarps; ++i) { a += red[0][i]; b += red[1][i]; c += red[2][i]; }
    if (want_flux) atomicAdd(&qxy[z], a);
    if (stats != nullptr && c > 0.0) {
      atomicAdd(&stats[0], b);
      atomicAdd(&stats[1], c);
    }
  }
  if (want_flux)
    for (int r = t; r < nrow; r += kBfThreads) {
      double a = 0.0;
#pragma unroll
      for (int i = 0; i < kWarps; ++i) a += row_part[r * kWarps + i];
      atomicAdd(&qxz[y0 + r], a);
    }
}

template <typename Tf>
static int launch_div_flux(const void* u, const void* v, const void* w, const uint8_t* mask, int nx, int ny, int nz,
                           double dx, double dy, double dz, const void* w_below, const void* w_above,
                           const uint8_t* mask_above, void* div, double* stats, double* qxy, double* qxz, double* qyz,
                           cudaStream_t s) {
  // rows per CTA: enough CTAs to fill 148 SMs several times over, few enough that the per-column
  // float64 atomics (nz * chunks_y * nx of them) stay negligible
  int chunks_y = 1;
  while ((int64_t)nz * chunks_y < 148 * 8 && chunks_y * 16 < ny) chunks_y *= 2;
  int rows = (ny + chunks_y - 1) / chunks_y;
  chunks_y = (ny + rows - 1) / rows;
  const auto al = [](const void* p, size_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
  const bool vec = (nx % 4 == 0) && al(u, 16) && al(v, 16) && al(w, 16) && al(div, 16) && al(w_below, 16) &&
                   al(w_above, 16) && al(mask, 4) && al(mask_above, 4);
  const bool unit = dx == 1.0 && dy == 1.0 && dz == 1.0;
  const bool bulk = tuning().stencil_bulk != 0 && (nx % 16 == 0) && al(u, 16) && al(v, 16) && al(w, 16) && al(div, 16) &&
                    al(w_below, 16) && al(w_above, 16) && al(mask, 16) && al(mask_above, 16) &&
                    (unit || (spacing_ok(dx) && spacing_ok(dy) && spacing_ok(dz)));
  if (bulk) {
    using L = BfLayout<Tf>;
    while (rows > kBfMaxRows) { chunks_y *= 2; rows = (ny + chunks_y - 1) / chunks_y; }
    chunks_y = (ny + rows - 1) / rows;
    const size_t smem_b = (size_t)L::kRing + ((size_t)rows + 8 * kBfVsStride) * (kBfConsumers / 32) * sizeof(double);
    const unsigned grid_b = (unsigned)((int64_t)nz * chunks_y);
    // rows requested ahead of their use: one less than the ring could hold, so that the warp on producer duty finds
    // the stage it refills released two rows ago -- with stages - 1 it waits for the slowest warp of the CTA every
    // row and the others queue up behind it (1024^3: 4.5 ms; with the slack: 3.35 ms)
    const int la = max(1, min(L::kStages - 1, tuning().stencil_la > 0 ? tuning().stencil_la : L::kStages - 2));
    Recip3 rh;
    rh.x = 1.0 / dx; rh.y = 1.0 / dy; rh.z = 1.0 / dz;
#define PTV_BF_LAUNCH(UNIT)                                                                                       \
  do {                                                                                                            \
    auto kern = div_flux_bulk_kernel<Tf, UNIT>;                                                                   \
    PTV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));               \
    kern<<<grid_b, kBfThreads, smem_b, s>>>((const Tf*)u, (const Tf*)v, (const Tf*)w, mask, nx, ny, nz, dx, dy,   \
                                            dz, rh, (const Tf*)w_below, (const Tf*)w_above, mask_above, (Tf*)div, \
                                            stats, qxy, qxz, qyz, rows, chunks_y, la);                            \
  } while (0)
    if (unit) PTV_BF_LAUNCH(true); else PTV_BF_LAUNCH(false);
#undef PTV_BF_LAUNCH
    count_launches(1);
    PTV_CUDA(cudaGetLastError());
    return PTV_OK;
  }
  const unsigned grid = (unsigned)((int64_t)nz * chunks_y);
  const size_t smem = (size_t)rows * (kSfThreads / 32) * sizeof(double);
  if (smem > 200 * 1024) { set_error("ptv_divergence_flux: ny too large for one CTA sweep"); return PTV_ERR_INVALID; }
#define PTV_DF_ARGS (const Tf*)u, (const Tf*)v, (const Tf*)w, mask, nx, ny, nz, dx, dy, dz, (const Tf*)w_below, \
                    (const Tf*)w_above, mask_above, (Tf*)div, stats, qxy, qxz, qyz, rows, chunks_y
#define PTV_DF_LAUNCH(VEC, UNIT)                                                                                   \
  do {                                                                                                            \
    auto kern = div_flux_kernel<Tf, VEC, UNIT>;                                                                   \
    if (smem > 48 * 1024)                                                                                         \
      PTV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    kern<<<grid, kSfThreads, smem, s>>>(PTV_DF_ARGS);                                                             \
  } while (0)
  if (vec && unit) PTV_DF_LAUNCH(true, true);
  else if (vec) PTV_DF_LAUNCH(true, false);
  else if (unit) PTV_DF_LAUNCH(false, true);
  else PTV_DF_LAUNCH(false, false);
#undef PTV_DF_LAUNCH
#undef PTV_DF_ARGS
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

}  // namespace ptv

using namespace ptv;

extern "C" int ptv_divergence_flux(const void* d_u, const void* d_v, const void* d_w, const uint8_t* d_mask, int nx,
                                   int ny, int nz, double dx, double dy, double dz, const void* d_w_below,
                                   const void* d_w_above, const uint8_t* d_mask_above, int dtype, void* d_div,
                                   double* d_absdiv_sum, double* d_qxy, double* d_qxz, double* d_qyz, void* stream) {
  if (!d_u || !d_v || !d_w || !d_mask || !d_div) { set_error("ptv_divergence_flux: NULL argument"); return PTV_ERR_INVALID; }
  if (nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_divergence_flux: empty grid"); return PTV_ERR_INVALID; }
  if ((d_w_above == nullptr) != (d_mask_above == nullptr)) { set_error("ptv_divergence_flux: w_above and mask_above go together"); return PTV_ERR_INVALID; }
  const bool any_q = d_qxy || d_qxz || d_qyz;
  if (any_q && !(d_qxy && d_qxz && d_qyz)) { set_error("ptv_divergence_flux: give all three flux outputs or none"); return PTV_ERR_INVALID; }
  if (dtype == PTV_F32)
    return launch_div_flux<float>(d_u, d_v, d_w, d_mask, nx, ny, nz, dx, dy, dz, d_w_below, d_w_above, d_mask_above,
                                  d_div, d_absdiv_sum, d_qxy, d_qxz, d_qyz, (cudaStream_t)stream);
  if (dtype == PTV_F64)
    return launch_div_flux<double>(d_u, d_v, d_w, d_mask, nx, ny, nz, dx, dy, dz, d_w_below, d_w_above, d_mask_above,
                                   d_div, d_absdiv_sum, d_qxy, d_qxz, d_qyz, (cudaStream_t)stream);
  set_error("ptv_divergence_flux: bad dtype");
  return PTV_ERR_INVALID;
}

namespace ptv {
// 2^22 pseudo-random numerators per round (exponents spread over +-40 binades around 1, plus exact zeros): counts the
// results of div_by_spacing that differ from the IEEE division in any bit
__global__ void __launch_bounds__(256) division_selftest_kernel(double h, double rh, uint64_t seed, int rounds,
                                                                unsigned long long* __restrict__ bad) {
  uint64_t st = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x + 1));
  unsigned long long mism = 0;
  for (int i = 0; i < rounds; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;  // xorshift64
    const uint64_t mant = st & 0x000fffffffffffffull;
    const uint64_t ex = 1023 - 40 + ((st >> 52) % 81);
    const uint64_t sg = (st >> 63) << 63;
    double t = __longlong_as_double((long long)(sg | (ex << 52) | mant));
    if ((st & 0xff000) == 0) t = __longlong_as_double((long long)sg);  // a signed zero now and then
    const double a = div_by_spacing(t, h, rh), b = __ddiv_rn(t, h);
    mism += __double_as_longlong(a) != __double_as_longlong(b);
  }
  if (mism) atomicAdd(bad, mism);
}
}  // namespace ptv

extern "C" int ptv_selftest_division(double h, int64_t n, uint64_t seed, int64_t* mismatches) {
  if (!mismatches || n <= 0) { set_error("ptv_selftest_division: bad argument"); return PTV_ERR_INVALID; }
  if (!spacing_ok(h)) { set_error("ptv_selftest_division: divisor outside the fast path's range"); return PTV_ERR_INVALID; }
  unsigned long long* d_bad = nullptr;
  PTV_CUDA(cudaMalloc(&d_bad, sizeof(unsigned long long)));
  cudaError_t e = cudaMemset(d_bad, 0, sizeof(unsigned long long));
  const int threads = 256, blocks = 148 * 8;
  const int rounds = (int)((n + (int64_t)threads * blocks - 1) / ((int64_t)threads * blocks));
  if (e == cudaSuccess) {
    division_selftest_kernel<<<blocks, threads>>>(h, 1.0 / h, seed, rounds, d_bad);
    count_launches(1);
    e = cudaGetLastError();
  }
  unsigned long long bad = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
  cudaFree(d_bad);
  PTV_CUDA(e);
  *mismatches = (int64_t)bad;
  return PTV_OK;
}
