// HBM-bound grid kernels that sit either side of the interpolation: mask resampling gather,
// no-slip boundary-voxel extraction, solid zeroing, the masked finite-volume divergence stencil
// and the plane-flux / mean|div| reductions.  Each is a single streaming pass; the algorithmic
// bytes per voxel are listed in DESIGN.md.
#include <math.h>

#include "ptv_internal.cuh"

namespace ptv {

int launch_strain_vorticity_bulk(const float* u, const float* v, const float* w, const uint8_t* mask, int nx, int ny,
                                 int nz, double dx, double dy, double dz, float* strain, float* vort, const float* below,
                                 const float* above, cudaStream_t s);

// ------------------------------------------------------------------ mask gather (a2)
// out[z,y,x] = raw[iz[z], iy[y], ix[x]] != 0, 0 where any index is -1 (out of bounds ->
// fill_value 0, interpolator.py:230-231).  A warp owns 512 consecutive x of one output row: lane l handles
// x = 128 j + 4 l .. + 3 for j = 0..3, so every warp-wide byte gather and every 4-byte store touches one
// 128-byte line, and a thread has sixteen independent gathers in flight (the kernel is bound by the latency
// of its dependent loads, index -> byte, not by bytes).
static constexpr int kMgPer = 16;

__global__ void __launch_bounds__(256) mask_gather_kernel(const uint8_t* __restrict__ raw, int rnx, int rny,
                                                           const int32_t* __restrict__ mx,
                                                           const int32_t* __restrict__ my,
                                                           const int32_t* __restrict__ mz, int nx, int ny,
                                                           int nz, uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gx = (nx + 32 * kMgPer - 1) / (32 * kMgPer);  // 512-wide chunks per row
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= (int64_t)gx * ny * nz) return;
  const int xc = (int)(wid % gx);
  const int64_t row = wid / gx;
  const int y = (int)(row % ny), z = (int)(row / ny);
  const int xw = xc * 32 * kMgPer;
  const bool vec = (nx & 3) == 0 && (reinterpret_cast<uintptr_t>(mx) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  int sx[kMgPer];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = xw + 128 * j + 4 * lane;
    if (vec && x + 4 <= nx) {
      const int4 t4 = __ldg(reinterpret_cast<const int4*>(mx + x));
      sx[4 * j] = t4.x; sx[4 * j + 1] = t4.y; sx[4 * j + 2] = t4.z; sx[4 * j + 3] = t4.w;
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b) sx[4 * j + b] = x + b < nx ? __ldg(mx + x + b) : -1;
    }
  }
  const int sz = __ldg(mz + z), sy = __ldg(my + y);
  const bool rowok = sz >= 0 && sy >= 0;
  const uint8_t* src = raw + ((int64_t)(rowok ? sz : 0) * rny + (rowok ? sy : 0)) * rnx;
  uint8_t v[kMgPer];
#pragma unroll
  for (int i = 0; i < kMgPer; ++i) v[i] = (rowok && sx[i] >= 0) ? __ldg(src + sx[i]) : (uint8_t)0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = xw + 128 * j + 4 * lane;
    uint32_t w = 0u;
#pragma unroll
    for (int b = 0; b < 4; ++b) w |= (v[4 * j + b] != 0 ? 1u : 0u) << (8 * b);
    uint8_t* dst = out + row * nx + x;
    if (vec && x + 4 <= nx) {
      *reinterpret_cast<uint32_t*>(dst) = w;
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (x + b < nx) dst[b] = (uint8_t)(w >> (8 * b));
    }
  }
}

// ------------------------------------------------------------------ boundary voxels (a3)
// Solid voxels within `thickness` 6-connected dilation steps of fluid (scipy.ndimage.binary_dilation with
// border_value = 0, interpolator.py:256-262), as an ordered index list.  Everything runs on a BIT-PACKED copy
// of the mask (32 voxels of one x-row per word, rows padded to whole words): packing reads the byte mask
// once, a dilation step is seven word reads and one word write per 32 voxels, the flags are
// dilated & ~mask on words, and the ordered compaction walks set bits.
__global__ void __launch_bounds__(256) pack_mask_kernel(const uint8_t* __restrict__ mask, int nx, int wx, int64_t nrows,
                                                         uint32_t* __restrict__ bits, int* __restrict__ any_fluid) {
  const int64_t wid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (wid >= nrows * wx) return;
  const int64_t row = wid / wx;
  const int xw = (int)(wid % wx);
  const uint8_t* src = mask + row * nx + (int64_t)xw * 32;
  const int valid = min(32, nx - xw * 32);
  uint32_t w = 0u;
  if (valid == 32 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {  // four mask bytes -> four bits (multiply gathers the 0/1 bytes into a nibble)
      const uint32_t b = __vcmpne4(reinterpret_cast<const uint32_t*>(src)[q], 0u) & 0x01010101u;
      w |= ((b * 0x01020408u) >> 24 & 0xfu) << (4 * q);
    }
  } else {
    for (int j = 0; j < valid; ++j) w |= (src[j] != 0 ? 1u : 0u) << j;
  }
  bits[wid] = w;
  if (w != 0u && *any_fluid == 0) *any_fluid = 1;
}

// One 6-connected dilation step on the packed volume; `fill` != 0: every voxel becomes set if the input has
// any set voxel at all (binary_dilation with iterations < 1 repeats until nothing changes).
__global__ void __launch_bounds__(256) dilate_bits_kernel(const uint32_t* __restrict__ in, int nx, int wx, int ny, int nz,
                                                           uint32_t* __restrict__ out, const int* __restrict__ fill) {
  const int64_t wid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nwords = (int64_t)wx * ny * nz;
  if (wid >= nwords) return;
  const int xw = (int)(wid % wx);
  const int y = (int)((wid / wx) % ny);
  const int z = (int)(wid / ((int64_t)wx * ny));
  const uint32_t tail = (xw == wx - 1 && (nx & 31)) ? ((1u << (nx & 31)) - 1u) : 0xffffffffu;  // bits inside the row
  if (fill != nullptr) {
    out[wid] = *fill ? tail : 0u;
    return;
  }
  const uint32_t c = in[wid];
  uint32_t r = c | (c << 1) | (c >> 1);
  if (xw > 0) r |= in[wid - 1] >> 31;
  if (xw + 1 < wx) r |= in[wid + 1] << 31;
  if (y > 0) r |= in[wid - wx];
  if (y + 1 < ny) r |= in[wid + wx];
  const int64_t sz = (int64_t)wx * ny;
  if (z > 0) r |= in[wid - sz];
  if (z + 1 < nz) r |= in[wid + sz];
  out[wid] = r & tail;
}

static constexpr int kCompThreads = 256;
static constexpr int kCompItems = 8;                           // words per thread
static constexpr int kCompTile = kCompThreads * kCompItems;    // words per tile

__device__ __forceinline__ int comp_block_scan(int v, int* total_out) {
  __shared__ int warp_tot[kCompThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int w2 = 0; w2 < kCompThreads / 32; ++w2) {
    int t = warp_tot[w2];
    if (w2 < wid) woff += t;
    tot += t;
  }
  __syncthreads();
  *total_out = tot;
  return woff + inc - v;
}

// flag words = dilated & ~mask.  pass 0: flag words + per-tile counts; pass 1: ordered (C order) index write.
template <int PASS>
__global__ void __launch_bounds__(kCompThreads) boundary_compact_kernel(
    const uint32_t* __restrict__ dil, const uint32_t* __restrict__ mbits, uint32_t* __restrict__ flags, int64_t nwords,
    int nx, int wx, int64_t* __restrict__ tile_counts, const int64_t* __restrict__ tile_offsets,
    int64_t* __restrict__ indices, int64_t cap) {
  const int64_t base = (int64_t)blockIdx.x * kCompTile + (int64_t)threadIdx.x * kCompItems;
  uint32_t f[kCompItems];
  int c = 0;
#pragma unroll
  for (int j = 0; j < kCompItems; ++j) {
    const int64_t i = base + j;
    f[j] = 0u;
    if (i < nwords) f[j] = PASS == 0 ? (dil[i] & ~mbits[i]) : flags[i];
    c += __popc(f[j]);
  }
  int tot;
  const int ex = comp_block_scan(c, &tot);
  if (PASS == 0) {
#pragma unroll
    for (int j = 0; j < kCompItems; ++j)
      if (base + j < nwords) flags[base + j] = f[j];
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
  } else {
    int64_t o = tile_offsets[blockIdx.x] + ex;
#pragma unroll
    for (int j = 0; j < kCompItems; ++j) {
      uint32_t w = f[j];
      if (w == 0u) continue;
      const int64_t i = base + j;
      const int64_t vox0 = (i / wx) * nx + (int64_t)(i % wx) * 32;  // linear index of the word's first voxel
      while (w != 0u) {
        const int b = __ffs((int)w) - 1;
        w &= w - 1u;
        if (o < cap) indices[o] = vox0 + b;
        ++o;
      }
    }
  }
}

// exclusive scan of int64 tile counts by one block (tile count = nvox/2048 <= ~0.5M at 1024^3)
__global__ void __launch_bounds__(1024) scan_i64_kernel(const int64_t* __restrict__ in, int64_t n,
                                                         int64_t* __restrict__ out, int64_t* __restrict__ total) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < n ? in[i] : 0;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    int64_t woff = 0, tot = 0;
    for (int w2 = 0; w2 < 32; ++w2) {
      const int64_t t = warp_tot[w2];
      if (w2 < wid) woff += t;
      tot += t;
    }
    const int64_t carry = carry_s;
    if (i < n) out[i] = carry + woff + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

// ------------------------------------------------------------------ solid zeroing (a11)
template <typename Tf>
__global__ void __launch_bounds__(256) apply_mask_kernel(Tf* __restrict__ u, Tf* __restrict__ v, Tf* __restrict__ w,
                                                          const uint8_t* __restrict__ mask, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // main.py:195-207: NaN -> 0, then zero where the mask is solid
  const bool solid = mask != nullptr && mask[i] == 0;
  Tf a = u[i], b = v[i], c = w[i];
  if (solid || a != a) u[i] = (Tf)0;
  if (solid || b != b) v[i] = (Tf)0;
  if (solid || c != c) w[i] = (Tf)0;
}

// ------------------------------------------------------------------ divergence (a12)
// Closed form of physics.py:26-53, evaluated in float64 with the reference's
// operation order and no FMA contraction, so float64 fields reproduce NumPy bit for bit:
//   F+ = i == n-1 ? f[i] : (m[i+1] ? (f[i] + f[i+1]) / 2 : 0)
//   F- = i == 0   ? f[0] : (m[i]   ? (f[i-1] + f[i]) / 2 : 0)
//   div = ((F+x - F-x)/dx + (F+y - F-y)/dy) + (F+z - F-z)/dz
template <typename Tf>
__global__ void __launch_bounds__(256) divergence_kernel(
    const Tf* __restrict__ u, const Tf* __restrict__ v, const Tf* __restrict__ w,
    const uint8_t* __restrict__ mask, int nx, int ny, int nz, double dx, double dy, double dz,
    const Tf* __restrict__ w_below, const Tf* __restrict__ w_above, const uint8_t* __restrict__ mask_above,
    Tf* __restrict__ div, double* __restrict__ absdiv_sum) {
  const int64_t n = (int64_t)nx * ny * nz;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double absval = 0.0, cnt = 0.0;
  if (i < n) {
    const int x = (int)(i % nx);
    const int y = (int)((i / nx) % ny);
    const int z = (int)(i / ((int64_t)nx * ny));
    const int64_t sy = nx, sz = (int64_t)nx * ny;
    const bool m = mask[i] != 0;
    // x (u, axis 2)
    const double uc = (double)u[i];
    double fp, fm;
    if (x == nx - 1) fp = uc;
    else fp = mask[i + 1] ? __ddiv_rn(__dadd_rn(uc, (double)u[i + 1]), 2.0) : 0.0;
    if (x == 0) fm = uc;
    else fm = m ? __ddiv_rn(__dadd_rn((double)u[i - 1], uc), 2.0) : 0.0;
    const double tx = __ddiv_rn(__dsub_rn(fp, fm), dx);
    // y (v, axis 1)
    const double vc = (double)v[i];
    if (y == ny - 1) fp = vc;
    else fp = mask[i + sy] ? __ddiv_rn(__dadd_rn(vc, (double)v[i + sy]), 2.0) : 0.0;
    if (y == 0) fm = vc;
    else fm = m ? __ddiv_rn(__dadd_rn((double)v[i - sy], vc), 2.0) : 0.0;
    const double ty = __ddiv_rn(__dsub_rn(fp, fm), dy);
    // z (w, axis 0) -- slab halos stand in for the planes owned by the z-neighbours
    const double wc = (double)w[i];
    const int64_t pl = (int64_t)y * nx + x;
    if (z == nz - 1) {
      if (w_above != nullptr) fp = mask_above[pl] ? __ddiv_rn(__dadd_rn(wc, (double)w_above[pl]), 2.0) : 0.0;
      else fp = wc;
    } else {
      fp = mask[i + sz] ? __ddiv_rn(__dadd_rn(wc, (double)w[i + sz]), 2.0) : 0.0;
    }
    if (z == 0) {
      if (w_below != nullptr) fm = m ? __ddiv_rn(__dadd_rn((double)w_below[pl], wc), 2.0) : 0.0;
      else fm = wc;
    } else {
      fm = m ? __ddiv_rn(__dadd_rn((double)w[i - sz], wc), 2.0) : 0.0;
    }
    const double tz = __ddiv_rn(__dsub_rn(fp, fm), dz);
    const double d = __dadd_rn(__dadd_rn(tx, ty), tz);
    div[i] = (Tf)d;
    if (m) {
      absval = fabs((double)(Tf)d);
      cnt = 1.0;
    }
  }
  if (absdiv_sum != nullptr) {  // uniform branch: block reduction then one atomic pair per block
    __shared__ double sh[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      absval += __shfl_xor_sync(0xffffffffu, absval, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = absval; sh[1][wid] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, c = 0.0;
      for (int w2 = 0; w2 < 8; ++w2) { a += sh[0][w2]; c += sh[1][w2]; }
      if (c > 0.0) {
        atomicAdd(&absdiv_sum[0], a);
        atomicAdd(&absdiv_sum[1], c);
      }
    }
  }
}

// ------------------------------------------------------------------ strain rate / vorticity (N3)
// velocity_analysis.py:10-63 (compute_strain_rate) and :94-120 (compute_vorticity): nine np.gradient
// derivatives (second-order central inside, first-order one-sided at the edges, uniform spacing),
// then the shear-rate magnitude sqrt(0.5*(exx^2+eyy^2+ezz^2) + exy^2 + exz^2 + eyz^2) and |curl u|,
// zeroed in solid voxels.  float64 arithmetic in NumPy's operation order (bit-identical for float64
// fields).  HBM-bound: 13 B read + 4 B per output written per voxel.
// num / den with the IEEE result.  When den is a power of two (unit spacing: 1 and 2) its reciprocal is
// exact and the product is bit-identical to the quotient, so the division is skipped.
struct Divisor {
  double den, inv;
  bool pow2;
};
__host__ __device__ inline Divisor make_divisor(double den) {
  Divisor d;
  d.den = den;
  d.inv = 1.0 / den;
  int e;
  const double m = frexp(fabs(den), &e);
  d.pow2 = (m == 0.5) && e > -1000 && e < 1000;
  return d;
}
__device__ __forceinline__ double grad_div(double num, const Divisor& d) {
  if (d.pow2) return __dmul_rn(num, d.inv);
  if (num == 0.0) return d.den > 0.0 ? num : -num;  // signed zero without the division slow path
  return __ddiv_rn(num, d.den);
}

// kPow2: every divisor of the launch is a power of two (unit or power-of-two spacing): the run-time tests go
template <bool kPow2>
__device__ __forceinline__ double grad_div_t(double num, const Divisor& d) {
  if (kPow2) return __dmul_rn(num, d.inv);
  return grad_div(num, d);
}

struct Divisors6 { Divisor d[6]; };

template <typename Tf>
__device__ __forceinline__ double np_gradient(const Tf* __restrict__ f, int64_t i, int64_t stride, int pos, int n,
                                              const Divisor& h1, const Divisor& h2) {
  if (pos == 0) return grad_div(__dsub_rn((double)f[i + stride], (double)f[i]), h1);
  if (pos == n - 1) return grad_div(__dsub_rn((double)f[i], (double)f[i - stride]), h1);
  return grad_div(__dsub_rn((double)f[i + stride], (double)f[i - stride]), h2);
}

template <typename Tf>
__global__ void __launch_bounds__(256) strain_vorticity_kernel(const Tf* __restrict__ u, const Tf* __restrict__ v,
                                                                const Tf* __restrict__ w,
                                                                const uint8_t* __restrict__ mask, int nx, int ny, int nz,
                                                                const Divisors6 dv6,
                                                                Tf* __restrict__ strain, Tf* __restrict__ vort) {
  const int64_t n = (int64_t)nx * ny * nz;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask != nullptr && mask[i] == 0) {  // velocity_analysis.py:58-61,116-117
    if (strain) strain[i] = (Tf)0;
    if (vort) vort[i] = (Tf)0;
    return;
  }
  const int x = (int)(i % nx);
  const int y = (int)((i / nx) % ny);
  const int z = (int)(i / ((int64_t)nx * ny));
  const int64_t sy = nx, sz = (int64_t)nx * ny;
  // edge spacing h and interior spacing 2h (np.gradient: (f[2:] - f[:-2]) / (2. * h))
  const Divisor &x1 = dv6.d[0], &x2 = dv6.d[1], &y1 = dv6.d[2], &y2 = dv6.d[3], &z1 = dv6.d[4], &z2 = dv6.d[5];
  const double du_dx = np_gradient(u, i, 1, x, nx, x1, x2), du_dy = np_gradient(u, i, sy, y, ny, y1, y2),
               du_dz = np_gradient(u, i, sz, z, nz, z1, z2);
  const double dv_dx = np_gradient(v, i, 1, x, nx, x1, x2), dv_dy = np_gradient(v, i, sy, y, ny, y1, y2),
               dv_dz = np_gradient(v, i, sz, z, nz, z1, z2);
  const double dw_dx = np_gradient(w, i, 1, x, nx, x1, x2), dw_dy = np_gradient(w, i, sy, y, ny, y1, y2),
               dw_dz = np_gradient(w, i, sz, z, nz, z1, z2);
  if (strain) {
    const double exx = __dmul_rn(2.0, du_dx), eyy = __dmul_rn(2.0, dv_dy), ezz = __dmul_rn(2.0, dw_dz);
    const double exy = __dadd_rn(du_dy, dv_dx), exz = __dadd_rn(du_dz, dw_dx), eyz = __dadd_rn(dv_dz, dw_dy);
    const double diag = __dmul_rn(0.5, __dadd_rn(__dadd_rn(__dmul_rn(exx, exx), __dmul_rn(eyy, eyy)), __dmul_rn(ezz, ezz)));
    const double s2 = __dadd_rn(__dadd_rn(__dadd_rn(diag, __dmul_rn(exy, exy)), __dmul_rn(exz, exz)), __dmul_rn(eyz, eyz));
    strain[i] = (Tf)sqrt(s2);
  }
  if (vort) {
    const double vx = __dsub_rn(dw_dy, dv_dz), vy = __dsub_rn(du_dz, dw_dx), vz = __dsub_rn(dv_dx, du_dy);
    vort[i] = (Tf)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
  }
}

// Vectorised variant (nx % 4 == 0, 16-byte aligned rows): one thread = 4 consecutive x; every field
// contributes 5 vector loads (centre, y-1, y+1, z-1, z+1) and 2 scalar x-halo loads per 4 voxels instead
// of 28 scalar loads.  Same arithmetic, same operation order.
template <typename Tf>
__device__ __forceinline__ void ld4(const Tf* __restrict__ p, double (&o)[4]) {
  if (sizeof(Tf) == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  } else {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *(reinterpret_cast<const double2*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
}

template <typename Tf>
struct Grad4 { double dx[4], dy[4], dz[4]; };

// The 7-point neighbourhood of 4 consecutive x as raw loads.  All of a thread's loads (3 fields x 5 vectors
// + 2 scalars, and the mask) are issued before anything is converted or tested, so a thread has ~330 B in
// flight instead of walking mask -> field -> field -> field latency by latency.
template <typename Tf> struct Raw4 { Tf v[4]; };
template <typename Tf>
__device__ __forceinline__ Raw4<Tf> ldraw4(const Tf* __restrict__ p) {
  // volatile asm: the compiler must not sink the loads below the mask test that follows them
  Raw4<Tf> r;
  if (sizeof(Tf) == 4) {
    float a, b, c, d;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p));
    r.v[0] = (Tf)a; r.v[1] = (Tf)b; r.v[2] = (Tf)c; r.v[3] = (Tf)d;
  } else {
    double a, b, c, d;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(c), "=d"(d) : "l"(p + 2));
    r.v[0] = (Tf)a; r.v[1] = (Tf)b; r.v[2] = (Tf)c; r.v[3] = (Tf)d;
  }
  return r;
}
template <typename Tf> struct Stencil4 { Raw4<Tf> c, ym, yp, zm, zp; Tf xl, xr; };

template <typename Tf>
__device__ __forceinline__ Stencil4<Tf> load_stencil4(const Tf* __restrict__ f, int64_t i, int x, int y, int z, int nx,
                                                     int ny, int nz) {
  const int64_t sy = nx, sz = (int64_t)nx * ny;
  Stencil4<Tf> s;
  s.c = ldraw4(f + i);
  s.ym = ldraw4(f + (y > 0 ? i - sy : i));       // (edge rows re-read the centre: unused below)
  s.yp = ldraw4(f + (y < ny - 1 ? i + sy : i));
  s.zm = ldraw4(f + (z > 0 ? i - sz : i));
  s.zp = ldraw4(f + (z < nz - 1 ? i + sz : i));
  s.xl = __ldg(f + (x > 0 ? i - 1 : i));
  s.xr = __ldg(f + (x + 4 < nx ? i + 4 : i));
  return s;
}

template <typename Tf, bool kPow2>
__device__ __forceinline__ Grad4<Tf> gradients4(const Stencil4<Tf>& s, int x, int y, int z, int nx, int ny, int nz,
                                                const Divisors6& dv) {
  double e[6], c[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j] = (double)s.c.v[j];
  e[0] = (double)s.xl;
  e[5] = (double)s.xr;
#pragma unroll
  for (int j = 0; j < 4; ++j) e[j + 1] = c[j];
  Grad4<Tf> g;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pos = x + j;
    if (pos == 0) g.dx[j] = grad_div_t<kPow2>(__dsub_rn(e[j + 2], e[j + 1]), dv.d[0]);
    else if (pos == nx - 1) g.dx[j] = grad_div_t<kPow2>(__dsub_rn(e[j + 1], e[j]), dv.d[0]);
    else g.dx[j] = grad_div_t<kPow2>(__dsub_rn(e[j + 2], e[j]), dv.d[1]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double a = (double)s.ym.v[j], b = (double)s.yp.v[j];
    if (y == 0) g.dy[j] = grad_div_t<kPow2>(__dsub_rn(b, c[j]), dv.d[2]);
    else if (y == ny - 1) g.dy[j] = grad_div_t<kPow2>(__dsub_rn(c[j], a), dv.d[2]);
    else g.dy[j] = grad_div_t<kPow2>(__dsub_rn(b, a), dv.d[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double a = (double)s.zm.v[j], b = (double)s.zp.v[j];
    if (z == 0) g.dz[j] = grad_div_t<kPow2>(__dsub_rn(b, c[j]), dv.d[4]);
    else if (z == nz - 1) g.dz[j] = grad_div_t<kPow2>(__dsub_rn(c[j], a), dv.d[4]);
    else g.dz[j] = grad_div_t<kPow2>(__dsub_rn(b, a), dv.d[5]);
  }
  return g;
}

template <typename Tf, bool kPow2>
__global__ void __launch_bounds__(256) strain_vorticity_vec4_kernel(const Tf* __restrict__ u, const Tf* __restrict__ v,
                                                                     const Tf* __restrict__ w,
                                                                     const uint8_t* __restrict__ mask, int nx, int ny,
                                                                     int nz, const Divisors6 dv6, Tf* __restrict__ strain,
                                                                     Tf* __restrict__ vort) {
  const int gx = nx >> 2;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)gx * ny * nz) return;
  const int x = (int)(gid % gx) * 4;
  const int64_t row = gid / gx;
  const int y = (int)(row % ny), z = (int)(row / ny);
  const int64_t i = row * nx + x;
  uchar4 m = make_uchar4(1, 1, 1, 1);
  if (mask != nullptr) m = __ldg(reinterpret_cast<const uchar4*>(mask + i));
  const Stencil4<Tf> su = load_stencil4(u, i, x, y, z, nx, ny, nz);
  const Stencil4<Tf> sv = load_stencil4(v, i, x, y, z, nx, ny, nz);
  const Stencil4<Tf> sw = load_stencil4(w, i, x, y, z, nx, ny, nz);
  const bool mk[4] = {m.x != 0, m.y != 0, m.z != 0, m.w != 0};
  double so[4] = {0.0, 0.0, 0.0, 0.0}, vo[4] = {0.0, 0.0, 0.0, 0.0};
  if (mk[0] || mk[1] || mk[2] || mk[3]) {
    const Grad4<Tf> gu = gradients4<Tf, kPow2>(su, x, y, z, nx, ny, nz, dv6);
    const Grad4<Tf> gv = gradients4<Tf, kPow2>(sv, x, y, z, nx, ny, nz, dv6);
    const Grad4<Tf> gw = gradients4<Tf, kPow2>(sw, x, y, z, nx, ny, nz, dv6);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!mk[j]) continue;
      if (strain) {
        const double exx = __dmul_rn(2.0, gu.dx[j]), eyy = __dmul_rn(2.0, gv.dy[j]), ezz = __dmul_rn(2.0, gw.dz[j]);
        const double exy = __dadd_rn(gu.dy[j], gv.dx[j]), exz = __dadd_rn(gu.dz[j], gw.dx[j]),
                     eyz = __dadd_rn(gv.dz[j], gw.dy[j]);
        const double diag = __dmul_rn(0.5, __dadd_rn(__dadd_rn(__dmul_rn(exx, exx), __dmul_rn(eyy, eyy)), __dmul_rn(ezz, ezz)));
        so[j] = sqrt(__dadd_rn(__dadd_rn(__dadd_rn(diag, __dmul_rn(exy, exy)), __dmul_rn(exz, exz)), __dmul_rn(eyz, eyz)));
      }
      if (vort) {
        const double vx = __dsub_rn(gw.dy[j], gv.dz[j]), vy = __dsub_rn(gu.dz[j], gw.dx[j]), vz = __dsub_rn(gv.dx[j], gu.dy[j]);
        vo[j] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
      }
    }
  }
  auto st4 = [&](Tf* p, const double (&d)[4]) {
    if (sizeof(Tf) == 4) {
      *reinterpret_cast<float4*>(p) = make_float4((float)d[0], (float)d[1], (float)d[2], (float)d[3]);
    } else {
      *reinterpret_cast<double2*>(p) = make_double2(d[0], d[1]);
      *(reinterpret_cast<double2*>(p) + 1) = make_double2(d[2], d[3]);
    }
  };
  if (strain) st4(strain + i, so);
  if (vort) st4(vort + i, vo);
}

}  // namespace ptv

using namespace ptv;

extern "C" int ptv_strain_vorticity(const void* d_u, const void* d_v, const void* d_w, const uint8_t* d_mask, int nx,
                                    int ny, int nz, double dx, double dy, double dz, int dtype, void* d_strain,
                                    void* d_vorticity, void* stream) {
  if (!d_u || !d_v || !d_w) { set_error("ptv_strain_vorticity: NULL argument"); return PTV_ERR_INVALID; }
  if (!d_strain && !d_vorticity) { set_error("ptv_strain_vorticity: nothing to compute"); return PTV_ERR_INVALID; }
  if (nx < 2 || ny < 2 || nz < 2) {
    // np.gradient: "Shape of array too small to calculate a numerical gradient, at least 2 elements are required"
    set_error("Shape of array too small to calculate a numerical gradient, at least (edge_order + 1) elements are required.");
    return PTV_ERR_INVALID;
  }
  if (dtype == PTV_F32) {  // z-marching kernel over bulk-copied plane tiles (strain_bulk.cu) where shape and alignment allow
    const int rc = launch_strain_vorticity_bulk((const float*)d_u, (const float*)d_v, (const float*)d_w, d_mask, nx, ny, nz, dx,
                                                dy, dz, (float*)d_strain, (float*)d_vorticity, nullptr, nullptr, (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  const int64_t n = (int64_t)nx * ny * nz;
  const unsigned nb = (unsigned)((n + 255) / 256);
  Divisors6 dv6;
  dv6.d[0] = make_divisor(dx); dv6.d[1] = make_divisor(2.0 * dx);
  dv6.d[2] = make_divisor(dy); dv6.d[3] = make_divisor(2.0 * dy);
  dv6.d[4] = make_divisor(dz); dv6.d[5] = make_divisor(2.0 * dz);
  const auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = (nx % 4 == 0) && al16(d_u) && al16(d_v) && al16(d_w) && al16(d_strain) && al16(d_vorticity) &&
                   (d_mask == nullptr || (reinterpret_cast<uintptr_t>(d_mask) & 3) == 0);
  const unsigned nbv = (unsigned)((n / 4 + 255) / 256);
  bool pow2 = true;
  for (int c = 0; c < 6; ++c) pow2 = pow2 && dv6.d[c].pow2;
#define PTV_SV4(T, P2)                                                                                        \
  strain_vorticity_vec4_kernel<T, P2><<<nbv, 256, 0, (cudaStream_t)stream>>>(                                \
      (const T*)d_u, (const T*)d_v, (const T*)d_w, d_mask, nx, ny, nz, dv6, (T*)d_strain, (T*)d_vorticity)
  if (vec && dtype == PTV_F32) {
    if (pow2) PTV_SV4(float, true); else PTV_SV4(float, false);
  } else if (vec && dtype == PTV_F64) {
    if (pow2) PTV_SV4(double, true); else PTV_SV4(double, false);
  }
#undef PTV_SV4
  else if (dtype == PTV_F32)
    strain_vorticity_kernel<float><<<nb, 256, 0, (cudaStream_t)stream>>>(
        (const float*)d_u, (const float*)d_v, (const float*)d_w, d_mask, nx, ny, nz, dv6, (float*)d_strain,
        (float*)d_vorticity);
  else if (dtype == PTV_F64)
    strain_vorticity_kernel<double><<<nb, 256, 0, (cudaStream_t)stream>>>(
        (const double*)d_u, (const double*)d_v, (const double*)d_w, d_mask, nx, ny, nz, dv6, (double*)d_strain,
        (double*)d_vorticity);
  else { set_error("ptv_strain_vorticity: bad dtype"); return PTV_ERR_INVALID; }
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

namespace ptv {

// ------------------------------------------------------------------ flux profiles (a13)
// q_xy[z] += sum_{y,x} w   (plot_flux.py:6-8)
template <typename Tf>
__global__ void __launch_bounds__(256) flux_xy_kernel(const Tf* __restrict__ w, int64_t plane, int chunks,
                                                       double* __restrict__ qxy) {
  const int z = blockIdx.x / chunks, ch = blockIdx.x % chunks;
  const int64_t per = (plane + chunks - 1) / chunks;
  const int64_t lo = (int64_t)ch * per, hi = min(plane, lo + per);
  const Tf* src = w + (int64_t)z * plane;
  double acc = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc += (double)src[i];
  __shared__ double sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w2 = 0; w2 < 8; ++w2) a += sh[w2];
    atomicAdd(&qxy[z], a);
  }
}

// q_xz[y] += sum_{z,x} v   (plot_flux.py:10-12): one warp per (z, y) row
template <typename Tf>
__global__ void __launch_bounds__(256) flux_xz_kernel(const Tf* __restrict__ v, int nx, int ny, int nz,
                                                       double* __restrict__ qxz) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (int64_t)ny * nz) return;
  const int y = (int)(row % ny);
  const Tf* src = v + row * nx;
  double acc = 0.0;
  for (int x = threadIdx.x & 31; x < nx; x += 32) acc += (double)src[x];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&qxz[y], acc);
}

// q_yz[x] += sum_{z,y} u   (plot_flux.py:14-16): one thread per x, marching over y of one plane
template <typename Tf>
__global__ void __launch_bounds__(256) flux_yz_kernel(const Tf* __restrict__ u, int nx, int ny,
                                                       double* __restrict__ qyz) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  const int z = blockIdx.y;
  if (x >= nx) return;
  const Tf* src = u + (int64_t)z * ny * nx + x;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int y = 0;
  for (; y + 3 < ny; y += 4) {
    a0 += (double)src[(int64_t)y * nx];
    a1 += (double)src[(int64_t)(y + 1) * nx];
    a2 += (double)src[(int64_t)(y + 2) * nx];
    a3 += (double)src[(int64_t)(y + 3) * nx];
  }
  for (; y < ny; ++y) a0 += (double)src[(int64_t)y * nx];
  atomicAdd(&qyz[x], (a0 + a1) + (a2 + a3));
}

}  // namespace ptv

using namespace ptv;

extern "C" int ptv_mask_gather(const uint8_t* d_mask_raw, int rnx, int rny, int rnz, const int32_t* d_ix, int nx,
                               const int32_t* d_iy, int ny, const int32_t* d_iz, int nz, uint8_t* d_out,
                               void* stream) {
  if (!d_mask_raw || !d_ix || !d_iy || !d_iz || !d_out) { set_error("ptv_mask_gather: NULL argument"); return PTV_ERR_INVALID; }
  if (rnx <= 0 || rny <= 0 || rnz <= 0 || nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_mask_gather: empty array"); return PTV_ERR_INVALID; }
  const int64_t total = (int64_t)((nx + 32 * kMgPer - 1) / (32 * kMgPer)) * ny * nz * 32;  // one warp per 512 x
  mask_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_mask_raw, rnx, rny, d_ix, d_iy, d_iz, nx, ny, nz, d_out);
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

// Workspace layout (bytes): [mask bits | dilation ping | dilation pong / flag words | tile counts, offsets, total |
// any-fluid flag]; every word array holds wx * ny * nz uint32, wx = ceil(nx / 32).
static int64_t boundary_words(int nx, int ny, int nz) { return (int64_t)((nx + 31) / 32) * ny * nz; }
static int64_t boundary_tiles(int nx, int ny, int nz) { return (boundary_words(nx, ny, nz) + kCompTile - 1) / kCompTile; }

extern "C" int64_t ptv_boundary_workspace_bytes(int nx, int ny, int nz) {
  if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
  const int64_t nw = boundary_words(nx, ny, nz), nt = boundary_tiles(nx, ny, nz);
  return 3 * ((nw * 4 + 255) & ~(int64_t)255) + (2 * nt + 2) * 8 + 256;
}

// phase 0: pack, dilate, flag, count -> *h_count (synchronises); phase 1: write the ordered indices from the
// flag words phase 0 left in the workspace (nothing is recomputed).
extern "C" int ptv_boundary_voxels_ws(const uint8_t* d_mask, int nx, int ny, int nz, int thickness, void* d_work,
                                      int phase, int64_t* d_indices, int64_t cap, int64_t* h_count, void* stream_) {
  if (!d_mask || !d_work || !h_count) { set_error("ptv_boundary_voxels_ws: NULL argument"); return PTV_ERR_INVALID; }
  if (nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_boundary_voxels_ws: bad shape"); return PTV_ERR_INVALID; }
  if (phase != 0 && phase != 1) { set_error("ptv_boundary_voxels_ws: phase must be 0 or 1"); return PTV_ERR_INVALID; }
  cudaStream_t stream = (cudaStream_t)stream_;
  const int wx = (nx + 31) / 32;
  const int64_t nrows = (int64_t)ny * nz, nw = boundary_words(nx, ny, nz), nt = boundary_tiles(nx, ny, nz);
  const int64_t stride = (nw * 4 + 255) & ~(int64_t)255;
  char* wbase = reinterpret_cast<char*>(d_work);
  uint32_t* mbits = reinterpret_cast<uint32_t*>(wbase);
  uint32_t* ping = reinterpret_cast<uint32_t*>(wbase + stride);
  uint32_t* pong = reinterpret_cast<uint32_t*>(wbase + 2 * stride);
  int64_t* counts = reinterpret_cast<int64_t*>(wbase + 3 * stride);
  int64_t* offsets = counts + nt;  // nt offsets + the total
  int* any_fluid = reinterpret_cast<int*>(offsets + nt + 1);
  const unsigned gw = (unsigned)((nw + 255) / 256);
  if (phase == 0) {
    PTV_CUDA(cudaMemsetAsync(any_fluid, 0, sizeof(int), stream));
    pack_mask_kernel<<<gw, 256, 0, stream>>>(d_mask, nx, wx, nrows, mbits, any_fluid);
    const uint32_t* cur = mbits;
    int launches = 3;
    if (thickness < 1) {  // binary_dilation(iterations < 1): repeat until stable == fill the box if any fluid exists
      dilate_bits_kernel<<<gw, 256, 0, stream>>>(cur, nx, wx, ny, nz, ping, any_fluid);
      cur = ping;
      ++launches;
    }
    for (int it = 0; it < thickness; ++it) {
      uint32_t* dst = (it & 1) ? pong : ping;
      dilate_bits_kernel<<<gw, 256, 0, stream>>>(cur, nx, wx, ny, nz, dst, nullptr);
      cur = dst;
      ++launches;
    }
    // the flag words go to the buffer the last dilation did not write (pong unless it holds `cur`)
    uint32_t* flags = cur == pong ? ping : pong;
    boundary_compact_kernel<0><<<(unsigned)nt, kCompThreads, 0, stream>>>(cur, mbits, flags, nw, nx, wx, counts, nullptr,
                                                                        nullptr, 0);
    scan_i64_kernel<<<1, 1024, 0, stream>>>(counts, nt, offsets, offsets + nt);
    // remember where the flags are for phase 1: slot after the any-fluid flag
    const int which = flags == pong ? 1 : 0;
    PTV_CUDA(cudaMemcpyAsync(any_fluid + 1, &which, sizeof(int), cudaMemcpyHostToDevice, stream));
    count_launches(launches);
    PTV_CUDA(cudaGetLastError());
    PTV_CUDA(cudaMemcpyAsync(h_count, offsets + nt, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    PTV_CUDA(cudaStreamSynchronize(stream));
    return PTV_OK;
  }
  int which = 0;
  PTV_CUDA(cudaMemcpyAsync(&which, any_fluid + 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PTV_CUDA(cudaMemcpyAsync(h_count, offsets + nt, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  PTV_CUDA(cudaStreamSynchronize(stream));
  if (d_indices != nullptr && cap > 0 && *h_count > 0) {
    uint32_t* flags = which ? pong : ping;
    boundary_compact_kernel<1><<<(unsigned)nt, kCompThreads, 0, stream>>>(nullptr, nullptr, flags, nw, nx, wx, nullptr,
                                                                        offsets, d_indices, cap);
    count_launches(1);
    PTV_CUDA(cudaGetLastError());
  }
  return PTV_OK;
}

extern "C" int ptv_boundary_voxels(const uint8_t* d_mask, int nx, int ny, int nz, int thickness,
                                   int64_t* d_indices, int64_t cap, int64_t* h_count, void* stream_) {
  if (!d_mask || !h_count) { set_error("ptv_boundary_voxels: NULL argument"); return PTV_ERR_INVALID; }
  if (nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_boundary_voxels: bad shape"); return PTV_ERR_INVALID; }
  void* work = nullptr;
  cudaError_t e = cudaMalloc(&work, (size_t)ptv_boundary_workspace_bytes(nx, ny, nz));
  if (e != cudaSuccess) return cuda_fail(e, "ptv_boundary_voxels alloc", __FILE__, __LINE__);
  int rc = ptv_boundary_voxels_ws(d_mask, nx, ny, nz, thickness, work, 0, nullptr, 0, h_count, stream_);
  if (rc == PTV_OK && d_indices != nullptr && cap > 0)
    rc = ptv_boundary_voxels_ws(d_mask, nx, ny, nz, thickness, work, 1, d_indices, cap, h_count, stream_);
  if (rc == PTV_OK) cudaStreamSynchronize((cudaStream_t)stream_);
  cudaFree(work);
  return rc;
}

extern "C" int ptv_apply_mask(void* d_u, void* d_v, void* d_w, const uint8_t* d_mask, int64_t nvox, int dtype,
                              void* stream) {
  if (!d_u || !d_v || !d_w) { set_error("ptv_apply_mask: NULL argument"); return PTV_ERR_INVALID; }
  if (nvox <= 0) return PTV_OK;
  const unsigned nb = (unsigned)((nvox + 255) / 256);
  if (dtype == PTV_F32)
    apply_mask_kernel<float><<<nb, 256, 0, (cudaStream_t)stream>>>((float*)d_u, (float*)d_v, (float*)d_w, d_mask, nvox);
  else if (dtype == PTV_F64)
    apply_mask_kernel<double><<<nb, 256, 0, (cudaStream_t)stream>>>((double*)d_u, (double*)d_v, (double*)d_w, d_mask, nvox);
  else { set_error("ptv_apply_mask: bad dtype"); return PTV_ERR_INVALID; }
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

extern "C" int ptv_divergence(const void* d_u, const void* d_v, const void* d_w, const uint8_t* d_mask, int nx,
                              int ny, int nz, double dx, double dy, double dz, const void* d_w_below,
                              const void* d_w_above, const uint8_t* d_mask_above, int dtype, void* d_div,
                              double* d_absdiv_sum, void* stream) {
  if (!d_u || !d_v || !d_w || !d_mask || !d_div) { set_error("ptv_divergence: NULL argument"); return PTV_ERR_INVALID; }
  if (nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_divergence: empty grid"); return PTV_ERR_INVALID; }
  if ((d_w_above == nullptr) != (d_mask_above == nullptr)) { set_error("ptv_divergence: w_above and mask_above go together"); return PTV_ERR_INVALID; }
  const int64_t n = (int64_t)nx * ny * nz;
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (dtype == PTV_F32)
    divergence_kernel<float><<<nb, 256, 0, (cudaStream_t)stream>>>(
        (const float*)d_u, (const float*)d_v, (const float*)d_w, d_mask, nx, ny, nz, dx, dy, dz,
        (const float*)d_w_below, (const float*)d_w_above, d_mask_above, (float*)d_div, d_absdiv_sum);
  else if (dtype == PTV_F64)
    divergence_kernel<double><<<nb, 256, 0, (cudaStream_t)stream>>>(
        (const double*)d_u, (const double*)d_v, (const double*)d_w, d_mask, nx, ny, nz, dx, dy, dz,
        (const double*)d_w_below, (const double*)d_w_above, d_mask_above, (double*)d_div, d_absdiv_sum);
  else { set_error("ptv_divergence: bad dtype"); return PTV_ERR_INVALID; }
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

template <typename Tf>
static int flux_launch(const void* u, const void* v, const void* w, int nx, int ny, int nz, double* qxy,
                       double* qxz, double* qyz, cudaStream_t s) {
  const int64_t plane = (int64_t)nx * ny;
  if (w && qxy) {
    int chunks = (int)((plane + 65535) / 65536);
    if (chunks < 1) chunks = 1;
    flux_xy_kernel<Tf><<<(unsigned)(nz * chunks), 256, 0, s>>>((const Tf*)w, plane, chunks, qxy);
    count_launches(1);
  }
  if (v && qxz) {
    const int64_t rows = (int64_t)ny * nz;
    flux_xz_kernel<Tf><<<(unsigned)((rows + 7) / 8), 256, 0, s>>>((const Tf*)v, nx, ny, nz, qxz);
    count_launches(1);
  }
  if (u && qyz) {
    dim3 grid((unsigned)((nx + 255) / 256), (unsigned)nz);
    flux_yz_kernel<Tf><<<grid, 256, 0, s>>>((const Tf*)u, nx, ny, qyz);
    count_launches(1);
  }
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

extern "C" int ptv_flux_profiles(const void* d_u, const void* d_v, const void* d_w, int nx, int ny, int nz,
                                 int dtype, double* d_qxy, double* d_qxz, double* d_qyz, void* stream) {
  if (nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_flux_profiles: empty grid"); return PTV_ERR_INVALID; }
  if (nz > 65535) { set_error("ptv_flux_profiles: nz > 65535 planes per call"); return PTV_ERR_INVALID; }
  if (dtype == PTV_F32) return flux_launch<float>(d_u, d_v, d_w, nx, ny, nz, d_qxy, d_qxz, d_qyz, (cudaStream_t)stream);
  if (dtype == PTV_F64) return flux_launch<double>(d_u, d_v, d_w, nx, ny, nz, d_qxy, d_qxz, d_qyz, (cudaStream_t)stream);
  set_error("ptv_flux_profiles: bad dtype");
  return PTV_ERR_INVALID;
}

// z-slab form (one rank's planes of a sharded grid): the planes just below / above the slab come from the
// z-neighbours as (3, ny, nx) buffers (u, v, w) so that the slab's first and last planes get np.gradient's central
// differences; NULL = that side is a face of the whole domain.  Served by the bulk-tile kernel only; shapes it
// cannot take return PTV_ERR_INVALID and the caller pads the slab with its halo planes instead (engine.py does).
extern "C" int ptv_strain_vorticity_slab(const void* d_u, const void* d_v, const void* d_w, const uint8_t* d_mask, int nx,
                                         int ny, int nz, double dx, double dy, double dz, const void* d_below,
                                         const void* d_above, int dtype, void* d_strain, void* d_vorticity, void* stream) {
  if (!d_u || !d_v || !d_w) { set_error("ptv_strain_vorticity_slab: NULL argument"); return PTV_ERR_INVALID; }
  if (!d_strain && !d_vorticity) { set_error("ptv_strain_vorticity_slab: nothing to compute"); return PTV_ERR_INVALID; }
  const int nz_ext = nz + (d_below ? 1 : 0) + (d_above ? 1 : 0);
  if (nx < 2 || ny < 2 || nz < 1 || nz_ext < 2) {
    set_error("Shape of array too small to calculate a numerical gradient, at least (edge_order + 1) elements are required.");
    return PTV_ERR_INVALID;
  }
  if (dtype == PTV_F32) {
    const int rc = launch_strain_vorticity_bulk((const float*)d_u, (const float*)d_v, (const float*)d_w, d_mask, nx, ny, nz, dx,
                                                dy, dz, (float*)d_strain, (float*)d_vorticity, (const float*)d_below,
                                                (const float*)d_above, (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  set_error("ptv_strain_vorticity_slab: needs float32 fields with nx % 16 == 0 and 16-byte aligned buffers");
  return PTV_ERR_INVALID;
}
