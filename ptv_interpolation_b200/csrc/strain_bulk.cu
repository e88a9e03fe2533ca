// Shear-rate magnitude and vorticity magnitude (velocity_analysis.py:10-63, 94-120) for float32 fields, as a
// z-MARCHING kernel over shared-memory plane tiles filled by 1-D bulk copies (cp.async.bulk, the TMA engine).
//
// The nine np.gradient stencils of a voxel read 6 neighbours of each of u, v, w.  A CTA owns a column of 8 rows x
// 128 x-positions and walks up in z: a warp owns one row, a lane four consecutive x.  The z-neighbours of a voxel
// are the thread's own values of the planes before and after (registers), its x / y neighbours come from the
// current plane's tile in shared memory ((8 + 2) rows x (128 + 8) columns per field, halo included), so a field
// value crosses L2 -> SM 1.33 times instead of five (y-marching: three; one load per neighbour: seven).  The tiles
// travel through a ring of six stages three planes ahead of their use; a stage is read at two consecutive
// steps (first the thread's own values, as the "plane after", then the neighbours, as the centre plane).
// Arithmetic, operation order and divisions are those of the direct-load kernels in grid_ops.cu (bit-identical
// float64 results, rounded to float32 once).
#include <math.h>

#include "bulk_pipe.cuh"
#include "ptv_internal.cuh"

namespace ptv {

// num / den with the IEEE result (see grid_ops.cu): power-of-two divisors multiply by the exact reciprocal
struct SbDivisor {
  double den, inv;
};
struct SbDivisors6 { SbDivisor d[6]; };  // x edge, x interior (2 h), y edge, y interior, z edge, z interior

template <bool kPow2>
__device__ __forceinline__ double sb_div(double num, const SbDivisor& d) {
  if (kPow2) return __dmul_rn(num, d.inv);
  if (num == 0.0) return d.den > 0.0 ? num : -num;  // signed zero without the division slow path
  return __ddiv_rn(num, d.den);
}

static constexpr int kSbRows = 8;            // tile rows = warps
static constexpr int kSbCols = 128;          // tile columns = 32 lanes x 4
static constexpr int kSbThreads = kSbRows * 32;
static constexpr int kSbFRow = kSbCols + 8;  // floats per staged row (4-column halo on both sides)
static constexpr int kSbFRows = kSbRows + 2; // staged rows per field (one halo row on both sides)
static constexpr int kSbStages = 6;         // 2 CTAs x 6 x 17 KB; a stage is held for two steps
static constexpr int kSbFieldBytes = kSbFRows * kSbFRow * 4;                         // 5440
static constexpr int kSbStageBytes = ((3 * kSbFieldBytes + kSbRows * kSbCols) + 127) / 128 * 128;  // + mask tile

template <bool kPow2>
__global__ void __launch_bounds__(kSbThreads, 2) strain_vorticity_bulk_kernel(
    const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
    const uint8_t* __restrict__ mask, int nx, int ny, int nz, const SbDivisors6 dv, float* __restrict__ strain,
    float* __restrict__ vort, int tiles_x, int tiles_y, int zseg, int la) {
  extern __shared__ __align__(128) unsigned char sb_smem[];
  __shared__ __align__(8) uint64_t full_bar[kSbStages], empty_bar[kSbStages];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int b = blockIdx.x;
  const int tx = b % tiles_x; b /= tiles_x;
  const int ty = b % tiles_y; b /= tiles_y;
  const int zs = b * zseg, ze = min(nz, zs + zseg);  // output planes [zs, ze)
  const int x0 = tx * kSbCols, y0 = ty * kSbRows;
  const int cw = min(kSbCols, nx - x0);              // live columns of the tile (multiple of 16)
  const int left = x0 > 0 ? 4 : 0, right = x0 + cw < nx ? 4 : 0;
  const int ra = y0 > 0 ? 0 : 1;                     // first / last staged row that exists
  const int rb = min(kSbFRows - 1, ny - y0);         // (tile row r holds y = y0 - 1 + r)
  const int64_t plane = (int64_t)nx * ny;
  const int nsteps = ze - zs + 2;                    // planes zs - 1 .. ze

  if (t == 0) {
#pragma unroll
    for (int s = 0; s < kSbStages; ++s) {
      mbar_init(smem_addr(&full_bar[s]), 1);
      mbar_init(smem_addr(&empty_bar[s]), kSbRows);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---------------- producer duty (one warp per step, in turn): every lane issues its own row copies
  const uint32_t frow_bytes = (uint32_t)(cw + left + right) * 4u;
  const uint32_t nfrows = (uint32_t)(rb - ra + 1);
  const uint32_t nmrows = mask != nullptr ? (uint32_t)min(kSbRows, ny - y0) : 0u;
  auto produce = [&](int k) {  // warp-collective
    const int s = k % kSbStages;
    const uint32_t bar = smem_addr(&full_bar[s]);
    if (k >= kSbStages) mbar_wait(smem_addr(&empty_bar[s]), ((k / kSbStages) & 1) ^ 1);
    const int zp = zs - 1 + k;
    if (zp < 0 || zp >= nz) {  // no such plane: the step still completes its barrier phase
      if (lane == 0) mbar_arrive(bar);
      return;
    }
    if (lane == 0) mbar_expect_tx(bar, 3u * nfrows * frow_bytes + nmrows * (uint32_t)cw);
    __syncwarp();
    const uint32_t sb = smem_addr(sb_smem + (size_t)s * kSbStageBytes);
    for (int i = lane; i < 3 * kSbFRows + kSbRows; i += 32) {
      if (i < 3 * kSbFRows) {
        const int f = i / kSbFRows, r = i - f * kSbFRows;
        if (r >= ra && r <= rb) {
          const float* src = (f == 0 ? u : f == 1 ? v : w) + (int64_t)zp * plane + (int64_t)(y0 - 1 + r) * nx + (x0 - left);
          bulk_g2s(sb + (uint32_t)(f * kSbFieldBytes + (r * kSbFRow + 4 - left) * 4), src, frow_bytes, bar);
        }
      } else {
        const int r = i - 3 * kSbFRows;
        if ((uint32_t)r < nmrows)
          bulk_g2s(sb + (uint32_t)(3 * kSbFieldBytes + r * kSbCols), mask + (int64_t)zp * plane + (int64_t)(y0 + r) * nx + x0,
                   (uint32_t)cw, bar);
      }
    }
  };
  for (int k = 0; k < la && k < nsteps; ++k)
    if (warp == (k & (kSbRows - 1))) produce(k);

  // ---------------- consumers
  const int y = y0 + warp, x = x0 + 4 * lane;
  const bool in = y < ny && x < nx;
  const bool x_first = x == 0, x_last = x + 4 >= nx, y_first = y == 0, y_last = y == ny - 1;
  const int own = ((warp + 1) * kSbFRow + 4 + 4 * lane) * 4;  // byte offset of the thread's four values in a field tile
  float prv[3][4], cen[3][4], nxt[3][4];
#pragma unroll
  for (int f = 0; f < 3; ++f)
#pragma unroll
    for (int j = 0; j < 4; ++j) prv[f][j] = cen[f][j] = nxt[f][j] = 0.0f;
  const SbDivisor dxe = dv.d[0], dxi = dv.d[1], dye = dv.d[2], dyi = dv.d[3], dze = dv.d[4], dzi = dv.d[5];
  const SbDivisor dyy = (y_first || y_last) ? dye : dyi;

#pragma unroll 1
  for (int k = 0; k < nsteps; ++k) {
    {
      const int kp = k + la;
      if (kp < nsteps && warp == (kp & (kSbRows - 1))) produce(kp);
    }
    const int s = k % kSbStages;
    const unsigned char* sb = sb_smem + (size_t)s * kSbStageBytes;
    mbar_wait(smem_addr(&full_bar[s]), (k / kSbStages) & 1);
    const int zp = zs - 1 + k;  // the plane that has just arrived: the "plane after" of the centre plane zp - 1
    if (in && zp >= 0 && zp < nz) {
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const float4 q = *reinterpret_cast<const float4*>(sb + f * kSbFieldBytes + own);
        nxt[f][0] = q.x; nxt[f][1] = q.y; nxt[f][2] = q.z; nxt[f][3] = q.w;
      }
    }
    if (k >= 1) {
      const int sc = (k - 1) % kSbStages;  // the centre plane's tile: neighbours in x and y, mask
      const unsigned char* sbc = sb_smem + (size_t)sc * kSbStageBytes;
      const int zc = zp - 1;
      const bool work = k >= 2 && in;
      uint32_t m4 = 0x01010101u;
      float up[3][4], dn[3][4], xl[3], xr[3];
      if (work) {
        if (mask != nullptr) m4 = *reinterpret_cast<const uint32_t*>(sbc + 3 * kSbFieldBytes + warp * kSbCols + 4 * lane);
        if (m4 != 0u) {
#pragma unroll
          for (int f = 0; f < 3; ++f) {
            const unsigned char* ft = sbc + f * kSbFieldBytes;
            const float4 a = *reinterpret_cast<const float4*>(ft + own - kSbFRow * 4);
            const float4 c = *reinterpret_cast<const float4*>(ft + own + kSbFRow * 4);
            up[f][0] = a.x; up[f][1] = a.y; up[f][2] = a.z; up[f][3] = a.w;
            dn[f][0] = c.x; dn[f][1] = c.y; dn[f][2] = c.z; dn[f][3] = c.w;
            xl[f] = *reinterpret_cast<const float*>(ft + own - 4);
            xr[f] = *reinterpret_cast<const float*>(ft + own + 16);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&empty_bar[sc]));  // the centre tile's values are in registers
      if (work) {
        float so[4] = {0.0f, 0.0f, 0.0f, 0.0f}, vo[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (m4 != 0u) {
          const bool z_first = zc == 0, z_last = zc == nz - 1;
          const SbDivisor dzz = (z_first || z_last) ? dze : dzi;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if ((m4 & (0xffu << (8 * j))) == 0u) continue;  // solid voxel: 0 (velocity_analysis.py:58-61,116-117)
            const bool xe = (j == 0 && x_first) || (j == 3 && x_last);
            const SbDivisor dxx = xe ? dxe : dxi;
            double gx[3], gy[3], gz[3];
#pragma unroll
            for (int f = 0; f < 3; ++f) {
              const double c = (double)cen[f][j];
              // np.gradient: central difference over 2 h inside, one-sided over h at the two ends of an axis
              const double xa = j == 0 ? (x_first ? c : (double)xl[f]) : (double)cen[f][j > 0 ? j - 1 : 0];
              const double xb = j == 3 ? (x_last ? c : (double)xr[f]) : (double)cen[f][j < 3 ? j + 1 : 3];
              gx[f] = sb_div<kPow2>(__dsub_rn(xb, xa), dxx);
              const double ya = y_first ? c : (double)up[f][j], yb = y_last ? c : (double)dn[f][j];
              gy[f] = sb_div<kPow2>(__dsub_rn(yb, ya), dyy);
              const double za = z_first ? c : (double)prv[f][j], zb = z_last ? c : (double)nxt[f][j];
              gz[f] = sb_div<kPow2>(__dsub_rn(zb, za), dzz);
            }
            if (strain != nullptr) {
              const double exx = __dmul_rn(2.0, gx[0]), eyy = __dmul_rn(2.0, gy[1]), ezz = __dmul_rn(2.0, gz[2]);
              const double exy = __dadd_rn(gy[0], gx[1]), exz = __dadd_rn(gz[0], gx[2]), eyz = __dadd_rn(gz[1], gy[2]);
              const double diag = __dmul_rn(0.5, __dadd_rn(__dadd_rn(__dmul_rn(exx, exx), __dmul_rn(eyy, eyy)), __dmul_rn(ezz, ezz)));
              so[j] = (float)sqrt(__dadd_rn(__dadd_rn(__dadd_rn(diag, __dmul_rn(exy, exy)), __dmul_rn(exz, exz)), __dmul_rn(eyz, eyz)));
            }
            if (vort != nullptr) {
              const double vx = __dsub_rn(gy[2], gz[1]), vy = __dsub_rn(gz[0], gx[2]), vz = __dsub_rn(gx[1], gy[0]);
              vo[j] = (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
            }
          }
        }
        const int64_t o = (int64_t)zc * plane + (int64_t)y * nx + x;
        if (strain != nullptr) *reinterpret_cast<float4*>(strain + o) = make_float4(so[0], so[1], so[2], so[3]);
        if (vort != nullptr) *reinterpret_cast<float4*>(vort + o) = make_float4(vo[0], vo[1], vo[2], vo[3]);
      }
    }
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
      for (int j = 0; j < 4; ++j) { prv[f][j] = cen[f][j]; cen[f][j] = nxt[f][j]; }
  }
}

static SbDivisor sb_divisor(double den, bool* pow2) {
  SbDivisor d;
  d.den = den;
  d.inv = 1.0 / den;
  int e;
  const double m = frexp(fabs(den), &e);
  *pow2 = *pow2 && (m == 0.5) && e > -1000 && e < 1000;
  return d;
}

// Returns PTV_OK after launching, or -1 if shape / alignment rule the bulk path out (the caller falls back).
int launch_strain_vorticity_bulk(const float* u, const float* v, const float* w, const uint8_t* mask, int nx, int ny,
                                 int nz, double dx, double dy, double dz, float* strain, float* vort, cudaStream_t s) {
  const auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (tuning().stencil_bulk == 0 || nx % 16 != 0 || !al16(u) || !al16(v) || !al16(w) || !al16(mask) || !al16(strain) ||
      !al16(vort))
    return -1;
  bool pow2 = true;
  SbDivisors6 dv;
  dv.d[0] = sb_divisor(dx, &pow2); dv.d[1] = sb_divisor(2.0 * dx, &pow2);
  dv.d[2] = sb_divisor(dy, &pow2); dv.d[3] = sb_divisor(2.0 * dy, &pow2);
  dv.d[4] = sb_divisor(dz, &pow2); dv.d[5] = sb_divisor(2.0 * dz, &pow2);
  const int tiles_x = (nx + kSbCols - 1) / kSbCols, tiles_y = (ny + kSbRows - 1) / kSbRows;
  // z is cut so that the grid holds several waves of 148 SMs x 2 CTAs (each cut re-reads two planes of its tile)
  int zsplit = (int)((148 * 2 * 8 + (int64_t)tiles_x * tiles_y - 1) / ((int64_t)tiles_x * tiles_y));
  zsplit = max(1, min(zsplit, nz / 32 > 0 ? nz / 32 : 1));
  const int zseg = (nz + zsplit - 1) / zsplit;
  zsplit = (nz + zseg - 1) / zseg;
  const int64_t grid = (int64_t)tiles_x * tiles_y * zsplit;
  if (grid > 2147483647LL) return -1;
  const size_t smem = (size_t)kSbStages * kSbStageBytes;
  // planes requested ahead: a tile is released one step after its own step, and two more steps of slack keep the
  // warp on producer duty from waiting for the slowest warp (stencil_fused.cu)
  const int la = max(1, min(kSbStages - 2, tuning().stencil_la > 0 ? tuning().stencil_la : kSbStages - 3));
#define PTV_SB_LAUNCH(P2)                                                                                   \
  do {                                                                                                      \
    auto kern = strain_vorticity_bulk_kernel<P2>;                                                           \
    PTV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    kern<<<(unsigned)grid, kSbThreads, smem, s>>>(u, v, w, mask, nx, ny, nz, dv, strain, vort, tiles_x,     \
                                                  tiles_y, zseg, la);                                       \
  } while (0)
  if (pow2) PTV_SB_LAUNCH(true); else PTV_SB_LAUNCH(false);
#undef PTV_SB_LAUNCH
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

}  // namespace ptv
