// Shear-rate magnitude and vorticity magnitude (velocity_analysis.py:10-63, 94-120) for float32 fields, as a
// z-MARCHING kernel over shared-memory plane tiles filled by 1-D bulk copies (cp.async.bulk, the TMA engine).
//
// The nine np.gradient stencils of a voxel read 6 neighbours of each of u, v, w.  A CTA owns a column of 8 rows x
// 128 x-positions and walks up in z; a warp owns one row.  Three consecutive plane tiles ((8 + 2) rows x (128 + 8)
// columns per field, halo included) are resident in shared memory: the centre plane and its z-neighbours, so a
// field value crosses L2 -> SM 1.33 times instead of seven (one load per neighbour) and every neighbour of a voxel is
// one shared-memory read.  The tiles travel through a ring of six stages two planes ahead of their use.
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

// kPow2 = 1: all six divisors are powers of two; 0: reciprocal + two exact FMA corrections (bulk_pipe.cuh, the
// divisors are inside its range); 2: IEEE division
template <int kPow2>
__device__ __forceinline__ double sb_div(double num, const SbDivisor& d) {
  if (kPow2 == 1) return __dmul_rn(num, d.inv);
  if (kPow2 == 0) return div_by_spacing(num, d.den, d.inv);
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
static constexpr int kSbWarpBytes = 2 * kSbCols * 4 + kSbCols;  // per warp: two result rows and the fluid-voxel list

template <int kPow2>
__global__ void __launch_bounds__(kSbThreads, 2) strain_vorticity_bulk_kernel(
    const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
    const uint8_t* __restrict__ mask, int nx, int ny, int nz, const SbDivisors6 dv, float* __restrict__ strain,
    float* __restrict__ vort, const float* __restrict__ below, const float* __restrict__ above, int tiles_x,
    int tiles_y, int zseg, int la) {
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
    // the planes just outside the slab come from the z-neighbours' halo buffers ((3, ny, nx): u, v, w), if any
    const float* halo = zp < 0 ? below : (zp >= nz ? above : nullptr);
    const bool inside = zp >= 0 && zp < nz;
    if (!inside && halo == nullptr) {  // no such plane: the step still completes its barrier phase
      if (lane == 0) mbar_arrive(bar);
      return;
    }
    if (lane == 0) mbar_expect_tx(bar, 3u * nfrows * frow_bytes + (inside ? nmrows * (uint32_t)cw : 0u));
    __syncwarp();
    const uint32_t sb = smem_addr(sb_smem + (size_t)s * kSbStageBytes);
    for (int i = lane; i < 3 * kSbFRows + kSbRows; i += 32) {
      if (i < 3 * kSbFRows) {
        const int f = i / kSbFRows, r = i - f * kSbFRows;
        if (r >= ra && r <= rb) {
          const float* pl = inside ? (f == 0 ? u : f == 1 ? v : w) + (int64_t)zp * plane : halo + (int64_t)f * plane;
          const float* src = pl + (int64_t)(y0 - 1 + r) * nx + (x0 - left);
          bulk_g2s(sb + (uint32_t)(f * kSbFieldBytes + (r * kSbFRow + 4 - left) * 4), src, frow_bytes, bar);
        }
      } else if (inside) {
        const int r = i - 3 * kSbFRows;
        if ((uint32_t)r < nmrows)
          bulk_g2s(sb + (uint32_t)(3 * kSbFieldBytes + r * kSbCols), mask + (int64_t)zp * plane + (int64_t)(y0 + r) * nx + x0,
                   (uint32_t)cw, bar);
      }
    }
  };
  for (int k = 0; k < la && k < nsteps; ++k)
    if (warp == (k & (kSbRows - 1))) produce(k);

  // ---------------- consumers.  Step k: plane zs - 1 + k has arrived; the centre plane is the one before it, its
  // z-neighbours sit in the stages before and after.  Only FLUID voxels cost arithmetic (velocity_analysis.py:58-61,
  // 116-117 zero the solid ones), and in a porous medium a warp's 128 voxels are a patchwork of both -- so the warp
  // compacts the fluid voxels of its row into a list and lane i takes list entries i, i + 32, ...: the ~130
  // float64 instructions per voxel run with (nearly) all lanes busy instead of the 40 % a fixed voxel-to-lane
  // mapping gives at porosity 0.4.  Every neighbour is one scalar read of a staged tile (domain edges: the
  // neighbour index is clamped to the voxel itself, np.gradient's one-sided difference), results go through a
  // per-warp row buffer so the global stores stay 16-byte vectors.
  const int y = y0 + warp, x = x0 + 4 * lane;
  const bool in = y < ny && x < nx;
  const bool y_first = y == 0, y_last = y == ny - 1;
  unsigned char* wsm = sb_smem + (size_t)kSbStages * kSbStageBytes + (size_t)warp * kSbWarpBytes;
  float* outs = reinterpret_cast<float*>(wsm);            // [128] shear rate of the row
  float* outv = outs + kSbCols;                           // [128] vorticity
  uint8_t* list = reinterpret_cast<uint8_t*>(outv + kSbCols);  // [128] x-offsets of the row's fluid voxels
  const SbDivisor dxe = dv.d[0], dxi = dv.d[1], dze = dv.d[4], dzi = dv.d[5];
  const SbDivisor dyy = (y_first || y_last) ? dv.d[2] : dv.d[3];
  const int rowc = (warp + 1) * kSbFRow, rowu = (y_first ? warp + 1 : warp) * kSbFRow, rowd = (y_last ? warp + 1 : warp + 2) * kSbFRow;
  const unsigned lt = (1u << lane) - 1u;

#pragma unroll 1
  for (int k = 0; k < nsteps; ++k) {
    {
      const int kp = k + la;
      if (kp < nsteps && warp == (kp & (kSbRows - 1))) produce(kp);
    }
    mbar_wait(smem_addr(&full_bar[k % kSbStages]), (k / kSbStages) & 1);
    if (k < 2) continue;
    const int zc = zs + k - 2;  // centre plane
    const int sa = (k - 2) % kSbStages, sc = (k - 1) % kSbStages, sn = k % kSbStages;
    const float* tc = reinterpret_cast<const float*>(sb_smem + (size_t)sc * kSbStageBytes);
    const bool z_first = zc == 0 && below == nullptr, z_last = zc == nz - 1 && above == nullptr;  // true domain faces
    const float* tlo = z_first ? tc : reinterpret_cast<const float*>(sb_smem + (size_t)sa * kSbStageBytes);
    const float* thi = z_last ? tc : reinterpret_cast<const float*>(sb_smem + (size_t)sn * kSbStageBytes);
    const SbDivisor dzz = (z_first || z_last) ? dze : dzi;
    // ---- fluid voxels of the row, in x order
    uint32_t m4 = 0u;
    if (in) m4 = mask != nullptr ? *reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(tc) + 3 * kSbFieldBytes + warp * kSbCols + 4 * lane)
                                 : 0x01010101u;
    int total = 0, rank = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool fl = (m4 & (0xffu << (8 * j))) != 0u;
      const unsigned bal = __ballot_sync(0xffffffffu, fl);
      // voxels are ordered lane-major (x = 4 lane + j): everything in lower lanes, then this lane's earlier j
      total += __popc(bal);
      rank += __popc(bal & lt);
    }
    {
      int r = rank;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((m4 & (0xffu << (8 * j))) != 0u) list[r++] = (uint8_t)(4 * lane + j);
    }
    *reinterpret_cast<float4*>(outs + 4 * lane) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    *reinterpret_cast<float4*>(outv + 4 * lane) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < total; i += 32) {
      const int xo = list[i];
      const int xg = x0 + xo, cx = 4 + xo;
      const bool xe0 = xg == 0, xe1 = xg == nx - 1;
      const int cl = xe0 ? cx : cx - 1, cr = xe1 ? cx : cx + 1;
      const SbDivisor dxx = (xe0 || xe1) ? dxe : dxi;
      double gx[3], gy[3], gz[3];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const int fo = f * (kSbFieldBytes / 4);
        gx[f] = sb_div<kPow2>(__dsub_rn((double)tc[fo + rowc + cr], (double)tc[fo + rowc + cl]), dxx);
        gy[f] = sb_div<kPow2>(__dsub_rn((double)tc[fo + rowd + cx], (double)tc[fo + rowu + cx]), dyy);
        gz[f] = sb_div<kPow2>(__dsub_rn((double)thi[fo + rowc + cx], (double)tlo[fo + rowc + cx]), dzz);
      }
      if (strain != nullptr) {
        const double exx = __dmul_rn(2.0, gx[0]), eyy = __dmul_rn(2.0, gy[1]), ezz = __dmul_rn(2.0, gz[2]);
        const double exy = __dadd_rn(gy[0], gx[1]), exz = __dadd_rn(gz[0], gx[2]), eyz = __dadd_rn(gz[1], gy[2]);
        const double diag = __dmul_rn(0.5, __dadd_rn(__dadd_rn(__dmul_rn(exx, exx), __dmul_rn(eyy, eyy)), __dmul_rn(ezz, ezz)));
        outs[xo] = (float)sqrt(__dadd_rn(__dadd_rn(__dadd_rn(diag, __dmul_rn(exy, exy)), __dmul_rn(exz, exz)), __dmul_rn(eyz, eyz)));
      }
      if (vort != nullptr) {
        const double vx = __dsub_rn(gy[2], gz[1]), vy = __dsub_rn(gz[0], gx[2]), vz = __dsub_rn(gx[1], gy[0]);
        outv[xo] = (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_addr(&empty_bar[sa]));  // the plane below the centre is done with
    if (in) {
      const int64_t o = (int64_t)zc * plane + (int64_t)y * nx + x;
      if (strain != nullptr) *reinterpret_cast<float4*>(strain + o) = *reinterpret_cast<const float4*>(outs + 4 * lane);
      if (vort != nullptr) *reinterpret_cast<float4*>(vort + o) = *reinterpret_cast<const float4*>(outv + 4 * lane);
    }
    __syncwarp();  // the row buffers are rewritten in the next step
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
                                 int nz, double dx, double dy, double dz, float* strain, float* vort, const float* below,
                                 const float* above, cudaStream_t s) {
  const auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (tuning().stencil_bulk == 0 || nx % 16 != 0 || !al16(u) || !al16(v) || !al16(w) || !al16(mask) || !al16(strain) ||
      !al16(vort) || !al16(below) || !al16(above))
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
  const size_t smem = (size_t)kSbStages * kSbStageBytes + (size_t)kSbRows * kSbWarpBytes;
  // planes requested ahead: a tile is held for three steps (plane above, centre, plane below); one more step of
  // slack keeps the warp on producer duty from waiting for the slowest warp (stencil_fused.cu)
  const int la = max(1, min(kSbStages - 3, tuning().stencil_la > 0 ? tuning().stencil_la : kSbStages - 4));
#define PTV_SB_LAUNCH(P2)                                                                                   \
  do {                                                                                                      \
    auto kern = strain_vorticity_bulk_kernel<P2>;                                                           \
    PTV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    kern<<<(unsigned)grid, kSbThreads, smem, s>>>(u, v, w, mask, nx, ny, nz, dv, strain, vort, below, above, \
                                                  tiles_x, tiles_y, zseg, la);                              \
  } while (0)
  bool fast = true;
  for (int c = 0; c < 6; ++c) fast = fast && spacing_ok(dv.d[c].den);
  if (pow2) PTV_SB_LAUNCH(1); else if (fast) PTV_SB_LAUNCH(0); else PTV_SB_LAUNCH(2);
#undef PTV_SB_LAUNCH
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

}  // namespace ptv
