// Internal declarations shared by the .cu translation units of libptvb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "ptv_b200.h"

namespace ptv {

// One cell-sorted particle record: 32 B, 16-B aligned so a row of cells is one contiguous,
// bulk-copyable range.  `idx` is the particle's ORIGINAL row (ties are broken on it).
struct __align__(16) ParticleRec {
  double x, y, z;
  int32_t idx;
  int32_t pad;
};
static_assert(sizeof(ParticleRec) == 32, "ParticleRec must be 32 bytes");

struct __align__(16) Value4 {  // (u,v,w,0) in ORIGINAL particle order
  double u, v, w, pad;
};

struct HashGrid {  // plain-old-data view passed to kernels by value
  const ParticleRec* rec;
  const int32_t* cell_start;  // [ncells+1], cell id = (cz*cny + cy)*cnx + cx
  const Value4* vals;         // (u,v,w,0) float64 in ORIGINAL particle order
  const Value4* vals_s64;     // the same in cell-sorted order (parallel to rec)
  const float4* vals_s32;     // float32 copy in cell-sorted order
  const double* pts;          // original (n,3) float64 rows (caller-owned; valid during the call that set it)
  double ox, oy, oz;          // origin = particle bbox minimum
  double cell, inv_cell;
  int cnx, cny, cnz;
  int64_t n;
  // slab hash (ptv_hash_build_slab): only particles with clip_lo <= z <= clip_hi are binned; a voxel's
  // neighbours are certified only if its k-th distance stays inside that range (+-INFINITY = side not clipped)
  double clip_lo, clip_hi;
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

struct Tuning {
  double ppc = 0.5;  // target particles per cell (finer cells -> tighter scan regions)
  int r0 = 1;        // rings merged into the first staging batch
  int tile = 128;    // threads (= voxels) per tile, heap kernel
  int stream = 2;    // idw/sibson with k >= 8: 2 = warp-private streaming kernel (knn_duo.cu), 1 = the CTA-wide one
                     // (knn_stream.cu), 0 = heap kernel only; the heap kernel is always the exact fallback
  int stream_tile = 128;
  int stats = 0;     // 1 = count streamed tiles
  int linear_k = 64;   // method='linear': the first candidate radius is the one expected to hold this many particles (2.5 spacings)
  int linear_occ = 3;  // method='linear': CTAs per SM the kernel is compiled for (3: 168 registers, 4: 128 + spills)
  int hull = 2;      // method='linear': hull membership decided on the hull-candidate list (2: cell + particle dominance, 1: cell dominance only), 0 = scan all particles
  double rscale = 1.3;  // stream kernel: first scan radius^2 = rscale * r_est^2
  int rbf_regs = 1;      // local RBF with k + tail <= 32: 1 = register-resident elimination, 0 = shared-memory matrix
  int stencil_la = 0;    // bulk stencil kernel: rows requested ahead (0 = stages - 2)
  int stencil_bulk = 1;  // stencil kernels: 1 = bulk-async (TMA) row pipelines where shape and alignment allow, 0 = direct loads
};
Tuning& tuning();
void count_launches(int n);

#define PTV_CUDA(expr)                                                    \
  do {                                                                    \
    cudaError_t _e = (expr);                                              \
    if (_e != cudaSuccess) return ::ptv::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

}  // namespace ptv

struct ptv_hash {
  // device buffers (grown on demand, reused across builds)
  ptv::ParticleRec* rec = nullptr;
  ptv::Value4* vals = nullptr;
  ptv::Value4* vals_s64 = nullptr;
  float4* vals_s32 = nullptr;
  int32_t* cid = nullptr;
  int32_t* sorted_idx = nullptr;
  int32_t* slot = nullptr;        // arrival order of each particle inside its cell (count pass)
  int32_t* cell_start = nullptr;  // ncells+1
  int32_t* cell_fill = nullptr;   // ncells
  int32_t* scan_tmp = nullptr;
  double* bbox_dev = nullptr;     // 6 doubles + scratch partials
  double* bbox_host = nullptr;    // pinned, 8 doubles
  int* err_flag = nullptr;        // device, RBF singularity flag
  int* err_host = nullptr;        // pinned
  int* fail_list = nullptr;       // tiles handed from the streaming to the heap kernel
  int64_t fail_cap = 0;
  unsigned* fail_flags = nullptr;  // one bit per heap tile (de-duplicates the fail list)
  int64_t fail_flags_cap = 0;
  unsigned long long* fail_count = nullptr;  // [0] low 32 bits: fail count; [1]: streamed tiles (stats)
  // method='linear': hull-candidate records and the dominance tables they come from (built on demand)
  ptv::ParticleRec* hull_rec = nullptr;   // stage 1: particles of undominated cells
  ptv::ParticleRec* hull_rec2 = nullptr;  // stage 2: particles with an empty closed octant
  ptv::ParticleRec* hull_list = nullptr;  // the list the kernel uses (one of the two)
  uint8_t* hull_keep = nullptr;
  int64_t hull_cap2 = 0;
  double* hull_box = nullptr;
  int64_t hull_cap = 0;
  int* hull_tab = nullptr;
  int hull_cap_rows = 0;
  int hull_n = 0;
  bool hull_valid = false;
  bool last_used_stream = false;
  double clip_lo = -1.0 / 0.0, clip_hi = 1.0 / 0.0;  // z-range of the binned particles (slab hash)
  int* clip_count = nullptr;                          // device: voxels whose search left the range
  int64_t last_stage_counts[2] = {0, 0};
  int64_t cap_n = 0;
  int64_t cap_cells = 0;
  int64_t cap_scan = 0;
  // geometry of the last build
  int64_t n = 0;
  int dims[3] = {0, 0, 0};
  double origin[3] = {0, 0, 0};
  double cell = 0;
  int max_cell_count = 0;
  const double* pts = nullptr;
  bool built = false;
  ptv::HashGrid view() const;
};
