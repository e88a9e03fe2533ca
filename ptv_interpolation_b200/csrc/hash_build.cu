// Uniform-grid spatial hash of the particle cloud (replaces the cKDTree build at
// interpolator.py:90,132).  One-pass radix (counting) sort on the cell id:
//   bbox reduce (one launch, last block folds) -> cell id + histogram + slot-in-cell + original-order values
//   -> exclusive prefix scan -> scatter (no atomics: start + slot) -> per-cell canonical order (by original
//   index) -> 32-byte record fill.  Seven launches.
// All kernels are HBM/L2-bound streaming passes over the particle arrays; algorithmic traffic
// is ~112 B/particle.
#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <cmath>

#include "ptv_internal.cuh"

namespace ptv {

static constexpr int kBboxBlocks = 296;  // 2 x 148 SMs
static constexpr int kBboxThreads = 256;
static constexpr int kScanThreads = 256;
static constexpr int kScanItems = 8;
static constexpr int kScanTile = kScanThreads * kScanItems;

// ------------------------------------------------------------------------------ bbox
// partial layout per block: minx,miny,minz,maxx,maxy,maxz,bad
__global__ void __launch_bounds__(kBboxThreads) bbox_partial_kernel(const double* __restrict__ pts,
                                                                     int64_t n,
                                                                     double* __restrict__ partial,
                                                                     unsigned* __restrict__ done_count,
                                                                     double* __restrict__ out) {
  double mn[3] = {INFINITY, INFINITY, INFINITY};
  double mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  double bad = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double v = pts[i * 3 + c];
      if (!isfinite(v)) bad = 1.0;
      mn[c] = fmin(mn[c], v);
      mx[c] = fmax(mx[c], v);
    }
  }
  __shared__ double sh[7][kBboxThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mn[c] = fmin(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    bad = fmax(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  }
  if (lane == 0) {
    for (int c = 0; c < 3; ++c) {
      sh[c][wid] = mn[c];
      sh[3 + c][wid] = mx[c];
    }
    sh[6][wid] = bad;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    const int q = threadIdx.x;
    double acc = sh[q][0];
    for (int w2 = 1; w2 < kBboxThreads / 32; ++w2)
      acc = (q < 3) ? fmin(acc, sh[q][w2]) : fmax(acc, sh[q][w2]);
    partial[blockIdx.x * 7 + q] = acc;
  }
  // the last block to arrive folds the per-block partials (no second launch)
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(done_count, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last && threadIdx.x < 7) {
    const int q = threadIdx.x;
    const volatile double* vp = partial;
    double acc = vp[q];
    for (int b = 1; b < (int)gridDim.x; ++b) {
      const double v = vp[b * 7 + q];
      acc = (q < 3) ? fmin(acc, v) : fmax(acc, v);
    }
    out[q] = acc;
    if (q == 0) *done_count = 0u;  // ready for the next build
  }
}

// ------------------------------------------------------------------------------ binning
__device__ __forceinline__ int cell_coord(double p, double o, double inv_cell, int n) {
  int c = (int)floor((p - o) * inv_cell);
  return min(max(c, 0), n - 1);
}

__global__ void __launch_bounds__(256) cell_count_kernel(const double* __restrict__ pts, int64_t n,
                                                          double ox, double oy, double oz,
                                                          double inv_cell, int cnx, int cny, int cnz,
                                                          double zlo, double zhi,
                                                          const double* __restrict__ vals_in,
                                                          Value4* __restrict__ vals,
                                                          int32_t* __restrict__ cid,
                                                          int32_t* __restrict__ slot,
                                                          int32_t* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  {  // (u, v, w, 0) in ORIGINAL order for the gathers by particle row
    Value4 v;
    v.u = vals_in[i * 3 + 0];
    v.v = vals_in[i * 3 + 1];
    v.w = vals_in[i * 3 + 2];
    v.pad = 0.0;
    vals[i] = v;
  }
  const double pz = pts[i * 3 + 2];
  if (pz < zlo || pz > zhi) {  // slab hash: outside the binned z-range
    cid[i] = -1;
    return;
  }
  const int cx = cell_coord(pts[i * 3 + 0], ox, inv_cell, cnx);
  const int cy = cell_coord(pts[i * 3 + 1], oy, inv_cell, cny);
  const int cz = cell_coord(pz, oz, inv_cell, cnz);
  const int32_t c = (cz * cny + cy) * cnx + cx;
  cid[i] = c;
  slot[i] = atomicAdd(&counts[c], 1);  // arrival order inside the cell; made canonical after the scatter
}

// ------------------------------------------------------------------------------ scan
__device__ __forceinline__ int block_exclusive_scan(int v, int* total_out) {
  // exclusive scan of one int per thread across a kScanThreads block
  __shared__ int warp_tot[kScanThreads / 32];
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
  for (int w2 = 0; w2 < kScanThreads / 32; ++w2) {
    int t = warp_tot[w2];
    if (w2 < wid) woff += t;
    tot += t;
  }
  __syncthreads();
  *total_out = tot;
  return woff + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int32_t* __restrict__ data,
                                                                    int64_t n,
                                                                    int32_t* __restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
    if (base + j < n) s += data[base + j];
  int tot;
  (void)block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) scan_blocksums_kernel(int32_t* __restrict__ block_sums,
                                                                       int nb) {
  int carry = 0;
  for (int base = 0; base < nb; base += kScanThreads) {
    const int i = base + threadIdx.x;
    const int v = (i < nb) ? block_sums[i] : 0;
    int tot;
    const int ex = block_exclusive_scan(v, &tot);
    if (i < nb) block_sums[i] = carry + ex;
    carry += tot;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(int32_t* __restrict__ data, int64_t n,
                                                                   const int32_t* __restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    v[j] = (base + j < n) ? data[base + j] : 0;
    s += v[j];
  }
  int tot;
  int ex = block_exclusive_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    if (base + j < n) data[base + j] = ex;
    ex += v[j];
  }
}

// ------------------------------------------------------------------------------ scatter
__global__ void __launch_bounds__(256) scatter_kernel(const int32_t* __restrict__ cid, int64_t n,
                                                       const int32_t* __restrict__ cell_start,
                                                       const int32_t* __restrict__ slot,
                                                       int32_t* __restrict__ sorted_idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t c = cid[i];
  if (c < 0) return;
  sorted_idx[cell_start[c] + slot[i]] = (int32_t)i;
}

// Canonical order inside each cell (ascending original index, == a stable sort on the cell id)
// so the staged order, and with it the floating-point summation order of the weights, does not
// depend on atomic arrival order.
__global__ void __launch_bounds__(256) cell_canonical_kernel(const int32_t* __restrict__ cell_start,
                                                              int64_t ncells,
                                                              int32_t* __restrict__ sorted_idx,
                                                              int32_t* __restrict__ max_count) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = c < ncells;
  const int32_t s = valid ? cell_start[c] : 0;
  const int32_t m = valid ? cell_start[c + 1] - s : 0;
  if (m > 1) {
    int32_t* a = sorted_idx + s;
    if (m <= 48) {
      for (int i = 1; i < m; ++i) {
        const int32_t key = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > key) {
          a[j + 1] = a[j];
          --j;
        }
        a[j + 1] = key;
      }
    } else {  // heap sort for pathological, heavily clustered cells
      for (int start = m / 2 - 1; start >= 0; --start) {
        int root = start;
        for (;;) {
          int child = 2 * root + 1;
          if (child >= m) break;
          if (child + 1 < m && a[child] < a[child + 1]) ++child;
          if (a[root] >= a[child]) break;
          int32_t t = a[root]; a[root] = a[child]; a[child] = t;
          root = child;
        }
      }
      for (int end = m - 1; end > 0; --end) {
        int32_t t = a[0]; a[0] = a[end]; a[end] = t;
        int root = 0;
        for (;;) {
          int child = 2 * root + 1;
          if (child >= end) break;
          if (child + 1 < end && a[child] < a[child + 1]) ++child;
          if (a[root] >= a[child]) break;
          int32_t t2 = a[root]; a[root] = a[child]; a[child] = t2;
          root = child;
        }
      }
    }
  }
  // one atomic per warp keeps contention negligible
  int wm = m;
  for (int o = 16; o > 0; o >>= 1) wm = max(wm, __shfl_xor_sync(0xffffffffu, wm, o));
  if ((threadIdx.x & 31) == 0 && wm > 0) atomicMax(max_count, wm);
}

__global__ void __launch_bounds__(256) fill_records_kernel(const double* __restrict__ pts,
                                                            const double* __restrict__ vals_in,
                                                            const int32_t* __restrict__ sorted_idx,
                                                            int64_t n, const int32_t* __restrict__ n_binned,
                                                            ParticleRec* __restrict__ rec,
                                                            Value4* __restrict__ vals_s64,
                                                            float4* __restrict__ vals_s32) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n || p >= *n_binned) return;  // slab hash: fewer than n particles are binned
  const int32_t i = sorted_idx[p];
  ParticleRec r;
  r.x = pts[(int64_t)i * 3 + 0];
  r.y = pts[(int64_t)i * 3 + 1];
  r.z = pts[(int64_t)i * 3 + 2];
  r.idx = i;
  r.pad = 0;
  rec[p] = r;
  Value4 v;
  v.u = vals_in[(int64_t)i * 3 + 0];
  v.v = vals_in[(int64_t)i * 3 + 1];
  v.w = vals_in[(int64_t)i * 3 + 2];
  v.pad = 0.0;
  vals_s64[p] = v;
  vals_s32[p] = make_float4((float)v.u, (float)v.v, (float)v.w, 0.0f);
}

static int ensure_capacity(ptv_hash* h, int64_t n, int64_t ncells) {
  if (n > h->cap_n) {
    const int64_t cap = n + n / 8 + 1024;
    cudaFree(h->rec); cudaFree(h->vals); cudaFree(h->cid); cudaFree(h->sorted_idx); cudaFree(h->slot);
    cudaFree(h->vals_s64); cudaFree(h->vals_s32);
    h->rec = nullptr; h->vals = nullptr; h->cid = nullptr; h->sorted_idx = nullptr; h->slot = nullptr;
    h->vals_s64 = nullptr; h->vals_s32 = nullptr;
    h->cap_n = 0;
    PTV_CUDA(cudaMalloc(&h->rec, cap * sizeof(ParticleRec)));
    PTV_CUDA(cudaMalloc(&h->vals, cap * sizeof(Value4)));
    PTV_CUDA(cudaMalloc(&h->vals_s64, cap * sizeof(Value4)));
    PTV_CUDA(cudaMalloc(&h->vals_s32, cap * sizeof(float4)));
    PTV_CUDA(cudaMalloc(&h->cid, cap * sizeof(int32_t)));
    PTV_CUDA(cudaMalloc(&h->sorted_idx, cap * sizeof(int32_t)));
    PTV_CUDA(cudaMalloc(&h->slot, cap * sizeof(int32_t)));
    h->cap_n = cap;
  }
  if (ncells + 1 > h->cap_cells) {
    const int64_t cap = ncells + ncells / 8 + 1024;
    cudaFree(h->cell_start); cudaFree(h->cell_fill); cudaFree(h->scan_tmp);
    h->cell_start = nullptr; h->cell_fill = nullptr; h->scan_tmp = nullptr;
    h->cap_cells = 0;
    PTV_CUDA(cudaMalloc(&h->cell_start, cap * sizeof(int32_t)));
    PTV_CUDA(cudaMalloc(&h->cell_fill, (cap + 1) * sizeof(int32_t)));  // +1: max_count slot
    PTV_CUDA(cudaMalloc(&h->scan_tmp, (cap / kScanTile + 2) * sizeof(int32_t)));
    h->cap_cells = cap;
  }
  return PTV_OK;
}

static void choose_cells(const double bb[6], int64_t n, double cell_size, double ppc, int dims[3],
                         double* cell_out) {
  double ext[3];
  for (int c = 0; c < 3; ++c) ext[c] = std::max(0.0, bb[3 + c] - bb[c]);
  double cell = cell_size;
  if (!(cell > 0.0)) {
    double vol = 1.0;
    int d = 0;
    for (int c = 0; c < 3; ++c)
      if (ext[c] > 0.0) { vol *= ext[c]; ++d; }
    if (d == 0) cell = 1.0;
    else cell = std::pow(vol * ppc / (double)std::max<int64_t>(n, 1), 1.0 / d);
    if (!(cell > 0.0) || !std::isfinite(cell)) cell = 1.0;
  }
  const double max_cells = std::min<double>(4.0 * (double)n + 4096.0, 1.0e9);
  for (;;) {
    double tot = 1.0;
    for (int c = 0; c < 3; ++c) {
      double m = std::floor(ext[c] / cell) + 1.0;
      m = std::min(m, 2.0e9);
      dims[c] = (int)std::max(1.0, std::min(m, 1.0e9));
      tot *= (double)dims[c];
    }
    if (tot <= max_cells) break;
    cell *= 1.26;
  }
  *cell_out = cell;
}

}  // namespace ptv

using namespace ptv;

ptv::HashGrid ptv_hash::view() const {
  HashGrid g;
  g.rec = rec; g.cell_start = cell_start; g.vals = vals; g.vals_s64 = vals_s64; g.vals_s32 = vals_s32; g.pts = pts;
  g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2];
  g.cell = cell; g.inv_cell = 1.0 / cell;
  g.cnx = dims[0]; g.cny = dims[1]; g.cnz = dims[2];
  g.n = n;
  g.clip_lo = clip_lo;
  g.clip_hi = clip_hi;
  return g;
}

extern "C" int ptv_hash_create(ptv_hash** out) {
  if (!out) { set_error("ptv_hash_create: out is NULL"); return PTV_ERR_INVALID; }
  int dev = 0;
  PTV_CUDA(cudaGetDevice(&dev));
  ptv_hash* h = new ptv_hash();
  cudaError_t e = cudaMalloc(&h->bbox_dev, (kBboxBlocks * 7 + 8) * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(h->bbox_dev, 0, 8 * sizeof(double));  // [7]: the bbox kernel's arrival counter
  if (e == cudaSuccess) e = cudaMallocHost(&h->bbox_host, 8 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&h->err_flag, sizeof(int));
  if (e == cudaSuccess) e = cudaMallocHost(&h->err_host, sizeof(int));
  if (e != cudaSuccess) {
    cudaFree(h->bbox_dev);
    delete h;
    return cuda_fail(e, "ptv_hash_create alloc", __FILE__, __LINE__);
  }
  *out = h;
  return PTV_OK;
}

extern "C" int ptv_hash_destroy(ptv_hash* h) {
  if (!h) return PTV_OK;
  cudaFree(h->rec); cudaFree(h->vals); cudaFree(h->cid); cudaFree(h->sorted_idx); cudaFree(h->slot);
  cudaFree(h->vals_s64); cudaFree(h->vals_s32);
  cudaFree(h->cell_start); cudaFree(h->cell_fill); cudaFree(h->scan_tmp);
  cudaFree(h->bbox_dev);
  cudaFree(h->err_flag);
  cudaFree(h->fail_list);
  cudaFree(h->fail_count);
  cudaFree(h->fail_flags);
  cudaFree(h->clip_count);
  cudaFree(h->hull_rec);
  cudaFree(h->hull_box);
  cudaFree(h->hull_rec2);
  cudaFree(h->hull_keep);
  cudaFree(h->hull_tab);
  if (h->bbox_host) cudaFreeHost(h->bbox_host);
  if (h->err_host) cudaFreeHost(h->err_host);
  delete h;
  return PTV_OK;
}

static int build_impl(ptv_hash* h, const double* d_points, const double* d_values, int64_t n, double cell_size,
                      bool slab, double z_lo, double z_hi, int k, double halo_factor, void* stream_);

extern "C" int ptv_hash_build(ptv_hash* h, const double* d_points, const double* d_values, int64_t n,
                              double cell_size, void* stream_) {
  return build_impl(h, d_points, d_values, n, cell_size, false, 0.0, 0.0, 0, 0.0, stream_);
}

extern "C" int ptv_hash_build_slab(ptv_hash* h, const double* d_points, const double* d_values, int64_t n,
                                   double cell_size, double z_lo, double z_hi, int k, double halo_factor,
                                   void* stream_) {
  if (!(z_hi >= z_lo) || k < 1 || !(halo_factor > 0.0)) {
    set_error("ptv_hash_build_slab: need z_lo <= z_hi, k >= 1, halo_factor > 0");
    return PTV_ERR_INVALID;
  }
  return build_impl(h, d_points, d_values, n, cell_size, true, z_lo, z_hi, k, halo_factor, stream_);
}

extern "C" int ptv_hash_clip_violations(const ptv_hash* h, int64_t* count) {
  if (!h || !count) { set_error("ptv_hash_clip_violations: NULL argument"); return PTV_ERR_INVALID; }
  int c = 0;
  if (h->clip_count != nullptr) PTV_CUDA(cudaMemcpy(&c, h->clip_count, sizeof(int), cudaMemcpyDeviceToHost));
  *count = c;
  return PTV_OK;
}

__global__ void clip_to_double_kernel(const int* __restrict__ count, double* __restrict__ dst) { *dst = (double)*count; }

extern "C" int ptv_hash_clip_violations_to(const ptv_hash* h, double* d_dst, void* stream_) {
  if (!h || !d_dst || h->clip_count == nullptr) { set_error("ptv_hash_clip_violations_to: NULL argument / hash not built"); return PTV_ERR_INVALID; }
  clip_to_double_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>(h->clip_count, d_dst);
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}

static int build_impl(ptv_hash* h, const double* d_points, const double* d_values, int64_t n, double cell_size,
                      bool slab, double z_lo, double z_hi, int k, double halo_factor, void* stream_) {
  if (!h || !d_points || !d_values) { set_error("ptv_hash_build: NULL argument"); return PTV_ERR_INVALID; }
  if (n <= 0) { set_error("ptv_hash_build: no particles"); return PTV_ERR_TOO_FEW; }
  if (n >= (int64_t)2147483000) { set_error("ptv_hash_build: more than 2^31 particles"); return PTV_ERR_INVALID; }
  cudaStream_t stream = (cudaStream_t)stream_;
  h->built = false;
  h->hull_valid = false;

  // 1. bounding box (and finiteness) of the cloud
  double* partial = h->bbox_dev + 8;
  bbox_partial_kernel<<<kBboxBlocks, kBboxThreads, 0, stream>>>(d_points, n, partial,
                                                                reinterpret_cast<unsigned*>(h->bbox_dev + 7), h->bbox_dev);
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  PTV_CUDA(cudaMemcpyAsync(h->bbox_host, h->bbox_dev, 7 * sizeof(double), cudaMemcpyDeviceToHost, stream));
  PTV_CUDA(cudaStreamSynchronize(stream));
  if (h->bbox_host[6] != 0.0) {
    // KDTree(points) raises ValueError("data must be finite ...") here
    set_error("data must be finite, check for nan or inf values");
    return PTV_ERR_INVALID;
  }
  double cell;
  int dims[3];
  choose_cells(h->bbox_host, n, cell_size, tuning().ppc, dims, &cell);
  // slab hash: bin only z in [z_lo - halo, z_hi + halo], halo = halo_factor x the radius expected to hold k
  // particles at the cloud's mean density; the cell edge stays the one of the whole cloud
  double zlo = -INFINITY, zhi = INFINITY;
  h->clip_lo = -INFINITY;
  h->clip_hi = INFINITY;
  if (slab) {
    double vol = 1.0;
    for (int c = 0; c < 3; ++c) vol *= std::max(h->bbox_host[3 + c] - h->bbox_host[c], cell);
    const double halo = halo_factor * std::cbrt(0.238732414637843 * k * vol / (double)n);
    zlo = z_lo - halo;
    zhi = z_hi + halo;
    if (zlo > h->bbox_host[2]) { h->clip_lo = zlo; h->bbox_host[2] = std::min(zlo, h->bbox_host[5]); } else zlo = -INFINITY;
    if (zhi < h->bbox_host[5]) { h->clip_hi = zhi; h->bbox_host[5] = std::max(zhi, h->bbox_host[2]); } else zhi = INFINITY;
    dims[2] = (int)std::max(1.0, std::floor((h->bbox_host[5] - h->bbox_host[2]) / cell) + 1.0);
  }
  if (h->clip_count == nullptr) PTV_CUDA(cudaMalloc(&h->clip_count, sizeof(int)));
  PTV_CUDA(cudaMemsetAsync(h->clip_count, 0, sizeof(int), stream));
  const int64_t ncells = (int64_t)dims[0] * dims[1] * dims[2];
  int rc = ensure_capacity(h, n, ncells);
  if (rc != PTV_OK) return rc;
  h->n = n;
  h->cell = cell;
  for (int c = 0; c < 3; ++c) { h->dims[c] = dims[c]; h->origin[c] = h->bbox_host[c]; }
  h->pts = d_points;

  // 2. cell ids + histogram
  PTV_CUDA(cudaMemsetAsync(h->cell_start, 0, (ncells + 1) * sizeof(int32_t), stream));
  PTV_CUDA(cudaMemsetAsync(h->cell_fill + ncells, 0, sizeof(int32_t), stream));  // max_count slot
  const int nb_p = (int)((n + 255) / 256);
  cell_count_kernel<<<nb_p, 256, 0, stream>>>(d_points, n, h->origin[0], h->origin[1], h->origin[2],
                                              1.0 / cell, dims[0], dims[1], dims[2], zlo, zhi, d_values, h->vals, h->cid,
                                              h->slot, h->cell_start);
  // 3. exclusive scan over ncells+1 entries (last entry becomes n)
  const int64_t nscan = ncells + 1;
  const int nb_s = (int)((nscan + kScanTile - 1) / kScanTile);
  scan_reduce_kernel<<<nb_s, kScanThreads, 0, stream>>>(h->cell_start, nscan, h->scan_tmp);
  scan_blocksums_kernel<<<1, kScanThreads, 0, stream>>>(h->scan_tmp, nb_s);
  scan_apply_kernel<<<nb_s, kScanThreads, 0, stream>>>(h->cell_start, nscan, h->scan_tmp);
  // 4. scatter, canonical order, records, values
  scatter_kernel<<<nb_p, 256, 0, stream>>>(h->cid, n, h->cell_start, h->slot, h->sorted_idx);
  int32_t* max_count = h->cell_fill + ncells;  // zeroed above
  cell_canonical_kernel<<<(int)((ncells + 255) / 256), 256, 0, stream>>>(h->cell_start, ncells,
                                                                        h->sorted_idx, max_count);
  fill_records_kernel<<<nb_p, 256, 0, stream>>>(d_points, d_values, h->sorted_idx, n, h->cell_start + ncells, h->rec,
                                                h->vals_s64, h->vals_s32);
  count_launches(7);
  PTV_CUDA(cudaGetLastError());
  h->max_cell_count = -1;  // fetched lazily by ptv_hash_info
  h->built = true;
  return PTV_OK;
}

extern "C" int ptv_hash_info(const ptv_hash* h, int64_t* n, int dims[3], double origin[3],
                             double* cell_size, int* max_cell_count) {
  if (!h || !h->built) { set_error("ptv_hash_info: hash not built"); return PTV_ERR_INVALID; }
  if (n) *n = h->n;
  for (int c = 0; c < 3; ++c) {
    if (dims) dims[c] = h->dims[c];
    if (origin) origin[c] = h->origin[c];
  }
  if (cell_size) *cell_size = h->cell;
  if (max_cell_count) {
    const int64_t ncells = (int64_t)h->dims[0] * h->dims[1] * h->dims[2];
    int32_t m = 0;
    PTV_CUDA(cudaMemcpy(&m, h->cell_fill + ncells, sizeof(int32_t), cudaMemcpyDeviceToHost));
    *max_cell_count = m;
  }
  return PTV_OK;
}
