// ptv_knn_interp: argument checks and the choice between the streaming kernel (IDW / sibson,
// k >= 8, no neighbour-list output) with its exact heap-kernel fallback, and the heap kernel alone.
#include <string>

#include "knn_common.cuh"

using namespace ptv;

static constexpr int kStatSlots = 24;  // [0] fail count, [1..] stats (see KnnParams::stats)

static int ensure_fail_buffers(ptv_hash* h, int64_t ntiles) {
  if (h->fail_cap < ntiles) {
    cudaFree(h->fail_list);
    h->fail_list = nullptr;
    h->fail_cap = 0;
    PTV_CUDA(cudaMalloc(&h->fail_list, (size_t)(ntiles + 1024) * sizeof(int)));
    h->fail_cap = ntiles + 1024;
  }
  if (h->fail_count == nullptr) PTV_CUDA(cudaMalloc(&h->fail_count, kStatSlots * sizeof(unsigned long long)));
  const int64_t words = (ntiles + 31) / 32 + 1;
  if (h->fail_flags_cap < words) {
    cudaFree(h->fail_flags);
    h->fail_flags = nullptr;
    h->fail_flags_cap = 0;
    PTV_CUDA(cudaMalloc(&h->fail_flags, (size_t)words * sizeof(unsigned)));
    h->fail_flags_cap = words;
  }
  return PTV_OK;
}

// PTV_METHOD_RBF* -> (PTV_METHOD_RBF, kernel id, monomials of the tail, minimum degree)
static bool rbf_variant(int& method, int& kern, int& npoly, int& degree) {
  kern = 0; npoly = 4; degree = 1;
  switch (method) {
    case PTV_METHOD_RBF: return true;
    case PTV_METHOD_RBF_CUBIC: kern = 1; break;
    case PTV_METHOD_RBF_LINEAR: kern = 2; npoly = 1; degree = 0; break;
    case PTV_METHOD_RBF_QUINTIC: kern = 3; npoly = 10; degree = 2; break;
    default: return false;
  }
  method = PTV_METHOD_RBF;
  return true;
}

static int rbf_check(int64_t n_particles, int& k, int npoly, int degree, double smoothing, const char* who) {
  if ((int64_t)k > n_particles) k = (int)n_particles;  // RBFInterpolator clamps neighbors to Np (scipy _rbfinterp.py:313)
  if (k < npoly) {
    set_error("At least " + std::to_string(npoly) + " data points are required when `degree` is " +
              std::to_string(degree) + " and the number of dimensions is 3.");
    return PTV_ERR_INVALID;
  }
  if (k + npoly > 64) {
    set_error(std::string(who) + ": rbf_neighbors > " + std::to_string(64 - npoly) + " is not supported on the CUDA path");
    return PTV_ERR_INVALID;
  }
  if (!(smoothing >= 0.0)) { set_error(std::string(who) + ": smoothing must be >= 0"); return PTV_ERR_INVALID; }
  return PTV_OK;
}

// griddata(method='linear') needs a non-degenerate triangulation: Qhull wants at least 5 points in 3-D
static int linear_check(int64_t n) {
  if (n < 5) {
    set_error("QH6214 qhull input error: not enough points(" + std::to_string(n) + ") to construct initial simplex (need 5)");
    return PTV_ERR_QHULL;
  }
  return PTV_OK;
}

// method='linear': hull-candidate list (once per hash build), statistics slots, one launch
static int run_linear(ptv_hash* h, KnnParams& p, bool f32, cudaStream_t stream) {
  // a cloud without extent along an axis has no tetrahedra: Qhull stops with QH6154 (coplanar / collinear input)
  for (int c = 0; c < 3; ++c) {
    if (h->bbox_host[3 + c] - h->bbox_host[c] == 0.0) {
      set_error("QH6154 Qhull precision error: Initial simplex is flat (the particles have no extent along axis " +
                std::to_string(c) + ")");
      return PTV_ERR_QHULL;
    }
  }
  int rc = ensure_fail_buffers(h, 0);
  if (rc != PTV_OK) return rc;
  PTV_CUDA(cudaMemsetAsync(h->fail_count, 0, kStatSlots * sizeof(unsigned long long), stream));
  p.stats = tuning().stats != 0 ? h->fail_count + 1 : nullptr;
  if (tuning().hull != 0) {
    rc = ensure_hull_list(h, stream);
    if (rc != PTV_OK) return rc;
    p.hull_rec = h->hull_list;
    p.hull_box = h->hull_box;
    p.hull_n = h->hull_n;
  }
  h->last_used_stream = false;
  p.k = tuning().linear_k;  // first radius = the one expected to hold this many particles
  PTV_CUDA(cudaMemsetAsync(h->err_flag, 0, sizeof(int), stream));
  rc = launch_delaunay_linear(p, f32, stream);
  if (rc != PTV_OK) return rc;
  // a voxel whose programme hit the pivot limit was written as 0: never silently (Qhull reports its own
  // precision failures as QhullError too)
  PTV_CUDA(cudaMemcpyAsync(h->err_host, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PTV_CUDA(cudaStreamSynchronize(stream));
  if (*h->err_host != 0) {
    set_error("method='linear': " + std::to_string(*h->err_host) +
              " voxel(s) could not be resolved (pivot limit reached: degenerate input?)");
    return PTV_ERR_QHULL;
  }
  return PTV_OK;
}

static void tile_shape(int T, int& tx, int& ty, int& tz) {
  tx = T == 128 ? 8 : 4; ty = 4; tz = T == 32 ? 2 : 4;
}

extern "C" int ptv_knn_interp(const ptv_hash* hc, const double* d_ax_x, int nx, const double* d_ax_y, int ny,
                              const double* d_ax_z, int nz, const uint8_t* d_mask, int method, int k,
                              double idw_power, double rbf_smoothing, int out_dtype, void* d_u, void* d_v,
                              void* d_w, int64_t* d_knn_idx, double* d_knn_dist, void* stream_) {
  ptv_hash* h = const_cast<ptv_hash*>(hc);  // scratch buffers of the handle are grown on demand
  if (!h || !h->built) { set_error("ptv_knn_interp: hash not built"); return PTV_ERR_INVALID; }
  if (method == PTV_METHOD_LINEAR && (h->clip_lo > -INFINITY || h->clip_hi < INFINITY)) {
    set_error("ptv_knn_interp: method='linear' needs the full hash (ptv_hash_build), not a slab hash");
    return PTV_ERR_INVALID;
  }
  if (!d_ax_x || !d_ax_y || !d_ax_z || !d_u || !d_v || !d_w) { set_error("ptv_knn_interp: NULL argument"); return PTV_ERR_INVALID; }
  if (nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_knn_interp: empty grid"); return PTV_ERR_INVALID; }
  if ((d_knn_idx == nullptr) != (d_knn_dist == nullptr)) { set_error("ptv_knn_interp: knn_idx and knn_dist must be given together"); return PTV_ERR_INVALID; }
  if (method == PTV_METHOD_NEAREST) k = 1;
  if (method == PTV_METHOD_LINEAR) {
    k = 4;
    const int rc0 = linear_check(h->n);
    if (rc0 != PTV_OK) return rc0;
  }
  int rbf_kern = 0, rbf_npoly = 4, rbf_degree = 1;
  const bool is_rbf = rbf_variant(method, rbf_kern, rbf_npoly, rbf_degree);
  if (method != PTV_METHOD_IDW && method != PTV_METHOD_SIBSON && method != PTV_METHOD_NEAREST && !is_rbf &&
      method != PTV_METHOD_LINEAR) {
    set_error("ptv_knn_interp: unsupported method");
    return PTV_ERR_INVALID;
  }
  if (is_rbf) {
    const int rc0 = rbf_check(h->n, k, rbf_npoly, rbf_degree, rbf_smoothing, "ptv_knn_interp");
    if (rc0 != PTV_OK) return rc0;
  }
  if (out_dtype != PTV_F32 && out_dtype != PTV_F64) { set_error("ptv_knn_interp: bad out_dtype"); return PTV_ERR_INVALID; }
  if (k < 1) { set_error("ptv_knn_interp: k must be >= 1"); return PTV_ERR_INVALID; }
  if ((int64_t)k > h->n) {
    // values[indices] with index == Np: IndexError in the reference (interpolator.py:150)
    set_error("index " + std::to_string(h->n) + " is out of bounds for axis 0 with size " + std::to_string(h->n));
    return PTV_ERR_TOO_FEW;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  KnnParams p;
  p.g = h->view();
  p.ax = d_ax_x; p.ay = d_ax_y; p.az = d_ax_z;
  p.nx = nx; p.ny = ny; p.nz = nz;
  p.mask = d_mask; p.method = method; p.k = k; p.power = idw_power;
  p.u = d_u; p.v = d_v; p.w = d_w;
  p.knn_idx = d_knn_idx; p.knn_dist = d_knn_dist;
  p.r0 = tuning().r0 < 0 ? 0 : tuning().r0;
  p.smoothing = rbf_smoothing;
  p.rbf_kernel = rbf_kern;
  p.rbf_npoly = rbf_npoly;
  p.rscale = tuning().rscale;
  p.err_flag = h->err_flag;
  p.clip_count = h->clip_count;
  p.fail_list = nullptr; p.fail_count = nullptr; p.tile_list = nullptr; p.tile_count = nullptr; p.stats = nullptr;
  p.fail_flags = nullptr; p.fail_tx = p.fail_ty = p.fail_tz = 1;
  p.qrec = nullptr; p.nq = 0; p.keep = nullptr; p.kth_dist = nullptr; p.mad_threshold = 0.0;
  p.tiles_x = p.tiles_y = p.tiles_z = 0;
  p.hull_rec = nullptr; p.hull_box = nullptr; p.hull_n = 0;
  if (method == PTV_METHOD_RBF) PTV_CUDA(cudaMemsetAsync(h->err_flag, 0, sizeof(int), stream));
  if (method == PTV_METHOD_LINEAR) return run_linear(h, p, out_dtype == PTV_F32, stream);

  const size_t smem_max = 227 * 1024;
  const bool f32 = out_dtype == PTV_F32;
  // the CTA-wide streaming kernel (tuning stream = 1, kept for comparison) cannot output neighbour lists
  const bool duo = tuning().stream >= 2;
  const bool use_stream = tuning().stream != 0 && (method == PTV_METHOD_IDW || method == PTV_METHOD_SIBSON) &&
                          k >= 8 && (duo || d_knn_idx == nullptr);
  int T = use_stream ? tuning().stream_tile : tuning().tile;
  if (T != 32 && T != 64 && T != 128) T = 128;
  if (use_stream) {
    while (T > 32 && knn_heap_smem_bytes(T, k, method) > smem_max) T >>= 1;  // the fallback must fit
  } else {
    while (T > 32 && knn_heap_smem_bytes(T, k, method) > smem_max / 2) T >>= 1;  // >= 2 CTAs per SM if possible
  }
  if (knn_heap_smem_bytes(T, k, method) > smem_max) {
    set_error("ptv_knn_interp: k too large for shared memory (max ~580)");
    return PTV_ERR_INVALID;
  }
  h->last_used_stream = use_stream;
  int rc;
  if (use_stream) {
    int tx, ty, tz;
    tile_shape(T, tx, ty, tz);
    const int64_t ntiles = (int64_t)((nx + tx - 1) / tx) * ((ny + ty - 1) / ty) * ((nz + tz - 1) / tz);
    if (ntiles > 2147483647LL) { set_error("ptv_knn_interp: grid too large for one launch"); return PTV_ERR_INVALID; }
    rc = ensure_fail_buffers(h, ntiles);
    if (rc != PTV_OK) return rc;
    PTV_CUDA(cudaMemsetAsync(h->fail_count, 0, kStatSlots * sizeof(unsigned long long), stream));
    p.fail_list = h->fail_list;
    p.fail_count = reinterpret_cast<int*>(h->fail_count);
    p.stats = tuning().stats != 0 ? h->fail_count + 1 : nullptr;
    if (duo) {
      PTV_CUDA(cudaMemsetAsync(h->fail_flags, 0, (size_t)((ntiles + 31) / 32 + 1) * sizeof(unsigned), stream));
      p.fail_flags = h->fail_flags;
      p.fail_tx = tx; p.fail_ty = ty; p.fail_tz = tz;
      rc = launch_knn_duo(p, f32, stream);
    } else {
      rc = launch_knn_stream(p, T, f32, stream);
    }
    if (rc != PTV_OK) return rc;
    // exact heap kernel over whatever the optimistic kernel could not finish
    p.tile_list = h->fail_list;
    p.tile_count = reinterpret_cast<const int*>(h->fail_count);
    p.stats = nullptr;
    rc = launch_knn_heap(p, T, f32, stream);
    // parity output: canonical order + distances for the rows the streaming kernel selected
    if (rc == PTV_OK && duo && d_knn_idx != nullptr) rc = launch_knn_sort_lists(p, stream);
  } else {
    rc = launch_knn_heap(p, T, f32, stream);
  }
  if (rc != PTV_OK) return rc;
  if (method == PTV_METHOD_RBF) {
    // a singular neighbourhood must surface as LinAlgError like scipy's dsysv info > 0 check
    PTV_CUDA(cudaMemcpyAsync(h->err_host, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PTV_CUDA(cudaStreamSynchronize(stream));
    if (*h->err_host != 0) { set_error("Singular matrix."); return PTV_ERR_SINGULAR; }
  }
  return PTV_OK;
}

static void init_point_params(KnnParams& p, const ptv_hash* h, const ptv_hash* q) {
  p.g = h->view();
  p.ax = p.ay = p.az = nullptr;
  p.nx = p.ny = p.nz = 1;
  p.mask = nullptr;
  p.power = 2.0;
  p.u = p.v = p.w = nullptr;
  p.knn_idx = nullptr; p.knn_dist = nullptr;
  p.tiles_x = p.tiles_y = p.tiles_z = 0;
  p.r0 = tuning().r0 < 0 ? 0 : tuning().r0;
  p.rscale = tuning().rscale;
  p.smoothing = 0.0;
  p.rbf_kernel = 0;
  p.rbf_npoly = 4;
  p.err_flag = h->err_flag;
  p.clip_count = h->clip_count;
  p.fail_list = nullptr; p.fail_count = nullptr; p.tile_list = nullptr; p.tile_count = nullptr; p.stats = nullptr;
  p.fail_flags = nullptr; p.fail_tx = p.fail_ty = p.fail_tz = 1;
  p.qrec = q->rec; p.nq = q->n;
  p.keep = nullptr; p.kth_dist = nullptr; p.mad_threshold = 0.0;
  p.hull_rec = nullptr; p.hull_box = nullptr; p.hull_n = 0;
}

static int pick_heap_tile(int k, int method) {
  const size_t smem_max = 227 * 1024;
  int T = 128;
  while (T > 32 && knn_heap_smem_bytes(T, k, method) > smem_max / 2) T >>= 1;
  return knn_heap_smem_bytes(T, k, method) > smem_max ? 0 : T;
}

extern "C" int ptv_knn_points(const ptv_hash* h, const ptv_hash* queries, int method, int k, double idw_power,
                              double rbf_smoothing, int out_dtype, void* d_u, void* d_v, void* d_w,
                              int64_t* d_knn_idx, double* d_knn_dist, void* stream_) {
  if (!h || !h->built || !queries || !queries->built) { set_error("ptv_knn_points: hash not built"); return PTV_ERR_INVALID; }
  const bool want_uvw = d_u || d_v || d_w;
  if (want_uvw && !(d_u && d_v && d_w)) { set_error("ptv_knn_points: give all of u, v, w or none"); return PTV_ERR_INVALID; }
  if ((d_knn_idx == nullptr) != (d_knn_dist == nullptr)) { set_error("ptv_knn_points: knn_idx and knn_dist must be given together"); return PTV_ERR_INVALID; }
  if (!want_uvw && !d_knn_idx) { set_error("ptv_knn_points: nothing to compute"); return PTV_ERR_INVALID; }
  if (method == PTV_METHOD_NEAREST) k = 1;
  if (method == PTV_METHOD_LINEAR) {
    k = 4;
    const int rc0 = linear_check(h->n);
    if (rc0 != PTV_OK) return rc0;
  }
  int rbf_kern = 0, rbf_npoly = 4, rbf_degree = 1;
  const bool is_rbf = rbf_variant(method, rbf_kern, rbf_npoly, rbf_degree);
  if (method != PTV_METHOD_IDW && method != PTV_METHOD_SIBSON && method != PTV_METHOD_NEAREST && !is_rbf &&
      method != PTV_METHOD_LINEAR) {
    set_error("ptv_knn_points: unsupported method");
    return PTV_ERR_INVALID;
  }
  if (is_rbf) {
    const int rc0 = rbf_check(h->n, k, rbf_npoly, rbf_degree, rbf_smoothing, "ptv_knn_points");
    if (rc0 != PTV_OK) return rc0;
  }
  if (out_dtype != PTV_F32 && out_dtype != PTV_F64) { set_error("ptv_knn_points: bad out_dtype"); return PTV_ERR_INVALID; }
  if (k < 1) { set_error("ptv_knn_points: k must be >= 1"); return PTV_ERR_INVALID; }
  if ((int64_t)k > h->n) {
    set_error("index " + std::to_string(h->n) + " is out of bounds for axis 0 with size " + std::to_string(h->n));
    return PTV_ERR_TOO_FEW;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  KnnParams p;
  init_point_params(p, h, queries);
  p.method = method; p.k = k; p.power = idw_power; p.smoothing = rbf_smoothing;
  p.rbf_kernel = rbf_kern; p.rbf_npoly = rbf_npoly;
  // u, v, w are always written by the kernel: give it scratch when only the lists are wanted (freed on
  // every return path, after the stream has drained)
  struct Scratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~Scratch() { if (p) { cudaStreamSynchronize(s); cudaFree(p); } }
  } guard;
  guard.s = stream;
  void*& scratch = guard.p;
  if (!want_uvw) {
    PTV_CUDA(cudaMalloc(&scratch, (size_t)queries->n * 8 * 3));
    d_u = scratch; d_v = (char*)scratch + (size_t)queries->n * 8; d_w = (char*)scratch + (size_t)queries->n * 16;
  }
  p.u = d_u; p.v = d_v; p.w = d_w;
  p.knn_idx = d_knn_idx; p.knn_dist = d_knn_dist;
  if (method == PTV_METHOD_LINEAR) {
    return run_linear(const_cast<ptv_hash*>(h), p, out_dtype == PTV_F32, stream);
  }
  const int T = pick_heap_tile(k, method);
  if (T == 0) { set_error("ptv_knn_points: k too large for shared memory (max ~580)"); return PTV_ERR_INVALID; }
  if (method == PTV_METHOD_RBF) PTV_CUDA(cudaMemsetAsync(h->err_flag, 0, sizeof(int), stream));
  const_cast<ptv_hash*>(h)->last_used_stream = false;
  int rc = launch_knn_heap(p, T, out_dtype == PTV_F32, stream);
  if (rc == PTV_OK && method == PTV_METHOD_RBF) {
    cudaError_t e = cudaMemcpyAsync(h->err_host, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "ptv_knn_points", __FILE__, __LINE__);
    else if (*h->err_host != 0) { set_error("Singular matrix."); rc = PTV_ERR_SINGULAR; }
  }
  return rc;
}

extern "C" int ptv_outlier_filter(const ptv_hash* h, int k, double threshold, uint8_t* d_keep, double* d_kth_dist,
                                  void* stream_) {
  if (!h || !h->built || !d_keep || !d_kth_dist) { set_error("ptv_outlier_filter: NULL argument / hash not built"); return PTV_ERR_INVALID; }
  if (k < 1) { set_error("ptv_outlier_filter: k must be >= 1"); return PTV_ERR_INVALID; }
  if (h->n <= k) {  // filtering.py:12-14 skips the filter; the shim never gets here
    set_error("ptv_outlier_filter: fewer than k+1 particles");
    return PTV_ERR_TOO_FEW;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  KnnParams p;
  init_point_params(p, h, h);
  p.method = PTV_METHOD_MADFILTER; p.k = k + 1;  // the particle itself is its own nearest neighbour
  p.keep = d_keep; p.kth_dist = d_kth_dist; p.mad_threshold = threshold;
  const int T = pick_heap_tile(k + 1, p.method);
  if (T == 0) { set_error("ptv_outlier_filter: k too large for shared memory"); return PTV_ERR_INVALID; }
  const_cast<ptv_hash*>(h)->last_used_stream = false;
  return launch_knn_heap(p, T, true, stream);
}

extern "C" int ptv_knn_fail_reasons(const ptv_hash* hc, int64_t reasons[4]) {
  ptv_hash* h = const_cast<ptv_hash*>(hc);
  if (!h || !reasons) { set_error("ptv_knn_fail_reasons: NULL argument"); return PTV_ERR_INVALID; }
  unsigned long long host[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (h->last_used_stream && h->fail_count != nullptr)
    PTV_CUDA(cudaMemcpy(host, h->fail_count, sizeof(host), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) reasons[i] = (int64_t)host[2 + i];
  h->last_stage_counts[0] = (int64_t)host[6];
  h->last_stage_counts[1] = (int64_t)host[7];
  return PTV_OK;
}

extern "C" int ptv_knn_work_stats(const ptv_hash* h, int64_t work[8]) {
  if (!h || !work) { set_error("ptv_knn_work_stats: NULL argument"); return PTV_ERR_INVALID; }
  unsigned long long host[kStatSlots] = {0};
  if (h->last_used_stream && h->fail_count != nullptr)
    PTV_CUDA(cudaMemcpy(host, h->fail_count, sizeof(host), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 8; ++i) work[i] = (int64_t)host[9 + i];
  return PTV_OK;
}

extern "C" int ptv_linear_stats(const ptv_hash* h, int64_t stats[8]) {
  if (!h || !stats) { set_error("ptv_linear_stats: NULL argument"); return PTV_ERR_INVALID; }
  unsigned long long host[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (h->fail_count != nullptr) PTV_CUDA(cudaMemcpy(host, h->fail_count, sizeof(host), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 7; ++i) stats[i] = (int64_t)host[1 + i];
  stats[7] = h->hull_valid ? h->hull_n : 0;
  return PTV_OK;
}

extern "C" int ptv_knn_stats(const ptv_hash* h, int64_t* used_stream, int64_t* tiles_failed, int64_t* tiles_streamed) {
  if (!h) { set_error("ptv_knn_stats: NULL handle"); return PTV_ERR_INVALID; }
  if (used_stream) *used_stream = h->last_used_stream ? 1 : 0;
  unsigned long long host[2] = {0, 0};
  if (h->last_used_stream && h->fail_count != nullptr)
    PTV_CUDA(cudaMemcpy(host, h->fail_count, sizeof(host), cudaMemcpyDeviceToHost));
  if (tiles_failed) *tiles_failed = (int64_t)(host[0] & 0xffffffffULL);
  if (tiles_streamed) *tiles_streamed = (int64_t)host[1];
  return PTV_OK;
}
