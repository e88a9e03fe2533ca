// Library-level C ABI: error slot, tuning knobs, device info and the host-buffer convenience
// entry point.  See include/ptv_b200.h for the contract of every symbol.
#include <string.h>

#include <atomic>
#include <string>

#include "ptv_internal.cuh"

namespace ptv {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
                 std::to_string(line) + ")";
  return e == cudaErrorMemoryAllocation ? PTV_ERR_NOMEM : PTV_ERR_CUDA;
}

Tuning& tuning() {
  static Tuning t;
  return t;
}

}  // namespace ptv

using namespace ptv;

extern "C" int ptv_version(void) { return 100; }

extern "C" int64_t ptv_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" const char* ptv_last_error(void) { return g_last_error.c_str(); }

extern "C" int ptv_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  cudaDeviceProp prop;
  PTV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  return PTV_OK;
}

extern "C" int ptv_set_tuning(const char* key, double value) {
  if (!key) { set_error("ptv_set_tuning: NULL key"); return PTV_ERR_INVALID; }
  Tuning& t = tuning();
  if (!strcmp(key, "ppc")) { if (!(value > 0)) { set_error("ppc must be > 0"); return PTV_ERR_INVALID; } t.ppc = value; }
  else if (!strcmp(key, "r0")) t.r0 = (int)value;
  else if (!strcmp(key, "tile")) t.tile = (int)value;
  else if (!strcmp(key, "stream")) t.stream = (int)value;
  else if (!strcmp(key, "stream_tile")) t.stream_tile = (int)value;
  else if (!strcmp(key, "stats")) t.stats = (int)value;
  else if (!strcmp(key, "hull")) t.hull = (int)value;
  else if (!strcmp(key, "linear_occ")) t.linear_occ = (int)value;
  else if (!strcmp(key, "linear_k")) { if (!(value >= 4 && value <= 4096)) { set_error("linear_k must be in [4, 4096]"); return PTV_ERR_INVALID; } t.linear_k = (int)value; }
  else if (!strcmp(key, "rscale")) { if (!(value >= 1.0 && value <= 4.0)) { set_error("rscale must be in [1, 4]"); return PTV_ERR_INVALID; } t.rscale = value; }
  else if (!strcmp(key, "stencil_bulk")) t.stencil_bulk = (int)value;
  else if (!strcmp(key, "rbf_regs")) t.rbf_regs = (int)value;
  else if (!strcmp(key, "stencil_la")) t.stencil_la = (int)value;
  else { set_error(std::string("ptv_set_tuning: unknown key ") + key); return PTV_ERR_INVALID; }
  return PTV_OK;
}

extern "C" double ptv_get_tuning(const char* key) {
  if (!key) return nan("");
  const Tuning& t = tuning();
  if (!strcmp(key, "ppc")) return t.ppc;
  if (!strcmp(key, "r0")) return t.r0;
  if (!strcmp(key, "tile")) return t.tile;
  if (!strcmp(key, "stream")) return t.stream;
  if (!strcmp(key, "stream_tile")) return t.stream_tile;
  if (!strcmp(key, "stats")) return t.stats;
  if (!strcmp(key, "hull")) return t.hull;
  if (!strcmp(key, "linear_occ")) return t.linear_occ;
  if (!strcmp(key, "linear_k")) return t.linear_k;
  if (!strcmp(key, "rscale")) return t.rscale;
  if (!strcmp(key, "stencil_bulk")) return t.stencil_bulk;
  if (!strcmp(key, "rbf_regs")) return t.rbf_regs;
  return nan("");
}

namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename U> U* as() { return reinterpret_cast<U*>(p); }
};
struct HashGuard {
  ptv_hash* h = nullptr;
  ~HashGuard() { if (h) ptv_hash_destroy(h); }
};
}  // namespace

extern "C" int ptv_interpolate_host(const double* h_points, const double* h_values, int64_t n,
                                    const double* h_ax_x, int nx, const double* h_ax_y, int ny,
                                    const double* h_ax_z, int nz, const uint8_t* h_mask, int method, int k,
                                    double idw_power, double rbf_smoothing, int out_dtype, void* h_u,
                                    void* h_v, void* h_w) {
  if (!h_points || !h_values || !h_ax_x || !h_ax_y || !h_ax_z || !h_u || !h_v || !h_w) {
    set_error("ptv_interpolate_host: NULL argument");
    return PTV_ERR_INVALID;
  }
  if (n <= 0) { set_error("ptv_interpolate_host: no particles"); return PTV_ERR_TOO_FEW; }
  if (nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_interpolate_host: empty grid"); return PTV_ERR_INVALID; }
  if (out_dtype != PTV_F32 && out_dtype != PTV_F64) { set_error("ptv_interpolate_host: bad out_dtype"); return PTV_ERR_INVALID; }
  const int64_t nvox = (int64_t)nx * ny * nz;
  const size_t esz = out_dtype == PTV_F32 ? 4 : 8;
  DevBuf pts, vals, axes, mask, out;
  PTV_CUDA(pts.alloc((size_t)n * 24));
  PTV_CUDA(vals.alloc((size_t)n * 24));
  PTV_CUDA(axes.alloc((size_t)(nx + ny + nz) * 8));
  if (h_mask) PTV_CUDA(mask.alloc((size_t)nvox));
  PTV_CUDA(out.alloc((size_t)nvox * esz * 3));
  cudaStream_t s = nullptr;
  PTV_CUDA(cudaMemcpyAsync(pts.p, h_points, (size_t)n * 24, cudaMemcpyHostToDevice, s));
  PTV_CUDA(cudaMemcpyAsync(vals.p, h_values, (size_t)n * 24, cudaMemcpyHostToDevice, s));
  double* dax = axes.as<double>();
  PTV_CUDA(cudaMemcpyAsync(dax, h_ax_x, (size_t)nx * 8, cudaMemcpyHostToDevice, s));
  PTV_CUDA(cudaMemcpyAsync(dax + nx, h_ax_y, (size_t)ny * 8, cudaMemcpyHostToDevice, s));
  PTV_CUDA(cudaMemcpyAsync(dax + nx + ny, h_ax_z, (size_t)nz * 8, cudaMemcpyHostToDevice, s));
  if (h_mask) PTV_CUDA(cudaMemcpyAsync(mask.p, h_mask, (size_t)nvox, cudaMemcpyHostToDevice, s));
  HashGuard hg;
  int rc = ptv_hash_create(&hg.h);
  if (rc != PTV_OK) return rc;
  rc = ptv_hash_build(hg.h, pts.as<double>(), vals.as<double>(), n, 0.0, s);
  if (rc != PTV_OK) return rc;
  char* o = out.as<char>();
  rc = ptv_knn_interp(hg.h, dax, nx, dax + nx, ny, dax + nx + ny, nz, h_mask ? mask.as<uint8_t>() : nullptr,
                      method, k, idw_power, rbf_smoothing, out_dtype, o, o + (size_t)nvox * esz,
                      o + 2 * (size_t)nvox * esz, nullptr, nullptr, s);
  if (rc != PTV_OK) return rc;
  PTV_CUDA(cudaMemcpyAsync(h_u, o, (size_t)nvox * esz, cudaMemcpyDeviceToHost, s));
  PTV_CUDA(cudaMemcpyAsync(h_v, o + (size_t)nvox * esz, (size_t)nvox * esz, cudaMemcpyDeviceToHost, s));
  PTV_CUDA(cudaMemcpyAsync(h_w, o + 2 * (size_t)nvox * esz, (size_t)nvox * esz, cudaMemcpyDeviceToHost, s));
  PTV_CUDA(cudaStreamSynchronize(s));
  return PTV_OK;
}
