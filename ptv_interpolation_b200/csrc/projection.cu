// Projection cleaning of the interpolated field (physics.py:55-209):
//   matrix-free masked 7-point Laplacian  (build_laplacian_matrix, physics.py:55-108),
//   LSQR on it                            (scipy.sparse.linalg.lsqr as called at physics.py:186:
//                                          damp=1e-8, atol=btol=1e-10, iter_lim=3000, conlim=1e8),
//   staggered-gradient velocity correction (apply_consistent_correction, physics.py:110-147).
//
// All vectors live on the full (nz,ny,nx) grid in float64 with zeros in solid voxels, so A.v is a
// stencil and no index map is needed.  The LSQR scalars stay on the device: two one-thread kernels per
// iteration update the Golub-Kahan / Givens recurrences and the stopping tests exactly as SciPy does,
// every vector kernel reads them from the state block, and kernels become no-ops once istop != 0 --
// the host only polls the state every few iterations.  Norms are reduced deterministically
// (per-block partials summed in a fixed order).  Every kernel is one HBM-bound streaming pass.
#include <math.h>

#include "ptv_internal.cuh"

namespace ptv {

static constexpr int kRedBlocks = 1184;  // 8 x 148 SMs
static constexpr int kRedThreads = 256;

struct LsqrState {
  double alfa, beta, rhobar, phibar, anorm, ddnorm, res2, xnorm, xxnorm, z, cs2, sn2;
  double t1, t2, bnorm, damp, dampsq, atol, btol, ctol;
  double rnorm, arnorm, acond, r1norm, w2;
  double mean;  // mean of the right-hand side over fluid voxels (subtracted, physics.py:181)
  double n_fluid;
  int itn, istop, iter_lim, pad;
};

__device__ __forceinline__ void block_partial(double v, double* partial) {
  __shared__ double sh[kRedThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < kRedThreads / 32; ++i) a += sh[i];
    partial[blockIdx.x] = a;
  }
}

__device__ __forceinline__ double sum_partials(const double* partial) {  // fixed order, one thread
  double a = 0.0;
  for (int i = 0; i < kRedBlocks; ++i) a += partial[i];
  return a;
}

// (A x)_i = sum over the six neighbours j that are fluid of (x_j - x_i) / h^2   (physics.py:76-100)
__device__ __forceinline__ double lap_at(const double* __restrict__ x, const uint8_t* __restrict__ m, int64_t i, int ix,
                                         int iy, int iz, int nx, int ny, int nz, double ax, double ay, double az) {
  const int64_t sy = nx, sz = (int64_t)nx * ny;
  const double xc = x[i];
  double s = 0.0;
  if (ix > 0 && m[i - 1]) s += ax * (x[i - 1] - xc);
  if (ix + 1 < nx && m[i + 1]) s += ax * (x[i + 1] - xc);
  if (iy > 0 && m[i - sy]) s += ay * (x[i - sy] - xc);
  if (iy + 1 < ny && m[i + sy]) s += ay * (x[i + sy] - xc);
  if (iz > 0 && m[i - sz]) s += az * (x[i - sz] - xc);
  if (iz + 1 < nz && m[i + sz]) s += az * (x[i + sz] - xc);
  return s;
}

#define PTV_GRID_LOOP(i, n) \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

// ---- right-hand side: b = (div - mean(div[mask])) on fluid voxels, 0 elsewhere -------------------
template <typename Tf>
__global__ void __launch_bounds__(kRedThreads) rhs_sum_kernel(const Tf* __restrict__ div, const uint8_t* __restrict__ m,
                                                               int64_t n, double* __restrict__ p_sum,
                                                               double* __restrict__ p_cnt) {
  double s = 0.0, c = 0.0;
  PTV_GRID_LOOP(i, n) if (m[i]) { s += (double)div[i]; c += 1.0; }
  block_partial(s, p_sum);
  __syncthreads();
  block_partial(c, p_cnt);
}

__global__ void lsqr_setup_kernel(LsqrState* st, const double* p_sum, const double* p_cnt, double damp, double atol,
                                  double btol, double conlim, int iter_lim) {
  const double cnt = sum_partials(p_cnt);
  st->n_fluid = cnt;
  st->mean = cnt > 0.0 ? sum_partials(p_sum) / cnt : 0.0;
  st->damp = damp; st->dampsq = damp * damp; st->atol = atol; st->btol = btol;
  st->ctol = conlim > 0.0 ? 1.0 / conlim : 0.0;
  st->iter_lim = iter_lim;
  st->itn = 0; st->istop = 0;
  st->anorm = 0.0; st->acond = 0.0; st->ddnorm = 0.0; st->res2 = 0.0; st->xnorm = 0.0; st->xxnorm = 0.0;
  st->z = 0.0; st->cs2 = -1.0; st->sn2 = 0.0; st->t1 = 0.0; st->t2 = 0.0;
}

template <typename Tf>
__global__ void __launch_bounds__(kRedThreads) rhs_init_kernel(const Tf* __restrict__ div, const uint8_t* __restrict__ m,
                                                                int64_t n, const LsqrState* __restrict__ st,
                                                                double* __restrict__ u, double* __restrict__ x,
                                                                double* __restrict__ partial) {
  const double mean = st->mean;
  double s = 0.0;
  PTV_GRID_LOOP(i, n) {
    const double b = m[i] ? (double)div[i] - mean : 0.0;
    u[i] = b;
    x[i] = 0.0;
    s += b * b;
  }
  block_partial(s, partial);
}

__global__ void lsqr_beta0_kernel(LsqrState* st, const double* partial) {
  const double b = sqrt(sum_partials(partial));
  st->bnorm = b;
  st->beta = b;
}

// v_raw = A u_raw / beta - beta_scale * v_raw / alfa   (first call: beta_scale = 0)
__global__ void __launch_bounds__(kRedThreads) lsqr_v_kernel(const double* __restrict__ u, double* __restrict__ v,
                                                              const uint8_t* __restrict__ m, int nx, int ny, int nz,
                                                              double ax, double ay, double az,
                                                              const LsqrState* __restrict__ st, int first,
                                                              double* __restrict__ partial) {
  double s = 0.0;
  if (st->istop == 0) {
    const double beta = st->beta, alfa = st->alfa;
    const double ib = beta > 0.0 ? 1.0 / beta : 0.0;
    const double ia = (!first && alfa > 0.0) ? 1.0 / alfa : 0.0;
    const int64_t n = (int64_t)nx * ny * nz;
    PTV_GRID_LOOP(i, n) {
      double r = 0.0;
      if (m[i]) {
        const int ix = (int)(i % nx), iy = (int)((i / nx) % ny), iz = (int)(i / ((int64_t)nx * ny));
        r = lap_at(u, m, i, ix, iy, iz, nx, ny, nz, ax, ay, az) * ib;
        if (!first) r -= beta * (v[i] * ia);
        if (beta <= 0.0) r = first ? 0.0 : v[i] * ia;  // lsqr.py: v is left unchanged when beta == 0
      }
      v[i] = r;
      s += r * r;
    }
  }
  block_partial(s, partial);
}

// u_raw = A v_raw / alfa - alfa * u_raw / beta
__global__ void __launch_bounds__(kRedThreads) lsqr_u_kernel(double* __restrict__ u, const double* __restrict__ v,
                                                              const uint8_t* __restrict__ m, int nx, int ny, int nz,
                                                              double ax, double ay, double az,
                                                              const LsqrState* __restrict__ st,
                                                              double* __restrict__ partial) {
  double s = 0.0;
  if (st->istop == 0) {
    const double beta = st->beta, alfa = st->alfa;
    const double ib = beta > 0.0 ? 1.0 / beta : 0.0, ia = alfa > 0.0 ? 1.0 / alfa : 0.0;
    const int64_t n = (int64_t)nx * ny * nz;
    PTV_GRID_LOOP(i, n) {
      double r = 0.0;
      if (m[i]) {
        const int ix = (int)(i % nx), iy = (int)((i / nx) % ny), iz = (int)(i / ((int64_t)nx * ny));
        r = lap_at(v, m, i, ix, iy, iz, nx, ny, nz, ax, ay, az) * ia - alfa * (u[i] * ib);
      }
      u[i] = r;
      s += r * r;
    }
  }
  block_partial(s, partial);
}

// after lsqr_u_kernel: beta = ||u_raw||, anorm update (lsqr.py "if beta > 0")
__global__ void lsqr_scalar_beta_kernel(LsqrState* st, const double* partial) {
  if (st->istop != 0) return;
  st->itn += 1;
  const double beta = sqrt(sum_partials(partial));
  if (beta > 0.0) st->anorm = sqrt(st->anorm * st->anorm + st->alfa * st->alfa + beta * beta + st->dampsq);
  st->beta = beta;
}

__device__ __forceinline__ double sgn(double a) { return (a > 0.0) - (a < 0.0); }

// after lsqr_v_kernel: alfa = ||v_raw||, plane rotations, step lengths, stopping tests (lsqr.py main loop)
__global__ void lsqr_scalar_alfa_kernel(LsqrState* st, const double* partial, int first) {
  if (st->istop != 0) return;
  const double eps = 2.220446049250313e-16;
  const double alfa_new = sqrt(sum_partials(partial));
  if (first) {  // initialisation: v = A^T u, w = v, rhobar = alfa, phibar = beta
    st->alfa = alfa_new;
    st->rhobar = alfa_new;
    st->phibar = st->beta;
    st->rnorm = st->beta;
    st->arnorm = alfa_new * st->beta;
    st->w2 = alfa_new > 0.0 ? 1.0 : 0.0;  // ||w||^2 with w = v normalised
    if (st->arnorm == 0.0) st->istop = -1;  // "The exact solution is x = 0"
    return;
  }
  const double beta = st->beta;
  const double alfa = beta > 0.0 ? alfa_new : st->alfa;  // v (hence alfa) unchanged when beta == 0
  double rhobar1, psi;
  if (st->damp > 0.0) {
    rhobar1 = sqrt(st->rhobar * st->rhobar + st->dampsq);
    const double cs1 = st->rhobar / rhobar1, sn1 = st->damp / rhobar1;
    psi = sn1 * st->phibar;
    st->phibar = cs1 * st->phibar;
  } else {
    rhobar1 = st->rhobar;
    psi = 0.0;
  }
  double cs, sn, rho;  // _sym_ortho(rhobar1, beta)
  if (beta == 0.0) { cs = sgn(rhobar1); sn = 0.0; rho = fabs(rhobar1); }
  else if (rhobar1 == 0.0) { cs = 0.0; sn = sgn(beta); rho = fabs(beta); }
  else if (fabs(beta) > fabs(rhobar1)) {
    const double tau = rhobar1 / beta;
    sn = sgn(beta) / sqrt(1.0 + tau * tau); cs = sn * tau; rho = beta / sn;
  } else {
    const double tau = beta / rhobar1;
    cs = sgn(rhobar1) / sqrt(1.0 + tau * tau); sn = cs * tau; rho = rhobar1 / cs;
  }
  const double theta = sn * alfa;
  st->rhobar = -cs * alfa;
  const double phi = cs * st->phibar;
  st->phibar = sn * st->phibar;
  const double tau = sn * phi;
  st->t1 = phi / rho;
  st->t2 = -theta / rho;
  st->ddnorm = st->ddnorm + st->w2 / (rho * rho);  // ||dk||^2 with dk = w / rho
  const double delta = st->sn2 * rho, gambar = -st->cs2 * rho;
  const double rhs = phi - delta * st->z;
  const double zbar = rhs / gambar;
  st->xnorm = sqrt(st->xxnorm + zbar * zbar);
  const double gamma = sqrt(gambar * gambar + theta * theta);
  st->cs2 = gambar / gamma;
  st->sn2 = theta / gamma;
  st->z = rhs / gamma;
  st->xxnorm = st->xxnorm + st->z * st->z;
  st->acond = st->anorm * sqrt(st->ddnorm);
  const double res1 = st->phibar * st->phibar;
  st->res2 = st->res2 + psi * psi;
  st->rnorm = sqrt(res1 + st->res2);
  st->arnorm = alfa * fabs(tau);
  if (st->damp > 0.0) {
    const double r1sq = st->rnorm * st->rnorm - st->dampsq * st->xxnorm;
    st->r1norm = r1sq < 0.0 ? -sqrt(fabs(r1sq)) : sqrt(fabs(r1sq));
  } else {
    st->r1norm = st->rnorm;
  }
  const double test1 = st->rnorm / st->bnorm;
  const double test2 = st->arnorm / (st->anorm * st->rnorm + eps);
  const double test3 = 1.0 / (st->acond + eps);
  const double t1c = test1 / (1.0 + st->anorm * st->xnorm / st->bnorm);
  const double rtol = st->btol + st->atol * st->anorm * st->xnorm / st->bnorm;
  int istop = 0;
  if (st->itn >= st->iter_lim) istop = 7;
  if (1.0 + test3 <= 1.0) istop = 6;
  if (1.0 + test2 <= 1.0) istop = 5;
  if (1.0 + t1c <= 1.0) istop = 4;
  if (test3 <= st->ctol) istop = 3;
  if (test2 <= st->atol) istop = 2;
  if (test1 <= rtol) istop = 1;
  st->alfa = alfa;
  st->pad = istop;  // published by the x/w update so that this iteration's step is still applied
}

// x += t1 w ; w = v_raw / alfa + t2 w ; ||w||^2 for the next iteration's ddnorm
__global__ void __launch_bounds__(kRedThreads) lsqr_xw_kernel(double* __restrict__ x, double* __restrict__ w,
                                                               const double* __restrict__ v, int64_t n,
                                                               const LsqrState* __restrict__ st, int first,
                                                               double* __restrict__ partial) {
  double s = 0.0;
  if (st->istop == 0) {
    const double alfa = st->alfa, ia = alfa > 0.0 ? 1.0 / alfa : 0.0;
    const double t1 = st->t1, t2 = st->t2;
    PTV_GRID_LOOP(i, n) {
      const double vn = v[i] * ia;
      double wn;
      if (first) {
        wn = vn;
      } else {
        const double wo = w[i];
        x[i] += t1 * wo;
        wn = vn + t2 * wo;
      }
      w[i] = wn;
      s += wn * wn;
    }
  }
  block_partial(s, partial);
}

__global__ void lsqr_scalar_w_kernel(LsqrState* st, const double* partial, int first) {
  if (st->istop != 0) return;
  st->w2 = sum_partials(partial);
  if (!first) st->istop = st->pad;
}

// ---- velocity correction (physics.py:110-147) ---------------------------------------------------
template <typename Tf>
__device__ __forceinline__ double cell_grad(const double* __restrict__ p, const uint8_t* __restrict__ m, int64_t i,
                                            int64_t stride, int pos, int n, double h) {
  // g_next = (m[i+1] & m[i]) ? (p[i+1] - p[i]) / h : 0, zero on the last cell; g_prev = g_next of cell i-1
  double gn = 0.0, gp = 0.0;
  if (pos + 1 < n && m[i + stride] && m[i]) gn = (p[i + stride] - p[i]) / h;
  if (pos > 0 && m[i] && m[i - stride]) gp = (p[i] - p[i - stride]) / h;
  return (gn + gp) / 2.0;
}

template <typename Tf>
__global__ void __launch_bounds__(256) correct_kernel(const Tf* __restrict__ u, const Tf* __restrict__ v,
                                                       const Tf* __restrict__ w, const double* __restrict__ phi,
                                                       const uint8_t* __restrict__ m, int nx, int ny, int nz, double dx,
                                                       double dy, double dz, Tf* __restrict__ uo, Tf* __restrict__ vo,
                                                       Tf* __restrict__ wo) {
  const int64_t n = (int64_t)nx * ny * nz;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!m[i]) {
    uo[i] = (Tf)0; vo[i] = (Tf)0; wo[i] = (Tf)0;
    return;
  }
  const int ix = (int)(i % nx), iy = (int)((i / nx) % ny), iz = (int)(i / ((int64_t)nx * ny));
  uo[i] = (Tf)((double)u[i] - cell_grad<Tf>(phi, m, i, 1, ix, nx, dx));
  vo[i] = (Tf)((double)v[i] - cell_grad<Tf>(phi, m, i, nx, iy, ny, dy));
  wo[i] = (Tf)((double)w[i] - cell_grad<Tf>(phi, m, i, (int64_t)nx * ny, iz, nz, dz));
}

}  // namespace ptv

using namespace ptv;

extern "C" int64_t ptv_poisson_workspace_bytes(int nx, int ny, int nz) {
  const int64_t n = (int64_t)nx * ny * nz;
  return 3 * n * (int64_t)sizeof(double) + (int64_t)(2 * kRedBlocks) * sizeof(double) + 1024;
}

extern "C" int ptv_poisson_lsqr(const void* d_div, int dtype, const uint8_t* d_mask, int nx, int ny, int nz, double dx,
                                double dy, double dz, double damp, double atol, double btol, double conlim,
                                int iter_lim, double* d_phi, void* d_work, double* h_info, void* stream_) {
  if (!d_div || !d_mask || !d_phi || !d_work) { set_error("ptv_poisson_lsqr: NULL argument"); return PTV_ERR_INVALID; }
  if (nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_poisson_lsqr: empty grid"); return PTV_ERR_INVALID; }
  if (dtype != PTV_F32 && dtype != PTV_F64) { set_error("ptv_poisson_lsqr: bad dtype"); return PTV_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream_;
  const int64_t n = (int64_t)nx * ny * nz;
  double* u = reinterpret_cast<double*>(d_work);
  double* v = u + n;
  double* w = v + n;
  double* p0 = w + n;
  double* p1 = p0 + kRedBlocks;
  LsqrState* st = reinterpret_cast<LsqrState*>(p1 + kRedBlocks);
  const double ax = 1.0 / (dx * dx), ay = 1.0 / (dy * dy), az = 1.0 / (dz * dz);
  if (dtype == PTV_F32) rhs_sum_kernel<float><<<kRedBlocks, kRedThreads, 0, s>>>((const float*)d_div, d_mask, n, p0, p1);
  else rhs_sum_kernel<double><<<kRedBlocks, kRedThreads, 0, s>>>((const double*)d_div, d_mask, n, p0, p1);
  lsqr_setup_kernel<<<1, 1, 0, s>>>(st, p0, p1, damp, atol, btol, conlim, iter_lim);
  if (dtype == PTV_F32) rhs_init_kernel<float><<<kRedBlocks, kRedThreads, 0, s>>>((const float*)d_div, d_mask, n, st, u, d_phi, p0);
  else rhs_init_kernel<double><<<kRedBlocks, kRedThreads, 0, s>>>((const double*)d_div, d_mask, n, st, u, d_phi, p0);
  lsqr_beta0_kernel<<<1, 1, 0, s>>>(st, p0);
  lsqr_v_kernel<<<kRedBlocks, kRedThreads, 0, s>>>(u, v, d_mask, nx, ny, nz, ax, ay, az, st, 1, p0);
  lsqr_scalar_alfa_kernel<<<1, 1, 0, s>>>(st, p0, 1);
  lsqr_xw_kernel<<<kRedBlocks, kRedThreads, 0, s>>>(d_phi, w, v, n, st, 1, p0);
  lsqr_scalar_w_kernel<<<1, 1, 0, s>>>(st, p0, 1);
  count_launches(8);
  PTV_CUDA(cudaGetLastError());
  LsqrState host;
  const int batch = 16;
  for (int it = 0; it < iter_lim; it += batch) {
    for (int b = 0; b < batch && it + b < iter_lim; ++b) {
      lsqr_u_kernel<<<kRedBlocks, kRedThreads, 0, s>>>(u, v, d_mask, nx, ny, nz, ax, ay, az, st, p0);
      lsqr_scalar_beta_kernel<<<1, 1, 0, s>>>(st, p0);
      lsqr_v_kernel<<<kRedBlocks, kRedThreads, 0, s>>>(u, v, d_mask, nx, ny, nz, ax, ay, az, st, 0, p0);
      lsqr_scalar_alfa_kernel<<<1, 1, 0, s>>>(st, p0, 0);
      lsqr_xw_kernel<<<kRedBlocks, kRedThreads, 0, s>>>(d_phi, w, v, n, st, 0, p0);
      lsqr_scalar_w_kernel<<<1, 1, 0, s>>>(st, p0, 0);
      count_launches(6);
    }
    PTV_CUDA(cudaMemcpyAsync(&host, st, sizeof(LsqrState), cudaMemcpyDeviceToHost, s));
    PTV_CUDA(cudaStreamSynchronize(s));
    if (host.istop != 0) break;
  }
  PTV_CUDA(cudaMemcpyAsync(&host, st, sizeof(LsqrState), cudaMemcpyDeviceToHost, s));
  PTV_CUDA(cudaStreamSynchronize(s));
  if (h_info) {
    h_info[0] = host.istop < 0 ? 0 : host.istop;
    h_info[1] = host.itn;
    h_info[2] = host.r1norm; h_info[3] = host.rnorm; h_info[4] = host.anorm; h_info[5] = host.acond;
    h_info[6] = host.arnorm; h_info[7] = host.xnorm;
  }
  return PTV_OK;
}

extern "C" int ptv_projection_correct(const void* d_u, const void* d_v, const void* d_w, const double* d_phi,
                                      const uint8_t* d_mask, int nx, int ny, int nz, double dx, double dy, double dz,
                                      int dtype, void* d_uo, void* d_vo, void* d_wo, void* stream) {
  if (!d_u || !d_v || !d_w || !d_phi || !d_mask || !d_uo || !d_vo || !d_wo) { set_error("ptv_projection_correct: NULL argument"); return PTV_ERR_INVALID; }
  if (nx <= 0 || ny <= 0 || nz <= 0) { set_error("ptv_projection_correct: empty grid"); return PTV_ERR_INVALID; }
  const int64_t n = (int64_t)nx * ny * nz;
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (dtype == PTV_F32)
    correct_kernel<float><<<nb, 256, 0, (cudaStream_t)stream>>>((const float*)d_u, (const float*)d_v, (const float*)d_w,
                                                                d_phi, d_mask, nx, ny, nz, dx, dy, dz, (float*)d_uo,
                                                                (float*)d_vo, (float*)d_wo);
  else if (dtype == PTV_F64)
    correct_kernel<double><<<nb, 256, 0, (cudaStream_t)stream>>>((const double*)d_u, (const double*)d_v,
                                                                 (const double*)d_w, d_phi, d_mask, nx, ny, nz, dx, dy,
                                                                 dz, (double*)d_uo, (double*)d_vo, (double*)d_wo);
  else { set_error("ptv_projection_correct: bad dtype"); return PTV_ERR_INVALID; }
  count_launches(1);
  PTV_CUDA(cudaGetLastError());
  return PTV_OK;
}
