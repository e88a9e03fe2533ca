/*
 * ptv_b200.h -- C ABI of the B200-native PTV scattered-to-grid interpolation path.
 *
 * The reference (tombultreys/ptv_interpolation) is pure Python and has no FFI: its boundary
 * for this path is the module-level API of interpolator.py / physics.py.  Each entry point
 * below names the reference interface it replaces (file:line in the upstream tree); the
 * ctypes binding a maintainer would add is shown in INTEGRATION.md and shipped in
 * ptv_interpolation_b200/_cabi.py.
 *
 * Conventions
 *   - Every function returns a status (PTV_OK == 0).  On failure ptv_last_error() returns a
 *     thread-local, NUL-terminated message.  The Python shim maps PTV_ERR_INVALID ->
 *     ValueError, PTV_ERR_TOO_FEW -> IndexError (what interpolator.py:150 raises when
 *     Np < k), PTV_ERR_SINGULAR -> numpy.linalg.LinAlgError (scipy _rbfinterp_np.py:74-87),
 *     anything else -> RuntimeError.
 *   - Pointers named d_* are DEVICE pointers (sm_100a, current device); h_* are HOST
 *     pointers.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Grids are (nz, ny, nx) C-ordered, x fastest, exactly as interpolator.py:59 builds them.
 *     A mask byte != 0 means fluid/pore (interpolator.py:31-37).
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     PTV_ERR_CUDA.
 */
#ifndef PTV_B200_H
#define PTV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTV_OK 0
#define PTV_ERR_INVALID 1
#define PTV_ERR_TOO_FEW 2
#define PTV_ERR_CUDA 3
#define PTV_ERR_SINGULAR 4
#define PTV_ERR_NOMEM 5
#define PTV_ERR_QHULL 6 /* method='linear': what scipy.spatial.QhullError reports (too few points) */

/* interpolation methods (interpolator.py:65 `method=`) */
#define PTV_METHOD_IDW 0     /* interpolator.py:126-155 */
#define PTV_METHOD_SIBSON 1  /* interpolator.py:83-124  */
#define PTV_METHOD_NEAREST 2 /* interpolator.py:197 griddata(method='nearest') == k=1 */
#define PTV_METHOD_RBF 3     /* interpolator.py:157-195 local thin-plate-spline RBF */
#define PTV_METHOD_MADFILTER 4 /* internal: filtering.py:5-58, reached through ptv_outlier_filter */
/* local RBF with the other scale-invariant kernels RBFInterpolator accepts without epsilon
 * (interpolator.py:162-167 passes rbf_kernel through): minimum polynomial degree 1 / 0 / 2 */
#define PTV_METHOD_RBF_CUBIC 5
#define PTV_METHOD_RBF_LINEAR 6
#define PTV_METHOD_RBF_QUINTIC 7
/* interpolator.py:197 griddata(method='linear', fill_value=0.0): barycentric weights in the Delaunay
 * tetrahedron that holds the voxel, 0 outside the convex hull.  k is ignored; the optional lists hold 4
 * entries per voxel: the tetrahedron's vertex rows in ascending order (-1 outside the hull) in knn_idx and
 * their barycentric weights in knn_dist. */
#define PTV_METHOD_LINEAR 8

/* output element types for the velocity grids */
#define PTV_F32 0
#define PTV_F64 1

typedef struct ptv_hash ptv_hash; /* opaque: uniform-grid cell list of the particle cloud */

/* ---- library ------------------------------------------------------------------------- */
int ptv_version(void);
const char* ptv_last_error(void);
/* sm count, compute capability and memory of `device`; fails with PTV_ERR_CUDA if absent. */
int ptv_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
/* Tuning knobs for experiments ("ppc" particles per cell, "r0" rings merged into the first
 * staging batch, "tile" / "stream_tile" 32|64|128 threads per voxel tile of the heap / streaming
 * kernel, "stream" 0|1 use the streaming kernel, "stats" 0|1, "rscale" first-radius factor of the
 * streaming kernel; method='linear': "hull" 2|1|0 hull-candidate list with / without the particle-level
 * stage / none (scan all particles), "linear_k" the first candidate radius holds this many particles,
 * "linear_occ" 3|4 CTAs per SM; "stencil_bulk" 1|0 stencil kernels on bulk-async (TMA) shared-memory rings /
 * direct loads, "stencil_la" rows or planes requested ahead in those rings (0 = default); "rbf_regs" 1|0
 * register-resident / shared-memory local-RBF solve).  Results never depend on them.  Unknown key ->
 * PTV_ERR_INVALID. */
int ptv_set_tuning(const char* key, double value);
/* Number of CUDA kernels this library has launched in this process so far (bench.py's
 * gpu_launches is the difference across the timed region). */
int64_t ptv_launch_count(void);
double ptv_get_tuning(const char* key);

/* ---- spatial hash: replaces KDTree(points) at interpolator.py:90,132 and the tree that
 *      RBFInterpolator builds (scipy _rbfinterp.py:338) -------------------------------- */
int ptv_hash_create(ptv_hash** out);
int ptv_hash_destroy(ptv_hash* h);
/* d_points: (n,3) float64 rows (x,y,z) == df[['x','y','z']].values (interpolator.py:78);
 * d_values: (n,3) float64 rows (u,v,w) == df[['u','v','w']].values (interpolator.py:79).
 * cell_size <= 0 picks the cell edge from the particle density.  Buffers owned by the
 * handle are reused across builds (time-resolved sweeps rebuild per frame).            */
/* Slab hash for the z-slab sharding of SURVEY.md 8(e) (no counterpart in the reference, whose only parallel
 * split is interpolator.py:176-182): the same build, but only particles with
 *   z_lo - halo <= z <= z_hi + halo,  halo = halo_factor x (3 k / (4 pi rho))^(1/3),  rho = n / bbox volume
 * are binned -- a rank that interpolates the planes [z_lo, z_hi] bins its slab's particles plus a halo instead
 * of the whole cloud.  Searches stay exact: ptv_knn_interp counts every voxel whose k-th distance reaches
 * outside the binned range (ptv_hash_clip_violations, synchronises); if the count is not zero the caller
 * repeats the frame on the full hash.  method = PTV_METHOD_LINEAR is refused on a slab hash. */
int ptv_hash_build_slab(ptv_hash* h, const double* d_points, const double* d_values, int64_t n, double cell_size,
                        double z_lo, double z_hi, int k, double halo_factor, void* stream);
int ptv_hash_clip_violations(const ptv_hash* h, int64_t* count);
/* The same count written to a device double on the stream (no synchronisation): lets the caller fold it into
 * the all-reduce of the flux / divergence statistics and decide about the fall-back collectively. */
int ptv_hash_clip_violations_to(const ptv_hash* h, double* d_dst, void* stream);
int ptv_hash_build(ptv_hash* h, const double* d_points, const double* d_values, int64_t n,
                   double cell_size, void* stream);
int ptv_hash_info(const ptv_hash* h, int64_t* n, int dims[3], double origin[3], double* cell_size,
                  int* max_cell_count);

/* ---- fused kNN + weights + accumulate + mask: replaces tree.query + the NumPy weighting at
 *      interpolator.py:97-122 / 139-153 and the solid zeroing at main.py:195-207 --------
 * Queries are the rectilinear grid ax_x (x) ax_y (x) ax_z given by its three float64 axes
 * (what create_grid returns as (x, y, z), interpolator.py:54-56).  d_mask may be NULL (all
 * voxels interpolated, like the reference); with a mask, voxels whose byte is 0 are written
 * as 0 and skipped.  d_knn_idx / d_knn_dist (nullable, (nvox,k), ascending by (d2, index))
 * expose the neighbour lists for the bit-exact parity tests; solid voxels get -1 / NaN.  */
int ptv_knn_interp(const ptv_hash* h, const double* d_ax_x, int nx, const double* d_ax_y, int ny,
                   const double* d_ax_z, int nz, const uint8_t* d_mask, int method, int k,
                   double idw_power, double rbf_smoothing, int out_dtype, void* d_u, void* d_v,
                   void* d_w, int64_t* d_knn_idx, double* d_knn_dist, void* stream);

/* ---- the same search and weights for ARBITRARY query points (non-rectilinear grid_tuple, or the
 *      particles themselves).  `queries` is a second hash built over the query points (its values are
 *      ignored; pass h itself for a self-query): its cell order keeps tiles of 128 queries compact.
 *      Outputs are indexed by the query's original row: d_u/d_v/d_w [nq] (nullable together),
 *      d_knn_idx/d_knn_dist [nq][k] (nullable together). ------------------------------------------ */
int ptv_knn_points(const ptv_hash* h, const ptv_hash* queries, int method, int k, double idw_power,
                   double rbf_smoothing, int out_dtype, void* d_u, void* d_v, void* d_w, int64_t* d_knn_idx,
                   double* d_knn_dist, void* stream);

/* ---- kNN median/MAD outlier filter: replaces remove_outliers_knn (filtering.py:5-58).  Self-query
 *      with k+1 neighbours, first neighbour dropped, z = |speed - median| / (MAD + 1e-6);
 *      d_keep[i] = z <= threshold, d_kth_dist[i] = distance to the k-th neighbour (filtering.py:33). */
int ptv_outlier_filter(const ptv_hash* h, int k, double threshold, uint8_t* d_keep, double* d_kth_dist,
                       void* stream);

/* Diagnostics of the last ptv_knn_interp call on this handle: whether the streaming kernel ran,
 * how many voxel tiles it handed to the exact heap kernel, and (with tuning "stats" = 1) how many it
 * finished itself.  Synchronises the device. */
int ptv_knn_stats(const ptv_hash* h, int64_t* used_stream, int64_t* tiles_failed, int64_t* tiles_streamed);
/* With tuning "stats" = 1: why tiles were handed over -- [0] no local density estimate, [1] k-th
 * neighbour beyond the histogram range, [2] crossing bin larger than the short list, [3] exact
 * verification failed. */
int ptv_knn_fail_reasons(const ptv_hash* h, int64_t reasons[4]);
/* With tuning "stats" = 1: work done by the streaming kernel in the last call -- [0] voxel-candidate pairs
 * of the histogram passes, [1] pairs of the classification passes (float32 pre-test), [2] exact float64
 * keys evaluated, [3] crossing-bin list entries, [4] pore voxels finished, [5] rounds (warp passes),
 * [6] histogram points, [7] retries.  The reference has no counterpart (its search is
 * interpolator.py:139, one cKDTree.query). */
int ptv_knn_work_stats(const ptv_hash* h, int64_t work[8]);
/* With tuning "stats" = 1, after a PTV_METHOD_LINEAR call: [0] voxels finished on the warp-shared candidate
 * set, [1] of those, voxels that re-used the previous voxel's tetrahedron, [2] voxels solved on their own
 * (growing region), [3] candidate sets read from global memory, [4] voxels outside the convex hull,
 * [5] unresolved voxels (pivot limit; written as 0), [6] pivots, [7] size of the hull-candidate list. */
int ptv_linear_stats(const ptv_hash* h, int64_t stats[8]);

/* ---- mask resampling: replaces sample_mask_on_grid (interpolator.py:205-238).  The
 *      per-axis nearest index maps (-1 == out of bounds) are computed by the host shim with
 *      the RegularGridInterpolator rule (scipy _rgi.py:551-554); this is the gather.     */
int ptv_mask_gather(const uint8_t* d_mask_raw, int rnx, int rny, int rnz, const int32_t* d_ix,
                    int nx, const int32_t* d_iy, int ny, const int32_t* d_iz, int nz,
                    uint8_t* d_out, void* stream);

/* ---- no-slip wall particles: replaces extract_boundary_particles (interpolator.py:240-284).
 *      Flags solid voxels within `thickness` 6-connected dilation steps of fluid, compacts
 *      their linear indices in C order into d_indices (capacity `cap`), total in *h_count. */
/* The same with a caller-owned workspace of ptv_boundary_workspace_bytes() bytes, in two phases so that the
 * caller can size the index array: phase 0 packs the mask into bits, dilates, flags and counts (-> *h_count,
 * synchronises); phase 1 writes the indices from the flag words phase 0 left in the workspace.  thickness < 1
 * behaves like binary_dilation(iterations < 1): dilate until nothing changes. */
int64_t ptv_boundary_workspace_bytes(int nx, int ny, int nz);
int ptv_boundary_voxels_ws(const uint8_t* d_mask, int nx, int ny, int nz, int thickness, void* d_work, int phase,
                           int64_t* d_indices, int64_t cap, int64_t* h_count, void* stream);
int ptv_boundary_voxels(const uint8_t* d_mask, int nx, int ny, int nz, int thickness,
                        int64_t* d_indices, int64_t cap, int64_t* h_count, void* stream);

/* ---- solid zeroing alone (main.py:202-207) for fields produced elsewhere ---------------- */
int ptv_apply_mask(void* d_u, void* d_v, void* d_w, const uint8_t* d_mask, int64_t nvox, int dtype,
                   void* stream);

/* ---- masked finite-volume divergence: replaces compute_consistent_divergence
 *      (physics.py:6-53).  Slab form: planes [0,nz) of the local slab; d_w_below/d_w_above and
 *      d_mask_above are the (ny,nx) halo planes from the z-neighbours, NULL at the domain
 *      edge (Neumann, physics.py:38-45).  If d_absdiv_sum != NULL also accumulates
 *      sum(|div|) over fluid voxels and their count into d_absdiv_sum[0..1] (float64,
 *      physics.py:174).  dtype selects f32/f64 fields.                                    */
int ptv_divergence(const void* d_u, const void* d_v, const void* d_w, const uint8_t* d_mask, int nx,
                   int ny, int nz, double dx, double dy, double dz, const void* d_w_below,
                   const void* d_w_above, const uint8_t* d_mask_above, int dtype, void* d_div,
                   double* d_absdiv_sum, void* stream);

/* ---- plane fluxes: replaces calculate_flux_xy/xz/yz (plot_flux.py:6-16) and the mid-plane
 *      X flux (physics.py:160-165).  Outputs are float64 sums WITHOUT the dx*dy area factor
 *      (the shim multiplies, as the reference does after np.sum): d_qxy[nz], d_qxz[ny],
 *      d_qyz[nx] must be zeroed by the caller (the kernel accumulates, so slabs can share). */
int ptv_flux_profiles(const void* d_u, const void* d_v, const void* d_w, int nx, int ny, int nz,
                      int dtype, double* d_qxy, double* d_qxz, double* d_qyz, void* stream);

/* ---- the two above in ONE pass over the fields (what the pipeline uses): divergence, sum|div| and
 *      fluid count, and the three unscaled flux profiles.  d_qxy/d_qxz/d_qyz: all three or all
 *      NULL; they and d_absdiv_sum are accumulated into (zero them first). -------------------- */
int ptv_divergence_flux(const void* d_u, const void* d_v, const void* d_w, const uint8_t* d_mask, int nx,
                        int ny, int nz, double dx, double dy, double dz, const void* d_w_below,
                        const void* d_w_above, const uint8_t* d_mask_above, int dtype, void* d_div,
                        double* d_absdiv_sum, double* d_qxy, double* d_qxz, double* d_qyz, void* stream);

/* ---- self-test of the stencil kernels' division: the kernels divide by the grid spacings through a
 *      correctly-rounded reciprocal sequence instead of the IEEE division instruction sequence (bit-identity
 *      with physics.py:26-53 and np.gradient depends on it).  Runs n pseudo-random numerators against
 *      divisor h on the device and returns how many results differ from the IEEE quotient (must be 0). */
int ptv_selftest_division(double h, int64_t n, uint64_t seed, int64_t* mismatches);

/* ---- shear-rate magnitude and vorticity magnitude: replaces compute_strain_rate / compute_vorticity
 *      (velocity_analysis.py:10-63, 94-120; nine np.gradient stencils).  d_mask nullable; either output
 *      nullable. ------------------------------------------------------------------------------- */
int ptv_strain_vorticity(const void* d_u, const void* d_v, const void* d_w, const uint8_t* d_mask, int nx,
                         int ny, int nz, double dx, double dy, double dz, int dtype, void* d_strain,
                         void* d_vorticity, void* stream);

/* z-slab form for a grid sharded over GPUs: d_below / d_above are the (3, ny, nx) planes (u, v, w) of the
 * z-neighbours just outside the slab (NULL = that side is a face of the whole domain), so the slab's outer
 * planes get np.gradient's central differences.  float32 fields with nx % 16 == 0 and 16-byte aligned buffers
 * only (PTV_ERR_INVALID otherwise: pad the slab with its halo planes and call ptv_strain_vorticity). */
int ptv_strain_vorticity_slab(const void* d_u, const void* d_v, const void* d_w, const uint8_t* d_mask, int nx,
                              int ny, int nz, double dx, double dy, double dz, const void* d_below,
                              const void* d_above, int dtype, void* d_strain, void* d_vorticity, void* stream);

/* ---- projection cleaning (physics.py:55-209).  ptv_poisson_lsqr solves
 *      A phi = div - mean(div[mask]) with the matrix-free masked 7-point Laplacian of
 *      build_laplacian_matrix (physics.py:55-108) by LSQR with SciPy's recurrences and stopping rules
 *      (physics.py:186: damp=1e-8, atol=btol=1e-10, iter_lim=3000; conlim SciPy default 1e8).
 *      d_phi: float64 (nz,ny,nx), zero in solid voxels; d_work: ptv_poisson_workspace_bytes() bytes;
 *      h_info[8] (host): istop, itn, r1norm, r2norm, anorm, acond, arnorm, xnorm.
 *      ptv_projection_correct applies apply_consistent_correction (physics.py:110-147). ---------- */
int64_t ptv_poisson_workspace_bytes(int nx, int ny, int nz);
int ptv_poisson_lsqr(const void* d_div, int dtype, const uint8_t* d_mask, int nx, int ny, int nz, double dx,
                     double dy, double dz, double damp, double atol, double btol, double conlim, int iter_lim,
                     double* d_phi, void* d_work, double* h_info, void* stream);
int ptv_projection_correct(const void* d_u, const void* d_v, const void* d_w, const double* d_phi,
                           const uint8_t* d_mask, int nx, int ny, int nz, double dx, double dy, double dz,
                           int dtype, void* d_uo, void* d_vo, void* d_wo, void* stream);

/* ---- host-buffer convenience (what a non-CUDA caller binds): copies in, runs
 *      ptv_hash_build + ptv_knn_interp, copies out.  All pointers are HOST pointers. ------ */
int ptv_interpolate_host(const double* h_points, const double* h_values, int64_t n,
                         const double* h_ax_x, int nx, const double* h_ax_y, int ny,
                         const double* h_ax_z, int nz, const uint8_t* h_mask, int method, int k,
                         double idw_power, double rbf_smoothing, int out_dtype, void* h_u, void* h_v,
                         void* h_w);

#ifdef __cplusplus
}
#endif
#endif /* PTV_B200_H */
