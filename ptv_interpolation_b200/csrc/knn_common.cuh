// Declarations shared by the two fused kNN kernels (heap: knn_interp.cu, streaming: knn_stream.cu).
#pragma once
#include <math.h>

#include "ptv_internal.cuh"

namespace ptv {

static constexpr int kStageCap = 256;  // particle records per staging chunk (8 KB)

struct KnnParams {
  HashGrid g;
  const double* ax;
  const double* ay;
  const double* az;
  int nx, ny, nz;
  const uint8_t* mask;
  int method;
  int k;
  double power;
  void* u;
  void* v;
  void* w;
  int64_t* knn_idx;
  double* knn_dist;
  int tiles_x, tiles_y, tiles_z;
  int r0;
  double smoothing;
  int* err_flag;
  // stream kernel -> heap kernel hand-off: tiles the optimistic kernel could not finish
  int* fail_list;        // [capacity]
  int* fail_count;       // [1]
  const int* tile_list;  // heap kernel: process only these tiles (NULL = all tiles of the grid)
  const int* tile_count;
  unsigned long long* stats;  // optional counters (tiles, failed tiles, candidates, accepted)
};

__device__ __forceinline__ bool key_greater(double ka, int ia, double kb, int ib) {
  return ka > kb || (ka == kb && ia > ib);
}

template <int T>
__device__ __forceinline__ int block_scan_excl(int v, int* warp_tot, int* total) {
  constexpr int NW = T / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int w2 = 0; w2 < NW; ++w2) {
    const int x = warp_tot[w2];
    if (w2 < wid) woff += x;
    tot += x;
  }
  __syncthreads();
  *total = tot;
  return woff + inc - v;
}

__device__ __forceinline__ int cell_of(double p, double o, double inv_cell, int n) {
  const int c = (int)floor((p - o) * inv_cell);
  return min(max(c, 0), n - 1);
}

template <typename OutT>
__device__ __forceinline__ void store_out(void* base, int64_t i, double v) {
  reinterpret_cast<OutT*>(base)[i] = (OutT)v;
}



// launchers defined next to their kernels
int launch_knn_heap(KnnParams& p, int T, bool f32, cudaStream_t stream);
int launch_knn_stream(KnnParams& p, int T, bool f32, cudaStream_t stream);
size_t knn_heap_smem_bytes(int T, int k, int method);

}  // namespace ptv
