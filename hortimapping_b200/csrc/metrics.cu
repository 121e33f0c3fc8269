// Nearest-neighbour distances between point clouds on the device: the arithmetic under the reference's
// evaluation metrics, metrics_3d/chamfer_distance.py:16-26 and metrics_3d/precision_recall.py:19-50, which call open3d's
// PointCloud.compute_point_cloud_distance (exact nearest neighbour, double precision) on ~1e5-point clouds per fruit.
//
// Brute force in fp64: every thread owns one query point, target points stream through shared memory in tiles.  The
// result is the exact minimum (no approximation, no atomics), so it agrees with any exact NN search to fp64 rounding.
#include "common.cuh"

namespace {

constexpr int kQ = 256;        // queries per block (one per thread)
constexpr int kTile = 1024;    // target points per shared-memory tile (24 KB)

__global__ void __launch_bounds__(kQ) nn_distance_kernel(const double* __restrict__ q, int64_t nq, const double* __restrict__ t, int64_t nt,
                                                         double* __restrict__ dist) {
  __shared__ double sx[kTile], sy[kTile], sz[kTile];
  const int64_t i = (int64_t)blockIdx.x * kQ + threadIdx.x;
  const bool live = i < nq;
  const double qx = live ? q[i * 3 + 0] : 0.0, qy = live ? q[i * 3 + 1] : 0.0, qz = live ? q[i * 3 + 2] : 0.0;
  double best = INFINITY;
  for (int64_t base = 0; base < nt; base += kTile) {
    const int m = (int)min((int64_t)kTile, nt - base);
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += kQ) {
      sx[j] = t[(base + j) * 3 + 0];
      sy[j] = t[(base + j) * 3 + 1];
      sz[j] = t[(base + j) * 3 + 2];
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < m; ++j) {
      const double dx = qx - sx[j], dy = qy - sy[j], dz = qz - sz[j];
      const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      best = fmin(best, d2);
    }
  }
  if (live) dist[i] = sqrt(best);
}

}  // namespace

// dist[i] = min_j || query[i] - target[j] ||_2 (all fp64, row-major [n][3], device pointers); ctx may be NULL.
extern "C" int hm_nn_distance(hm_context* ctx, const double* d_query, int64_t n_query, const double* d_target, int64_t n_target,
                              double* d_dist, void* stream) {
  HM_CHECK(n_query >= 0 && n_target > 0 && (n_query == 0 || (d_query && d_dist)) && d_target, "hm_nn_distance: bad argument");
  if (n_query == 0) return HM_OK;
  if (ctx) HM_CUDA(cudaSetDevice(ctx->device));     // ctx may be NULL (the metrics need no decoder): current device
  nn_distance_kernel<<<(unsigned)((n_query + kQ - 1) / kQ), kQ, 0, (cudaStream_t)stream>>>(d_query, n_query, d_target, n_target, d_dist);
  if (ctx) ctx->counters.kernel_launches += 1;
  HM_CUDA(cudaGetLastError());
  return HM_OK;
}
