// Device side of the point-cloud preparation that precedes every optimise call in the host scripts:
// wild_completion/utils.py:408-419 clean_pcd (open3d DBSCAN, keep the most frequent cluster) and :422-459 get_pose_init
// (axis-aligned bounding box, crop of the background cloud, mean offset -> initial yaw).
//
// DBSCAN is restated from open3d's PointCloud::ClusterDBSCAN (third-party, open3d==0.17, not vendored): neighbourhoods are
// {q : |p - q|^2 < eps^2} INCLUDING p itself (nanoflann radius search with the squared radius), a point is a core point when its
// neighbourhood has >= min_points members, clusters are numbered in the order in which the sequential scan over the point
// indices meets their first core point, a border point takes the label of the FIRST cluster whose expansion reaches it and the
// rest is noise (-1).  That outcome does not depend on the expansion order: core points -> connected component of the core
// graph, numbered by the component's smallest core index; border point -> smallest cluster number among its core neighbours.
// So it parallelises exactly: integer work, bit-identical to the sequential algorithm (oracle/preprocess_oracle.py).
// Sizes are a few thousand points per fruit (cfg opt.recon.n_pts = 2000): all-pairs distance tests from shared-memory tiles.
#include <float.h>

#include "common.cuh"

namespace {

constexpr int kTile = 256;

__device__ __forceinline__ double dist2(double ax, double ay, double az, double bx, double by, double bz) {
  const double dx = ax - bx, dy = ay - by, dz = az - bz;
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));      // Eigen's squaredNorm order, no contraction
}

// pass 1: neighbour counts -> core flags; comp[i] = i for core points, INT32_MAX otherwise
__global__ void __launch_bounds__(kTile) dbscan_core_kernel(const double* __restrict__ pts, int n, double eps2, int min_points,
                                                            int32_t* __restrict__ comp, uint8_t* __restrict__ core) {
  __shared__ double sx[kTile], sy[kTile], sz[kTile];
  const int i = blockIdx.x * kTile + threadIdx.x;
  const double x = i < n ? pts[3 * (size_t)i] : 0, y = i < n ? pts[3 * (size_t)i + 1] : 0, z = i < n ? pts[3 * (size_t)i + 2] : 0;
  int cnt = 0;
  for (int j0 = 0; j0 < n; j0 += kTile) {
    const int j = j0 + threadIdx.x;
    if (j < n) { sx[threadIdx.x] = pts[3 * (size_t)j]; sy[threadIdx.x] = pts[3 * (size_t)j + 1]; sz[threadIdx.x] = pts[3 * (size_t)j + 2]; }
    __syncthreads();
    const int m = min(kTile, n - j0);
    for (int k = 0; k < m; ++k) cnt += dist2(x, y, z, sx[k], sy[k], sz[k]) < eps2 ? 1 : 0;
    __syncthreads();
  }
  if (i < n) {
    const bool c = cnt >= min_points;
    core[i] = c ? 1 : 0;
    comp[i] = c ? i : INT32_MAX;
  }
}

// pass 2 (iterated): every core point takes the smallest component id among its core neighbours (and follows one pointer)
__global__ void __launch_bounds__(kTile) dbscan_propagate_kernel(const double* __restrict__ pts, int n, double eps2, const uint8_t* __restrict__ core,
                                                                 const int32_t* __restrict__ comp_in, int32_t* __restrict__ comp_out,
                                                                 int32_t* __restrict__ changed) {
  __shared__ double sx[kTile], sy[kTile], sz[kTile];
  __shared__ int32_t sc[kTile];
  const int i = blockIdx.x * kTile + threadIdx.x;
  const double x = i < n ? pts[3 * (size_t)i] : 0, y = i < n ? pts[3 * (size_t)i + 1] : 0, z = i < n ? pts[3 * (size_t)i + 2] : 0;
  const bool me = i < n && core[i];
  int32_t best = me ? comp_in[i] : INT32_MAX;
  for (int j0 = 0; j0 < n; j0 += kTile) {
    const int j = j0 + threadIdx.x;
    if (j < n) {
      sx[threadIdx.x] = pts[3 * (size_t)j]; sy[threadIdx.x] = pts[3 * (size_t)j + 1]; sz[threadIdx.x] = pts[3 * (size_t)j + 2];
      sc[threadIdx.x] = core[j] ? comp_in[j] : INT32_MAX;
    }
    __syncthreads();
    if (me) {
      const int m = min(kTile, n - j0);
      for (int k = 0; k < m; ++k)
        if (sc[k] < best && dist2(x, y, z, sx[k], sy[k], sz[k]) < eps2) best = sc[k];
    }
    __syncthreads();
  }
  if (i < n) {
    if (me) {
      const int32_t hop = comp_in[best];             // pointer jumping: the root of my best neighbour's current tree
      if (hop < best) best = hop;
      if (best != comp_in[i]) *changed = 1;
    }
    comp_out[i] = best;
  }
}

// pass 3: cluster numbers = rank of the component roots in index order (exclusive scan of the root flags, one block)
__global__ void __launch_bounds__(1024) dbscan_rank_kernel(int n, const uint8_t* __restrict__ core, const int32_t* __restrict__ comp,
                                                           int32_t* __restrict__ rank, int32_t* __restrict__ n_clusters) {
  __shared__ int32_t s[1024];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < n && core[i] && comp[i] == i) ? 1 : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      const int t = (threadIdx.x >= off) ? s[threadIdx.x - off] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) rank[i] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_clusters = carry;
}

// pass 4: labels.  core -> number of its component; border -> smallest number among its core neighbours; noise -> -1
__global__ void __launch_bounds__(kTile) dbscan_label_kernel(const double* __restrict__ pts, int n, double eps2, const uint8_t* __restrict__ core,
                                                             const int32_t* __restrict__ comp, const int32_t* __restrict__ rank,
                                                             int32_t* __restrict__ labels) {
  __shared__ double sx[kTile], sy[kTile], sz[kTile];
  __shared__ int32_t sl[kTile];
  const int i = blockIdx.x * kTile + threadIdx.x;
  const double x = i < n ? pts[3 * (size_t)i] : 0, y = i < n ? pts[3 * (size_t)i + 1] : 0, z = i < n ? pts[3 * (size_t)i + 2] : 0;
  const bool me_core = i < n && core[i];
  int32_t best = INT32_MAX;
  for (int j0 = 0; j0 < n; j0 += kTile) {
    const int j = j0 + threadIdx.x;
    if (j < n) {
      sx[threadIdx.x] = pts[3 * (size_t)j]; sy[threadIdx.x] = pts[3 * (size_t)j + 1]; sz[threadIdx.x] = pts[3 * (size_t)j + 2];
      sl[threadIdx.x] = core[j] ? rank[comp[j]] : INT32_MAX;
    }
    __syncthreads();
    if (i < n && !me_core) {
      const int m = min(kTile, n - j0);
      for (int k = 0; k < m; ++k)
        if (sl[k] < best && dist2(x, y, z, sx[k], sy[k], sz[k]) < eps2) best = sl[k];
    }
    __syncthreads();
  }
  if (i < n) labels[i] = me_core ? rank[comp[i]] : (best == INT32_MAX ? -1 : best);
}

// axis-aligned bounds of a cloud / count and coordinate sums of the points inside a box (inclusive bounds, like open3d's
// AxisAlignedBoundingBox crop): per-block partials in a fixed order, the host adds them in block order (deterministic)
__global__ void __launch_bounds__(256) bbox_kernel(const double* __restrict__ pts, int64_t n, double* __restrict__ part /* [blocks][6] */) {
  __shared__ double s[6][256];
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    for (int c = 0; c < 3; ++c) { const double v = pts[3 * i + c]; mn[c] = fmin(mn[c], v); mx[c] = fmax(mx[c], v); }
  for (int c = 0; c < 3; ++c) { s[c][threadIdx.x] = mn[c]; s[3 + c][threadIdx.x] = mx[c]; }
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off)
      for (int c = 0; c < 3; ++c) {
        s[c][threadIdx.x] = fmin(s[c][threadIdx.x], s[c][threadIdx.x + off]);
        s[3 + c][threadIdx.x] = fmax(s[3 + c][threadIdx.x], s[3 + c][threadIdx.x + off]);
      }
    __syncthreads();
  }
  if (threadIdx.x < 6) part[blockIdx.x * 6 + threadIdx.x] = s[threadIdx.x][0];
}

__global__ void __launch_bounds__(256) crop_sum_kernel(const double* __restrict__ pts, int64_t n, double x0, double y0, double z0, double x1, double y1,
                                                       double z1, double cx, double cy, double cz, double* __restrict__ part /* [blocks][4] */) {
  __shared__ double s[4][256];
  double a[4] = {0, 0, 0, 0};
  // contiguous chunk per block and per thread, so the summation order is a fixed function of (n, grid)
  const int64_t per_block = (n + gridDim.x - 1) / gridDim.x;
  const int64_t b0 = (int64_t)blockIdx.x * per_block, b1 = min(n, b0 + per_block);
  for (int64_t i = b0 + threadIdx.x; i < b1; i += 256) {
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    if (x >= x0 && x <= x1 && y >= y0 && y <= y1 && z >= z0 && z <= z1) { a[0] += 1.0; a[1] += x - cx; a[2] += y - cy; a[3] += z - cz; }
  }
  for (int c = 0; c < 4; ++c) s[c][threadIdx.x] = a[c];
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off)
      for (int c = 0; c < 4; ++c) s[c][threadIdx.x] += s[c][threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x < 4) part[blockIdx.x * 4 + threadIdx.x] = s[threadIdx.x][0];
}

int use_device(hm_context* ctx) {
  if (ctx) HM_CUDA(cudaSetDevice(ctx->device));
  return HM_OK;
}

}  // namespace

extern "C" int hm_dbscan(hm_context* ctx, const double* d_points, int64_t n, double eps, int32_t min_points, int32_t* d_labels,
                         int32_t* h_n_clusters, void* stream) {
  HM_CHECK(n >= 0 && n < ((int64_t)1 << 24) && eps >= 0 && (n == 0 || (d_points && d_labels)), "hm_dbscan: bad argument");
  if (h_n_clusters) *h_n_clusters = 0;
  if (n == 0) return HM_OK;
  int rc = use_device(ctx);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int nn = (int)n;
  int32_t *comp_a = nullptr, *comp_b = nullptr, *rank = nullptr, *flags = nullptr;
  uint8_t* core = nullptr;
  HM_CUDA(cudaMalloc(&comp_a, sizeof(int32_t) * (3 * n + 4) + n));
  comp_b = comp_a + n; rank = comp_b + n; flags = rank + n; core = reinterpret_cast<uint8_t*>(flags + 4);
  const double eps2 = eps * eps;
  const unsigned blocks = (unsigned)((n + kTile - 1) / kTile);
  auto fail = [&](cudaError_t e) { cudaFree(comp_a); hm_set_error("hm_dbscan: %s", cudaGetErrorString(e)); return HM_ERR_CUDA; };
  dbscan_core_kernel<<<blocks, kTile, 0, st>>>(d_points, nn, eps2, min_points, comp_a, core);
  int32_t *cur = comp_a, *nxt = comp_b;
  for (int it = 0; it < nn + 1; ++it) {                 // converges in O(log diameter) rounds; the bound is a safety net
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return fail(e);
    dbscan_propagate_kernel<<<blocks, kTile, 0, st>>>(d_points, nn, eps2, core, cur, nxt, flags);
    std::swap(cur, nxt);
    int32_t changed = 0;
    e = cudaMemcpyAsync(&changed, flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(e);
    if (!changed) break;
  }
  dbscan_rank_kernel<<<1, 1024, 0, st>>>(nn, core, cur, rank, flags + 1);
  dbscan_label_kernel<<<blocks, kTile, 0, st>>>(d_points, nn, eps2, core, cur, rank, d_labels);
  int32_t nc = 0;
  cudaError_t e = cudaMemcpyAsync(&nc, flags + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(e);
  cudaFree(comp_a);
  if (h_n_clusters) *h_n_clusters = nc;
  if (ctx) ctx->counters.kernel_launches += 3;
  return HM_OK;
}

extern "C" int hm_cloud_bounds(hm_context* ctx, const double* d_points, int64_t n, double* h_min3, double* h_max3, void* stream) {
  HM_CHECK(d_points && h_min3 && h_max3 && n > 0, "hm_cloud_bounds: bad argument");
  int rc = use_device(ctx);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>(296, (n + 255) / 256);
  double* d_part = nullptr;
  HM_CUDA(cudaMalloc(&d_part, sizeof(double) * 6 * blocks));
  bbox_kernel<<<blocks, 256, 0, st>>>(d_points, n, d_part);
  std::vector<double> h(6 * (size_t)blocks);
  cudaError_t e = cudaMemcpyAsync(h.data(), d_part, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_part);
  HM_CUDA(e);
  for (int c = 0; c < 3; ++c) { h_min3[c] = DBL_MAX; h_max3[c] = -DBL_MAX; }
  for (int b = 0; b < blocks; ++b)
    for (int c = 0; c < 3; ++c) { h_min3[c] = std::min(h_min3[c], h[6 * b + c]); h_max3[c] = std::max(h_max3[c], h[6 * b + 3 + c]); }
  return HM_OK;
}

extern "C" int hm_crop_mean_offset(hm_context* ctx, const double* d_points, int64_t n, const double* h_box_min3, const double* h_box_max3,
                                   const double* h_center3, int64_t* h_count, double* h_mean3, void* stream) {
  HM_CHECK(d_points && h_box_min3 && h_box_max3 && h_center3 && h_count && h_mean3 && n >= 0, "hm_crop_mean_offset: bad argument");
  *h_count = 0;
  h_mean3[0] = h_mean3[1] = h_mean3[2] = 0.0;
  if (n == 0) return HM_OK;
  int rc = use_device(ctx);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>(296, (n + 255) / 256);
  double* d_part = nullptr;
  HM_CUDA(cudaMalloc(&d_part, sizeof(double) * 4 * blocks));
  crop_sum_kernel<<<blocks, 256, 0, st>>>(d_points, n, h_box_min3[0], h_box_min3[1], h_box_min3[2], h_box_max3[0], h_box_max3[1], h_box_max3[2],
                                          h_center3[0], h_center3[1], h_center3[2], d_part);
  std::vector<double> h(4 * (size_t)blocks);
  cudaError_t e = cudaMemcpyAsync(h.data(), d_part, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_part);
  HM_CUDA(e);
  double a[4] = {0, 0, 0, 0};
  for (int b = 0; b < blocks; ++b)
    for (int c = 0; c < 4; ++c) a[c] += h[4 * b + c];
  *h_count = (int64_t)a[0];
  if (a[0] > 0)
    for (int c = 0; c < 3; ++c) h_mean3[c] = a[1 + c] / a[0];
  return HM_OK;
}
