// Device iso-surface extraction for the mesher (the step after the hot path).
//
// The reference hands the N^3 SDF grid to skimage.measure.marching_cubes on the host
// (wild_completion/utils.py:565-588; third-party, unpinned -> mesh-level parity is pinned at Chamfer level only,
// SURVEY.md 8c).  This file extracts the zero level set on the GPU so that only vertices and faces leave the device.
// It is the same marching-tetrahedra scheme as the host extractor hortimapping_b200/marching.py (6 tetrahedra per cell
// around the main diagonal, vertices welded per grid edge, faces oriented towards increasing field values) and
// reproduces its output ordering exactly: vertices sorted by (lower end point, upper end point) of their grid edge,
// faces in (cell, tetrahedron, triangle) order -- so the two can be compared index by index in the tests.
//
// HBM-bound integer/byte work: 4 passes over the grid (edge flags -> scan -> cell triangle counts -> scan -> emit),
// coalesced along z, no atomics (deterministic output).
#include <algorithm>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace {

// the 7 edge directions of the tetrahedral decomposition, ordered by their linear index offset (dx*n + dy)*n + dz
__constant__ int c_edge_dir[7][3] = {{0, 0, 1}, {0, 1, 0}, {0, 1, 1}, {1, 0, 0}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1}};
// corner c of a cell = (dx, dy, dz); marching.py _CORNERS
__constant__ int c_corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
// marching.py _TETS
__constant__ int c_tet[6][4] = {{0, 5, 1, 6}, {0, 1, 2, 6}, {0, 2, 3, 6}, {0, 3, 7, 6}, {0, 7, 4, 6}, {0, 4, 5, 6}};
// marching.py _EDGES (local vertex pairs of a tetrahedron)
__constant__ int c_tedge[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
// marching.py _TRI_TABLE: per inside-mask up to 2 triangles of 3 tetra-edge ids (-1 = none)
__constant__ int c_tri[16][2][3] = {
    {{-1, -1, -1}, {-1, -1, -1}}, {{0, 1, 2}, {-1, -1, -1}},  {{0, 3, 4}, {-1, -1, -1}},  {{1, 2, 3}, {2, 3, 4}},
    {{1, 3, 5}, {-1, -1, -1}},    {{0, 2, 3}, {2, 3, 5}},     {{0, 1, 4}, {1, 4, 5}},     {{2, 4, 5}, {-1, -1, -1}},
    {{2, 4, 5}, {-1, -1, -1}},    {{0, 1, 4}, {1, 4, 5}},     {{0, 2, 3}, {2, 3, 5}},     {{1, 3, 5}, {-1, -1, -1}},
    {{1, 2, 3}, {2, 3, 4}},       {{0, 3, 4}, {-1, -1, -1}},  {{0, 1, 2}, {-1, -1, -1}},  {{-1, -1, -1}, {-1, -1, -1}}};

__device__ __forceinline__ int edge_type_of(int dx, int dy, int dz) {
  // inverse of c_edge_dir: (0,0,1)->0 (0,1,0)->1 (0,1,1)->2 (1,0,0)->3 (1,0,1)->4 (1,1,0)->5 (1,1,1)->6
  return dx * 4 + dy * 2 + dz - 1;
}

// pass 1: flag[p*7 + t] = 1 if grid edge (p, p + dir[t]) exists and is crossed by the level set
__global__ void edge_flag_kernel(const float* __restrict__ sdf, int n, double level, int64_t n_edges, int32_t* __restrict__ flag) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_edges) return;
  const int64_t p = idx / 7;
  const int t = (int)(idx - p * 7);
  const int z = (int)(p % n), y = (int)((p / n) % n), x = (int)(p / ((int64_t)n * n));
  const int x2 = x + c_edge_dir[t][0], y2 = y + c_edge_dir[t][1], z2 = z + c_edge_dir[t][2];
  int f = 0;
  if (x2 < n && y2 < n && z2 < n) {
    const double va = (double)sdf[p] - level, vb = (double)sdf[((int64_t)x2 * n + y2) * n + z2] - level;
    f = ((va < 0.0) != (vb < 0.0)) ? 1 : 0;
  }
  flag[idx] = f;
}

__device__ __forceinline__ int tet_mask(const double* cv, int tet) {
  int m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) m |= (cv[c_tet[tet][i]] < 0.0) ? (1 << i) : 0;
  return m;
}

__device__ __forceinline__ void load_cell(const float* __restrict__ sdf, int n, double level, int x, int y, int z, double* cv) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    cv[c] = (double)sdf[((int64_t)(x + c_corner[c][0]) * n + (y + c_corner[c][1])) * n + (z + c_corner[c][2])] - level;
}

// pass 2: number of triangles of every cell
__global__ void cell_count_kernel(const float* __restrict__ sdf, int n, double level, int64_t n_cells, int32_t* __restrict__ count) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int m1 = n - 1;
  const int z = (int)(c % m1), y = (int)((c / m1) % m1), x = (int)(c / ((int64_t)m1 * m1));
  double cv[8];
  load_cell(sdf, n, level, x, y, z, cv);
  int cnt = 0;
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const int m = tet_mask(cv, t);
    cnt += (c_tri[m][0][0] >= 0) + (c_tri[m][1][0] >= 0);
  }
  count[c] = cnt;
}

// pass 3: one vertex per crossed edge, at index = exclusive scan of the flags
__global__ void vertex_kernel(const float* __restrict__ sdf, int n, double level, double spacing, int64_t n_edges,
                              const int32_t* __restrict__ flag, const int32_t* __restrict__ vid, float* __restrict__ verts) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_edges || !flag[idx]) return;
  const int64_t p = idx / 7;
  const int t = (int)(idx - p * 7);
  const int z = (int)(p % n), y = (int)((p / n) % n), x = (int)(p / ((int64_t)n * n));
  const int dx = c_edge_dir[t][0], dy = c_edge_dir[t][1], dz = c_edge_dir[t][2];
  const double va = (double)sdf[p] - level, vb = (double)sdf[((int64_t)(x + dx) * n + (y + dy)) * n + (z + dz)] - level;
  const double tt = va / (va - vb);
  float* o = verts + (int64_t)vid[idx] * 3;
  o[0] = (float)(((double)x + (double)dx * tt) * spacing);
  o[1] = (float)(((double)y + (double)dy * tt) * spacing);
  o[2] = (float)(((double)z + (double)dz * tt) * spacing);
}

// pass 4: faces of every cell, at offset = exclusive scan of the cell counts
__global__ void face_kernel(const float* __restrict__ sdf, int n, double level, int64_t n_cells, const int32_t* __restrict__ cell_off,
                            const int32_t* __restrict__ vid, int32_t* __restrict__ faces) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int m1 = n - 1;
  const int z = (int)(c % m1), y = (int)((c / m1) % m1), x = (int)(c / ((int64_t)m1 * m1));
  double cv[8];
  load_cell(sdf, n, level, x, y, z, cv);
  int64_t out = cell_off[c];
  for (int t = 0; t < 6; ++t) {
    const int m = tet_mask(cv, t);
    if (m == 0 || m == 15) continue;
    // centroids of the outside / inside vertices of the tetrahedron (cell-local coordinates)
    double co[3] = {0, 0, 0}, ci[3] = {0, 0, 0};
    int no = 0, ni = 0;
    for (int i = 0; i < 4; ++i) {
      const int cr = c_tet[t][i];
      if ((m >> i) & 1) { ++ni; for (int k = 0; k < 3; ++k) ci[k] += c_corner[cr][k]; }
      else { ++no; for (int k = 0; k < 3; ++k) co[k] += c_corner[cr][k]; }
    }
    double dirv[3];
    for (int k = 0; k < 3; ++k) dirv[k] = co[k] / no - ci[k] / ni;
    for (int tr = 0; tr < 2; ++tr) {
      if (c_tri[m][tr][0] < 0) continue;
      double pos[3][3];
      int32_t id[3];
      for (int k = 0; k < 3; ++k) {
        const int e = c_tri[m][tr][k];
        const int ca = c_tet[t][c_tedge[e][0]], cb = c_tet[t][c_tedge[e][1]];
        const double va = cv[ca], vb = cv[cb];
        const double tt = va / (va - vb);
        for (int q = 0; q < 3; ++q) pos[k][q] = (double)c_corner[ca][q] + (double)(c_corner[cb][q] - c_corner[ca][q]) * tt;
        // the welded vertex of this grid edge: lower end point + direction type
        const int lo = (c_corner[ca][0] <= c_corner[cb][0] && c_corner[ca][1] <= c_corner[cb][1] && c_corner[ca][2] <= c_corner[cb][2]) ? ca : cb;
        const int hi = (lo == ca) ? cb : ca;
        const int64_t p = ((int64_t)(x + c_corner[lo][0]) * n + (y + c_corner[lo][1])) * n + (z + c_corner[lo][2]);
        id[k] = vid[p * 7 + edge_type_of(c_corner[hi][0] - c_corner[lo][0], c_corner[hi][1] - c_corner[lo][1], c_corner[hi][2] - c_corner[lo][2])];
      }
      const double ax = pos[1][0] - pos[0][0], ay = pos[1][1] - pos[0][1], az = pos[1][2] - pos[0][2];
      const double bx = pos[2][0] - pos[0][0], by = pos[2][1] - pos[0][1], bz = pos[2][2] - pos[0][2];
      const double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
      const bool flip = (nx * dirv[0] + ny * dirv[1] + nz * dirv[2]) < 0.0;
      faces[out * 3 + 0] = id[0];
      faces[out * 3 + 1] = flip ? id[2] : id[1];
      faces[out * 3 + 2] = flip ? id[1] : id[2];
      ++out;
    }
  }
}

// (v + (-1)) * cube_radius in fp64 as wild_completion/utils.py:583-585, back to fp32 (mesher.py:22)
__global__ void vertex_affine_kernel(int64_t n3, double cube_radius, float* __restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  v[i] = (float)(((double)v[i] + -1.0) * cube_radius);
}

}  // namespace

// Extracts the level set of d_sdf [n][n][n] into the context's mesh workspace.  Returns the vertex / face counts (host);
// hm_isosurface_fetch copies them out.  Synchronises the stream.
extern "C" int hm_isosurface(hm_context* ctx, const float* d_sdf, int32_t n, double level, double spacing, int64_t* h_n_verts,
                             int64_t* h_n_faces, void* stream) {
  HM_CHECK(ctx && d_sdf && h_n_verts && h_n_faces && n >= 2 && n <= 640, "hm_isosurface: bad argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_pts = (int64_t)n * n * n, n_edges = n_pts * 7, n_cells = (int64_t)(n - 1) * (n - 1) * (n - 1);
  HM_CHECK(n_edges < (int64_t)1 << 31, "hm_isosurface: grid too large");
  size_t tmp1 = 0, tmp2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp1, (int32_t*)nullptr, (int32_t*)nullptr, (int)n_edges, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp2, (int32_t*)nullptr, (int32_t*)nullptr, (int)n_cells, st);
  const size_t tmp = std::max(tmp1, tmp2);
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t need = al(sizeof(int32_t) * n_edges) * 2 + al(sizeof(int32_t) * n_cells) * 2 + al(tmp);
  if (need > ctx->mesh_ws_bytes) {
    if (ctx->mesh_ws) { HM_CUDA(cudaDeviceSynchronize()); cudaFree(ctx->mesh_ws); ctx->mesh_ws = nullptr; ctx->mesh_ws_bytes = 0; }
    HM_CUDA(cudaMalloc(&ctx->mesh_ws, need));
    ctx->mesh_ws_bytes = need;
  }
  uint8_t* w = (uint8_t*)ctx->mesh_ws;
  int32_t* flag = (int32_t*)w; w += al(sizeof(int32_t) * n_edges);
  int32_t* vid = (int32_t*)w; w += al(sizeof(int32_t) * n_edges);
  int32_t* ccount = (int32_t*)w; w += al(sizeof(int32_t) * n_cells);
  int32_t* coff = (int32_t*)w; w += al(sizeof(int32_t) * n_cells);
  void* d_tmp = w;
  const int B = 256;
  edge_flag_kernel<<<(unsigned)((n_edges + B - 1) / B), B, 0, st>>>(d_sdf, n, level, n_edges, flag);
  size_t tb = tmp;
  cub::DeviceScan::ExclusiveSum(d_tmp, tb, flag, vid, (int)n_edges, st);
  cell_count_kernel<<<(unsigned)((n_cells + B - 1) / B), B, 0, st>>>(d_sdf, n, level, n_cells, ccount);
  tb = tmp;
  cub::DeviceScan::ExclusiveSum(d_tmp, tb, ccount, coff, (int)n_cells, st);
  ctx->counters.kernel_launches += 4;
  HM_CUDA(cudaGetLastError());
  int32_t last[4];
  HM_CUDA(cudaMemcpyAsync(&last[0], flag + n_edges - 1, 4, cudaMemcpyDeviceToHost, st));
  HM_CUDA(cudaMemcpyAsync(&last[1], vid + n_edges - 1, 4, cudaMemcpyDeviceToHost, st));
  HM_CUDA(cudaMemcpyAsync(&last[2], ccount + n_cells - 1, 4, cudaMemcpyDeviceToHost, st));
  HM_CUDA(cudaMemcpyAsync(&last[3], coff + n_cells - 1, 4, cudaMemcpyDeviceToHost, st));
  HM_CUDA(cudaStreamSynchronize(st));
  const int64_t nv = (int64_t)last[0] + last[1], nf = (int64_t)last[2] + last[3];
  // outputs live in a second allocation sized by the counts
  const size_t out_need = al(sizeof(float) * 3 * std::max<int64_t>(nv, 1)) + al(sizeof(int32_t) * 3 * std::max<int64_t>(nf, 1));
  if (out_need > ctx->mesh_out_bytes) {
    if (ctx->mesh_out) { cudaFree(ctx->mesh_out); ctx->mesh_out = nullptr; ctx->mesh_out_bytes = 0; }
    HM_CUDA(cudaMalloc(&ctx->mesh_out, out_need));
    ctx->mesh_out_bytes = out_need;
  }
  float* verts = (float*)ctx->mesh_out;
  int32_t* faces = (int32_t*)((uint8_t*)ctx->mesh_out + al(sizeof(float) * 3 * std::max<int64_t>(nv, 1)));
  if (nv > 0) vertex_kernel<<<(unsigned)((n_edges + B - 1) / B), B, 0, st>>>(d_sdf, n, level, spacing, n_edges, flag, vid, verts);
  if (nf > 0) face_kernel<<<(unsigned)((n_cells + B - 1) / B), B, 0, st>>>(d_sdf, n, level, n_cells, coff, vid, faces);
  ctx->counters.kernel_launches += 2;
  HM_CUDA(cudaGetLastError());
  ctx->mesh_n_verts = nv;
  ctx->mesh_n_faces = nf;
  *h_n_verts = nv;
  *h_n_faces = nf;
  return HM_OK;
}

// Copies the mesh of the last hm_isosurface call to d_verts [n_verts][3] (fp32) and d_faces [n_faces][3] (int32).  With
// apply_affine != 0 the vertices become (v - 1) * cube_radius (wild_completion/utils.py:583-585).
extern "C" int hm_isosurface_fetch(hm_context* ctx, float* d_verts, int32_t* d_faces, int32_t apply_affine, double cube_radius, void* stream) {
  HM_CHECK(ctx && ctx->mesh_out, "hm_isosurface_fetch: no mesh (call hm_isosurface first)");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  const int64_t nv = ctx->mesh_n_verts, nf = ctx->mesh_n_faces;
  const float* verts = (const float*)ctx->mesh_out;
  const int32_t* faces = (const int32_t*)((const uint8_t*)ctx->mesh_out + al(sizeof(float) * 3 * std::max<int64_t>(nv, 1)));
  if (nv > 0) {
    HM_CHECK(d_verts, "hm_isosurface_fetch: null vertex buffer");
    HM_CUDA(cudaMemcpyAsync(d_verts, verts, sizeof(float) * 3 * nv, cudaMemcpyDeviceToDevice, st));
    if (apply_affine) {
      vertex_affine_kernel<<<(unsigned)((nv * 3 + 255) / 256), 256, 0, st>>>(nv * 3, cube_radius, d_verts);
      ctx->counters.kernel_launches += 1;
    }
  }
  if (nf > 0) {
    HM_CHECK(d_faces, "hm_isosurface_fetch: null face buffer");
    HM_CUDA(cudaMemcpyAsync(d_faces, faces, sizeof(int32_t) * 3 * nf, cudaMemcpyDeviceToDevice, st));
  }
  HM_CUDA(cudaGetLastError());
  return HM_OK;
}
