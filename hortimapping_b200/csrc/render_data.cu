// Device side of wild_completion/utils.py:39-109 get_render_data + :23-38 get_rays (the step BEFORE the hot
// path).  The reference runs it per fruit x per frame in numpy on full images.  Here a frame's id / depth images are uploaded
// once, ONE pass builds the (count, bounding box) record of EVERY id in the frame, and per (fruit, frame) only the <= 300 x 300
// crop is touched: ordered compaction of the background / foreground candidates and the ray gather.  The random subsampling
// (np.random.choice, :77, :88) stays on the host so that a seeded run draws exactly the reference's indices.
// Integer / byte work, bit-exact; the ray directions repeat the reference's fp64 arithmetic.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kMaxIds = 1024;

// table[id] = (count, min_v, max_v, min_u, max_u) over pixels with img == id and depth > 0 (utils.py:50-59)
__global__ void id_bbox_kernel(const int32_t* __restrict__ id_img, const float* __restrict__ depth, int h, int w, int n_ids,
                               int32_t* __restrict__ table) {
  __shared__ int32_t s_tab[kMaxIds * 5];
  for (int i = threadIdx.x; i < n_ids * 5; i += blockDim.x) {
    const int f = i % 5;
    s_tab[i] = (f == 0) ? 0 : ((f == 1 || f == 3) ? INT32_MAX : -1);
  }
  __syncthreads();
  const int64_t total = (int64_t)h * w;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t id = id_img[p];
    if (id >= 0 && id < n_ids && depth[p] > 0.f) {
      const int v = (int)(p / w), u = (int)(p % w);
      atomicAdd(&s_tab[id * 5 + 0], 1);
      atomicMin(&s_tab[id * 5 + 1], v);
      atomicMax(&s_tab[id * 5 + 2], v);
      atomicMin(&s_tab[id * 5 + 3], u);
      atomicMax(&s_tab[id * 5 + 4], u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_ids; i += blockDim.x) {
    if (s_tab[i * 5] == 0) continue;
    atomicAdd(&table[i * 5 + 0], s_tab[i * 5 + 0]);
    atomicMin(&table[i * 5 + 1], s_tab[i * 5 + 1]);
    atomicMax(&table[i * 5 + 2], s_tab[i * 5 + 2]);
    atomicMin(&table[i * 5 + 3], s_tab[i * 5 + 3]);
    atomicMax(&table[i * 5 + 4], s_tab[i * 5 + 4]);
  }
}

__global__ void id_bbox_init_kernel(int n_ids, int32_t* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_ids * 5) return;
  const int f = i % 5;
  table[i] = (f == 0) ? 0 : ((f == 1 || f == 3) ? INT32_MAX : -1);
}

// Ordered compaction of the crop grid (rows hh[], columns ww[], row-major as utils.py:66-71): background candidates are the
// grid pixels whose id differs from the fruit's (:72), foreground those with the id AND a valid depth (:81).  One block walks
// the grid in chunks of blockDim.x and keeps running offsets, so the output order is the reference's.
__global__ void __launch_bounds__(1024) crop_candidates_kernel(const int32_t* __restrict__ id_img, const float* __restrict__ depth, int h, int w,
                                                               int32_t submap_id, const int32_t* __restrict__ hh, int crop_h,
                                                               const int32_t* __restrict__ ww, int crop_w, int32_t* __restrict__ pix_bg,
                                                               float* __restrict__ depth_bg, int32_t* __restrict__ pix_fg,
                                                               float* __restrict__ depth_fg, int32_t* __restrict__ counts) {
  __shared__ int s_warp[2][32];
  __shared__ int s_base[2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x == 0) s_base[0] = s_base[1] = 0;
  __syncthreads();
  const int total = crop_h * crop_w;
  for (int start = 0; start < total; start += blockDim.x) {
    const int i = start + threadIdx.x;
    int is_bg = 0, is_fg = 0, v = 0, u = 0;
    float d = 0.f;
    if (i < total) {
      v = hh[i / crop_w];
      u = ww[i % crop_w];
      if (v >= 0 && v < h && u >= 0 && u < w) {          // callers validate the crop against the image (render_data.py raises like
        const int32_t id = id_img[(int64_t)v * w + u];    // numpy's indexing would); a grid entry outside the image is never read
        d = depth[(int64_t)v * w + u];
        is_bg = (id != submap_id);
        is_fg = (id == submap_id) && (d > 0.f);
      }
    }
    const unsigned mb = __ballot_sync(0xffffffffu, is_bg), mf = __ballot_sync(0xffffffffu, is_fg);
    if (lane == 0) { s_warp[0][wid] = __popc(mb); s_warp[1][wid] = __popc(mf); }
    __syncthreads();
    int off_b = s_base[0], off_f = s_base[1];
    for (int k = 0; k < wid; ++k) { off_b += s_warp[0][k]; off_f += s_warp[1][k]; }
    off_b += __popc(mb & ((1u << lane) - 1u));
    off_f += __popc(mf & ((1u << lane) - 1u));
    if (is_bg) { pix_bg[off_b * 2] = u; pix_bg[off_b * 2 + 1] = v; depth_bg[off_b] = d; }      // [u, v] (:73)
    if (is_fg) { pix_fg[off_f * 2] = u; pix_fg[off_f * 2 + 1] = v; depth_fg[off_f] = d; }      // (:83)
    __syncthreads();
    if (threadIdx.x == 0) {
      int tb = 0, tf = 0;
      for (int k = 0; k < nwarp; ++k) { tb += s_warp[0][k]; tf += s_warp[1][k]; }
      s_base[0] += tb;
      s_base[1] += tf;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { counts[0] = s_base[0]; counts[1] = s_base[1]; }
}

struct InvK { double k[9]; };

// utils.py:23-38 get_rays on the selected candidates (:74-76 / :85-87): directions = float32(([u, v, 1] * invK).sum(-1)) in fp64
__global__ void gather_rays_kernel(const int32_t* __restrict__ pix, const float* __restrict__ depth, const int64_t* __restrict__ sel, int64_t k,
                                   InvK K, float* __restrict__ rays, float* __restrict__ depth_out, int32_t* __restrict__ pix_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const int64_t s = sel ? sel[i] : i;
  const int32_t u = pix[s * 2], v = pix[s * 2 + 1];
  const double du = (double)u, dv = (double)v;
#pragma unroll
  for (int r = 0; r < 3; ++r)
    rays[i * 3 + r] = (float)__dadd_rn(__dadd_rn(__dmul_rn(du, K.k[r * 3 + 0]), __dmul_rn(dv, K.k[r * 3 + 1])), __dmul_rn(1.0, K.k[r * 3 + 2]));
  depth_out[i] = depth[s];
  pix_out[i * 2] = u;
  pix_out[i * 2 + 1] = v;
}

}  // namespace

extern "C" int hm_frame_id_bboxes(hm_context* ctx, const int32_t* d_id_img, const float* d_depth, int32_t h, int32_t w, int32_t n_ids,
                                  int32_t* d_table, void* stream) {
  HM_CHECK(d_id_img && d_depth && d_table && h > 0 && w > 0 && n_ids > 0 && n_ids <= kMaxIds, "hm_frame_id_bboxes: bad argument (n_ids <= %d)", kMaxIds);
  if (ctx) HM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  id_bbox_init_kernel<<<(n_ids * 5 + 255) / 256, 256, 0, st>>>(n_ids, d_table);
  const int64_t total = (int64_t)h * w;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 4);
  id_bbox_kernel<<<blocks, 256, 0, st>>>(d_id_img, d_depth, h, w, n_ids, d_table);
  if (ctx) ctx->counters.kernel_launches += 2;
  HM_CUDA(cudaGetLastError());
  return HM_OK;
}

extern "C" int hm_crop_candidates(hm_context* ctx, const int32_t* d_id_img, const float* d_depth, int32_t h, int32_t w, int32_t submap_id,
                                  const int32_t* d_hh, int32_t crop_h, const int32_t* d_ww, int32_t crop_w, int32_t* d_pix_bg,
                                  float* d_depth_bg, int32_t* d_pix_fg, float* d_depth_fg, int32_t* d_counts, void* stream) {
  HM_CHECK(d_id_img && d_depth && d_hh && d_ww && d_pix_bg && d_depth_bg && d_pix_fg && d_depth_fg && d_counts && h > 0 && w > 0 &&
               crop_h > 0 && crop_w > 0, "hm_crop_candidates: bad argument");
  if (ctx) HM_CUDA(cudaSetDevice(ctx->device));
  crop_candidates_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_id_img, d_depth, h, w, submap_id, d_hh, crop_h, d_ww, crop_w, d_pix_bg, d_depth_bg,
                                                               d_pix_fg, d_depth_fg, d_counts);
  if (ctx) ctx->counters.kernel_launches += 1;
  HM_CUDA(cudaGetLastError());
  return HM_OK;
}

extern "C" int hm_gather_rays(hm_context* ctx, const int32_t* d_pix, const float* d_depth, const int64_t* d_sel, int64_t k,
                              const double* h_invK, float* d_rays, float* d_depth_out, int32_t* d_pix_out, void* stream) {
  HM_CHECK(h_invK && k >= 0 && (k == 0 || (d_pix && d_depth && d_rays && d_depth_out && d_pix_out)), "hm_gather_rays: bad argument");
  if (k == 0) return HM_OK;
  if (ctx) HM_CUDA(cudaSetDevice(ctx->device));
  InvK K;
  for (int i = 0; i < 9; ++i) K.k[i] = h_invK[i];
  gather_rays_kernel<<<(unsigned)((k + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_pix, d_depth, d_sel, k, K, d_rays, d_depth_out, d_pix_out);
  if (ctx) ctx->counters.kernel_launches += 1;
  HM_CUDA(cudaGetLastError());
  return HM_OK;
}
