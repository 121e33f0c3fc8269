// Device-side Levenberg-Marquardt loop of the fruit shape / pose optimisation, batched over fruits.
//
// Restates wild_completion/optimizer.py:28-302 (shape_pose_joint_opt), :306-429 (shape_opt_deepsdf),
// wild_completion/loss.py:8-243 and the math helpers of wild_completion/utils.py:197-358.  The whole
// max_iter loop is enqueued on one stream with no host synchronisation: convergence is tracked per
// fruit in device memory (the reference does 5 host syncs per iteration, optimizer.py:255-259).
//
// Per iteration (joint):  frame_setup -> sample -> decoder forward (all ray samples) -> composite
// (occupancy, transmittance, rendered depth, residuals, d/d sdf coefficients, band selection) ->
// scan/scatter (compaction of the in-band samples) -> decoder forward+Jacobian (recon points +
// in-band samples) -> ray / point Jacobians -> normal-equation partials -> per-fruit solve + update.
// All reductions are order-deterministic (no floating-point atomics).
#include <limits.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kMaxM = 64;        // max depth samples per ray
constexpr int kRedItems = 128;   // items (rays / points) per normal-equation block
constexpr int kE = HM_MAX_EST;   // 39
constexpr int kTri = kE * (kE + 1) / 2;
constexpr int kPartial = kTri + kE;   // upper triangle of J^T W J, then J^T W r

struct FrameState {
  float A[9];
  float t[3];
  float rho;             // object_bbx_radius = cube_radius * cur_scale (optimizer.py:107)
  float delta_d, d_term;
  int fruit;
  int n_fg;
  int64_t ray_begin;
  int32_t n_rays;
  int32_t valid_count;   // in-sphere samples of this frame (loss.py:38-45)
  float depths[kMaxM];
};

struct DevParams {          // hm_opt_params casts done once, where torch would do them
  int M, log_occ, occlusion_on, scale_on, pose_dim, est, robust_iter, min_valid, lm_on, lm_eye;
  float th, sigma, inv_2th, do_ds_lin, occl_th, min_grad, t_depth, t_recon;
  double eps_g, eps_c, eps_t, eps_r, eps_s, w_recon, w_depth, w_mask, w_codereg, lm_lambda_0, s_damp;
};

struct RedBlock {           // one normal-equation block
  int fruit, term;          // term: 0 depth, 1 mask, 2 recon
  int64_t start;            // first item (ray index or point index)
  int count;
};

__device__ __forceinline__ float det3(const float* T) {   // 3x3 block of a row-major 4x4
  return T[0] * (T[5] * T[10] - T[6] * T[9]) - T[1] * (T[4] * T[10] - T[6] * T[8]) + T[2] * (T[4] * T[9] - T[5] * T[8]);
}

// torch.inverse(T)[2,3] evaluated in fp64 by cofactors (the reference's fp32 LU is within ~1 ulp)
__device__ double inv23(const float* Tf) {
  double m[16];
  for (int i = 0; i < 16; ++i) m[i] = Tf[i];
  auto det3d = [&](int r0, int r1, int r2, int c0, int c1, int c2) {
    return m[r0 * 4 + c0] * (m[r1 * 4 + c1] * m[r2 * 4 + c2] - m[r1 * 4 + c2] * m[r2 * 4 + c1]) -
           m[r0 * 4 + c1] * (m[r1 * 4 + c0] * m[r2 * 4 + c2] - m[r1 * 4 + c2] * m[r2 * 4 + c0]) +
           m[r0 * 4 + c2] * (m[r1 * 4 + c0] * m[r2 * 4 + c1] - m[r1 * 4 + c1] * m[r2 * 4 + c0]);
  };
  double det = 0;
  {
    // Laplace expansion along row 3
    double c30 = -det3d(0, 1, 2, 1, 2, 3), c31 = det3d(0, 1, 2, 0, 2, 3), c32 = -det3d(0, 1, 2, 0, 1, 3), c33 = det3d(0, 1, 2, 0, 1, 2);
    det = m[12] * c30 + m[13] * c31 + m[14] * c32 + m[15] * c33;
  }
  // inv[2][3] = cofactor(3,2) / det ; cofactor(3,2) = (-1)^(3+2) * minor(3,2) (delete row 3, col 2)
  double minor = det3d(0, 1, 2, 0, 1, 3);
  return -minor / det;
}

// ---------------------------------------------------------------------------------------------
// optimizer.py:104-111: T_oc = T_ow @ T_wc, depth window from inverse(T_oc)[2,3], linspace samples
// ---------------------------------------------------------------------------------------------
__global__ void frame_setup_kernel(int n_frames, FrameState* __restrict__ fs, const float* __restrict__ T_ow,
                                   const float* __restrict__ T_wc, const float* __restrict__ cube_radius,
                                   const uint8_t* __restrict__ active, int M, int32_t* __restrict__ n_valid_total) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g == 0) *n_valid_total = 0;            // the sample kernel of this iteration compacts the in-sphere samples from slot 0
  if (g >= n_frames) return;
  FrameState& F = fs[g];
  F.valid_count = 0;
  if (!active[F.fruit]) return;
  const float* A = T_ow + (size_t)F.fruit * 16;
  const float* B = T_wc + (size_t)g * 16;
  float T[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s = fmaf(A[i * 4 + k], B[k * 4 + j], s);
      T[i * 4 + j] = s;
    }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) F.A[i * 3 + j] = T[i * 4 + j];
    F.t[i] = T[i * 4 + 3];
  }
  float cur_scale = powf(det3(A), (float)(-1.0 / 3.0));        // optimizer.py:66,250
  float rho = cube_radius[F.fruit] * cur_scale;                // :107
  float zc = (float)inv23(T);
  float d_min = zc - 1.0f * rho, d_max = zc + 0.8f * rho;      // :110
  // torch.linspace (:111): step = (end - start) / (steps - 1); first half from start, second half from end
  float step = (d_max - d_min) / (float)(M - 1);
  for (int i = 0; i < M; ++i) F.depths[i] = (i < M / 2) ? __fadd_rn(d_min, __fmul_rn(step, (float)i)) : __fsub_rn(d_max, __fmul_rn(step, (float)(M - 1 - i)));
  F.rho = rho;
  F.delta_d = (F.depths[M - 1] - F.depths[0]) / (float)(M - 1);   // loss.py:73-75
  F.d_term = F.depths[M - 1] + F.delta_d;                          // loss.py:78
}

// ---------------------------------------------------------------------------------------------
// loss.py:30-40: sample points on the rays, camera -> object, in-sphere test
// ---------------------------------------------------------------------------------------------
// The in-sphere samples are COMPACTED here for the decoder (the reference decodes only them, loss.py:47-49): a warp-aggregated
// integer atomic hands out slots, so the compact order varies from run to run, but every row is evaluated independently of its
// neighbours and its SDF is written back to the sample's own position (idx_c), so the results do not depend on that order.
__global__ void __launch_bounds__(256) sample_kernel(int64_t n_samples, int M, FrameState* __restrict__ fs, const int32_t* __restrict__ ray_frame,
                                                     const float* __restrict__ rays, const uint8_t* __restrict__ active, float* __restrict__ xyz,
                                                     uint8_t* __restrict__ valid, float* __restrict__ xyz_c, int32_t* __restrict__ idx_c,
                                                     int32_t* __restrict__ row_latent_c, int32_t* __restrict__ n_valid_total,
                                                     int32_t* __restrict__ cidx_of /* [n_samples] or NULL: sample -> compact row */) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool v = false;
  int g = -1, fruit = 0;
  float p[3] = {0.f, 0.f, 0.f};
  if (i < n_samples) {
    const int64_t r = i / M;
    const int m = (int)(i % M);
    g = ray_frame[r];
    const FrameState& F = fs[g];
    fruit = F.fruit;
    if (active[fruit]) {
      const float d = F.depths[m];
      const float cx = __fmul_rn(rays[r * 3 + 0], d), cy = __fmul_rn(rays[r * 3 + 1], d), cz = __fmul_rn(rays[r * 3 + 2], d);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float s = __fadd_rn(__fadd_rn(__fmul_rn(cx, F.A[k * 3 + 0]), __fmul_rn(cy, F.A[k * 3 + 1])), __fmul_rn(cz, F.A[k * 3 + 2]));
        p[k] = __fadd_rn(s, F.t[k]);
      }
      const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])), __fmul_rn(p[2], p[2])));
      v = nrm < F.rho;
      xyz[i * 3 + 0] = p[0];
      xyz[i * 3 + 1] = p[1];
      xyz[i * 3 + 2] = p[2];
    }
    valid[i] = v ? 1 : 0;
  }
  const unsigned vm = __ballot_sync(0xffffffffu, v);
  if (vm == 0u) return;
  int base = 0;
  const int leader = __ffs(vm) - 1;
  if (lane == leader) base = atomicAdd(n_valid_total, __popc(vm));
  base = __shfl_sync(0xffffffffu, base, leader);
  // per-frame in-sphere counts (loss.py:38-45): one integer atomic per (warp, frame)
  const unsigned same = __match_any_sync(0xffffffffu, v ? g : -1 - lane);
  if (v) {
    if (lane == __ffs(same) - 1) atomicAdd(&fs[g].valid_count, __popc(same));
    const int slot = base + __popc(vm & ((1u << lane) - 1u));
    xyz_c[(int64_t)slot * 3 + 0] = p[0];
    xyz_c[(int64_t)slot * 3 + 1] = p[1];
    xyz_c[(int64_t)slot * 3 + 2] = p[2];
    idx_c[slot] = (int32_t)i;
    row_latent_c[slot] = fruit;
    if (cidx_of) cidx_of[i] = slot;
  }
}

// ---------------------------------------------------------------------------------------------
// loss.py:55-176 per ray: occupancy, transmittance, rendered depth / occupancy, band selection,
// d(depth)/d(occ), d(mask)/d(occ), occlusion filter, per-ray residuals.  One thread per ray.
// Also the in-block exclusive scan of the per-ray survivor counts (for the compaction).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) composite_kernel(int64_t n_rays, DevParams P, const FrameState* __restrict__ fs,
                                                        const int32_t* __restrict__ ray_frame, const float* __restrict__ depth_obs,
                                                        const uint8_t* __restrict__ valid, const float* __restrict__ sdf,
                                                        const uint8_t* __restrict__ active, float* __restrict__ coef_e,
                                                        float* __restrict__ coef_m, unsigned long long* __restrict__ ray_mask,
                                                        float* __restrict__ res_d, float* __restrict__ res_m,
                                                        int32_t* __restrict__ ray_k, int32_t* __restrict__ ray_off_in_block,
                                                        int32_t* __restrict__ block_sum, int32_t* __restrict__ status) {
  __shared__ int32_t s_scan[256];
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  int k = 0;
  if (r < n_rays) {
    const int g = ray_frame[r];
    const FrameState& F = fs[g];
    const int M = P.M;
    unsigned long long keep = 0ull;
    float rd = 0.f, rm = 0.f;
    if (active[F.fruit]) {
      if (F.valid_count < P.min_valid) {
        if (r == F.ray_begin) atomicOr(&status[F.fruit], HM_STATUS_FRAME_SKIPPED);   // optimizer.py:130-132
      } else {
        const int64_t s0 = r * M;
        float o[kMaxM], tr[kMaxM];
        float T = 1.f, occ_ray = 0.f, d_u = 0.f;
        unsigned long long band = 0ull;
        for (int m = 0; m < M; ++m) {
          float om = 0.f;
          if (valid[s0 + m]) {
            float s = sdf[s0 + m];
            if (P.log_occ) {
              float z = -s / P.sigma;                         // utils.py:142 sigmoid(-sdf / sigma)
              om = 1.f / (1.f + expf(-z));
            } else {
              float c = fminf(fmaxf(s, -P.th), P.th);         // utils.py:131-132
              om = 0.5f - c / P.inv_2th;
            }
            if (s > -P.th && s < P.th) band |= 1ull << m;     // loss.py:66
          }
          float pq = om * T;                                  // o_q * Tr_{q-1}   (loss.py:91)
          occ_ray += pq;                                      // :93
          d_u = fmaf(F.depths[m], pq, d_u);                   // :96
          T *= (1.f - om);                                    // cumprod (:81)
          o[m] = om;
          tr[m] = T;
        }
        d_u = fmaf(F.d_term, T, d_u);                         // termination sample with o = 1
        const float tr_last = T;
        // suffix sums of the transmittance (loss.py:103-107)
        float suf = 0.f;
        float e_s[kMaxM];
        for (int m = M - 1; m >= 0; --m) {
          suf += tr[m];
          e_s[m] = suf;
        }
        const int64_t j = r - F.ray_begin;
        const bool is_bg = j >= F.n_fg;
        const float y = depth_obs[r];
        bool occluded = false;
        if (P.occlusion_on) occluded = is_bg && (y < d_u - P.occl_th) && (y > 0.f);   // loss.py:135
        for (int m = 0; m < M; ++m) {
          if (!((band >> m) & 1ull)) continue;
          float one_minus = 1.f - o[m];
          float dm_do = tr_last / one_minus;                  // :102
          float de_do = e_s[m] * F.delta_d / one_minus;       // :107
          if (!(de_do > P.min_grad)) continue;                // :111
          if (occluded) continue;                             // :136-139
          float do_ds = P.log_occ ? (-o[m] * (1.f - o[m]) / P.sigma) : P.do_ds_lin;   // :121-123
          coef_e[s0 + m] = de_do * do_ds;                     // :126
          coef_m[s0 + m] = dm_do * do_ds;                     // :127
          keep |= 1ull << m;
          ++k;
        }
        if (k > 0) {
          const float yy = is_bg ? F.d_term : y;              // :142
          const float res = yy - d_u;                         // :155
          // per-ray mean of k identical values, accumulated like scatter_add_ (:168-169)
          float acc_r = 0.f, acc_o = 0.f;
          for (int q = 0; q < k; ++q) { acc_r += res; acc_o += occ_ray; }
          rd = acc_r / (float)k;
          rm = acc_o / (float)k - (is_bg ? 0.f : 1.f);        // :175
        }
      }
    }
    ray_mask[r] = keep;
    res_d[r] = rd;
    res_m[r] = rm;
    ray_k[r] = k;
  }
  // exclusive scan of k within the block
  s_scan[threadIdx.x] = k;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    int v = (threadIdx.x >= off) ? s_scan[threadIdx.x - off] : 0;
    __syncthreads();
    s_scan[threadIdx.x] += v;
    __syncthreads();
  }
  if (r < n_rays) ray_off_in_block[r] = s_scan[threadIdx.x] - k;
  if (threadIdx.x == 255) block_sum[blockIdx.x] = s_scan[255];
}

// exclusive scan of the block sums (single block) + per-fruit output-ray counts + dynamic row count
__global__ void __launch_bounds__(1024) scan_blocks_kernel(int n_blocks, const int32_t* __restrict__ block_sum,
                                                           int32_t* __restrict__ block_base, int32_t* __restrict__ n_rows_dyn,
                                                           int32_t n_static_rows) {
  __shared__ int32_t s[1024];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < n_blocks) ? block_sum[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      int t = (threadIdx.x >= off) ? s[threadIdx.x - off] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n_blocks) block_base[i] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    n_rows_dyn[0] = n_static_rows + carry;       // all gradient rows: observed points + in-band samples
    n_rows_dyn[1] = carry;                       // the in-band samples alone
  }
}

// scatter the surviving samples into the compact grad-row list (after the static recon rows)
__global__ void scatter_kernel(int64_t n_rays, int M, const int32_t* __restrict__ ray_frame, const FrameState* __restrict__ fs,
                               const unsigned long long* __restrict__ ray_mask, const int32_t* __restrict__ ray_off_in_block,
                               const int32_t* __restrict__ block_base, const float* __restrict__ xyz_s, int32_t n_static_rows,
                               float* __restrict__ xyz_g, int32_t* __restrict__ row_latent_g, int32_t* __restrict__ ray_slot,
                               const int32_t* __restrict__ cidx_of, const float* __restrict__ sdf_s, int32_t* __restrict__ src_g,
                               float* __restrict__ sdf_g) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  unsigned long long keep = ray_mask[r];
  int slot = n_static_rows + block_base[r / 256] + ray_off_in_block[r];
  ray_slot[r] = slot;
  if (!keep) return;
  const int fruit = fs[ray_frame[r]].fruit;
  for (int m = 0; m < M; ++m) {
    if (!((keep >> m) & 1ull)) continue;
    const int64_t s = r * M + m;
    xyz_g[(int64_t)slot * 3 + 0] = xyz_s[s * 3 + 0];
    xyz_g[(int64_t)slot * 3 + 1] = xyz_s[s * 3 + 1];
    xyz_g[(int64_t)slot * 3 + 2] = xyz_s[s * 3 + 2];
    row_latent_g[slot] = fruit;
    if (src_g) {        // gradient-only decode of this row: where the forward pass evaluated it, and the SDF it got there
      src_g[slot - n_static_rows] = cidx_of[s];
      sdf_g[slot] = sdf_s[s];
    }
    ++slot;
  }
}

// optimizer.py:168 / :343: cur_points_o = (points_w[..., None, :] * T_ow[:3,:3]).sum(-1) + T_ow[:3,3]
__global__ void transform_points_kernel(int64_t n_points, const int32_t* __restrict__ point_fruit, const float* __restrict__ pts_w,
                                        const float* __restrict__ T_ow, float* __restrict__ xyz_g, int32_t* __restrict__ row_latent_g) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const int f = point_fruit[i];
  const float* T = T_ow + (size_t)f * 16;
  const float x = pts_w[i * 3], y = pts_w[i * 3 + 1], z = pts_w[i * 3 + 2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float s = __fadd_rn(__fadd_rn(__fmul_rn(x, T[k * 4 + 0]), __fmul_rn(y, T[k * 4 + 1])), __fmul_rn(z, T[k * 4 + 2]));
    xyz_g[i * 3 + k] = __fadd_rn(s, T[k * 4 + 3]);
  }
  row_latent_g[i] = f;
}

// huber weight squared (utils.py:327-358): w^2 with w = sqrt(huber(|r|)) / |r|; |r| == 0 -> 0
__device__ __forceinline__ float huber_w2(float r, float b) {
  float a = fabsf(r);
  float rn = (a <= b) ? a * a : (2.f * b * a - b * b);
  float den = (a == 0.f) ? 1.f : a;
  float w = sqrtf(rn) / den;
  return w * w;
}

// d x_o / d pose (left perturbation, [I | -hat(p) | p], utils.py:197-276) applied to a = coef * dsdf/dxyz
__device__ __forceinline__ void add_pose_jac(float* J, int pose_dim, float a0, float a1, float a2, float x, float y, float z) {
  J[0] += a0;
  J[1] += a1;
  J[2] += a2;
  J[3] += __fadd_rn(__fadd_rn(__fmul_rn(a0, 0.f), __fmul_rn(a1, -z)), __fmul_rn(a2, y));
  J[4] += __fadd_rn(__fadd_rn(__fmul_rn(a0, z), __fmul_rn(a1, 0.f)), __fmul_rn(a2, -x));
  J[5] += __fadd_rn(__fadd_rn(__fmul_rn(a0, -y), __fmul_rn(a1, x)), __fmul_rn(a2, 0.f));
  if (pose_dim == 7) J[6] += __fadd_rn(__fadd_rn(__fmul_rn(a0, x), __fmul_rn(a1, y)), __fmul_rn(a2, z));
}

// loss.py:185-215: per-ray Jacobians = sum over the ray's surviving samples (one thread per ray)
__global__ void ray_jacobian_kernel(int64_t n_rays, int M, int pose_dim, const unsigned long long* __restrict__ ray_mask,
                                    const int32_t* __restrict__ ray_slot, const float* __restrict__ coef_e,
                                    const float* __restrict__ coef_m, const float* __restrict__ xyz_g, const float* __restrict__ jac_g,
                                    float* __restrict__ J_d, float* __restrict__ J_m) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const int est = pose_dim + HM_LATENT;
  float jd[kE], jm[kE];
#pragma unroll
  for (int c = 0; c < kE; ++c) jd[c] = jm[c] = 0.f;
  unsigned long long keep = ray_mask[r];
  int slot = ray_slot[r];
  for (int m = 0; m < M && keep; ++m) {
    if (!((keep >> m) & 1ull)) continue;
    const float ce = coef_e[r * M + m], cm = coef_m[r * M + m];
    const float* g = jac_g + (int64_t)slot * HM_IN;
    const float x = xyz_g[(int64_t)slot * 3], y = xyz_g[(int64_t)slot * 3 + 1], z = xyz_g[(int64_t)slot * 3 + 2];
    add_pose_jac(jd, pose_dim, ce * g[32], ce * g[33], ce * g[34], x, y, z);
    add_pose_jac(jm, pose_dim, cm * g[32], cm * g[33], cm * g[34], x, y, z);
    for (int c = 0; c < HM_LATENT; ++c) {
      jd[pose_dim + c] += ce * g[c];
      jm[pose_dim + c] += cm * g[c];
    }
    ++slot;
  }
  for (int c = 0; c < est; ++c) {
    J_d[r * kE + c] = jd[c];
    J_m[r * kE + c] = jm[c];
  }
}

// loss.py:219-243 compute_sdf_loss (pose_dim = 0 for shape_opt_deepsdf: only the code Jacobian is used)
__global__ void point_jacobian_kernel(int64_t n_points, int pose_dim, const float* __restrict__ xyz_g, const float* __restrict__ sdf_g,
                                      const float* __restrict__ jac_g, float* __restrict__ res, float* __restrict__ J) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  float j[kE];
#pragma unroll
  for (int c = 0; c < kE; ++c) j[c] = 0.f;
  const float* g = jac_g + i * HM_IN;
  if (pose_dim) add_pose_jac(j, pose_dim, g[32], g[33], g[34], xyz_g[i * 3], xyz_g[i * 3 + 1], xyz_g[i * 3 + 2]);
  for (int c = 0; c < HM_LATENT; ++c) j[pose_dim + c] = g[c];
  res[i] = sdf_g[i];
  const int est = pose_dim + HM_LATENT;
  for (int c = 0; c < est; ++c) J[i * kE + c] = j[c];
}

// optimizer.py:152-159,189-190: partial sums of w * J^T J (upper triangle) and w * J^T r over <= 128 items (one block).
// The block's Jacobian rows are staged in shared memory with coalesced loads ([J | r] per item, rows padded to 44 floats), then
// each thread accumulates one 4 x 4 tile of J^T W [J | r] over the items in item order (two 16-byte shared loads and 16 FMAs per
// item).  Tiles below the diagonal are computed too (the arithmetic is trivial) but not written.  Every output is the same
// left-to-right fmaf chain over the items as a scalar loop, so partials are deterministic and independent of the tiling.
constexpr int kJRow = 44;       // kE + 1 (the residual column) rounded up to a multiple of 4

__global__ void __launch_bounds__(128) normal_eq_kernel(const RedBlock* __restrict__ blocks, int est, int iter, DevParams P,
                                                        const float* __restrict__ J_d, const float* __restrict__ J_m,
                                                        const float* __restrict__ res_d, const float* __restrict__ res_m,
                                                        const int32_t* __restrict__ ray_k, const float* __restrict__ J_r,
                                                        const float* __restrict__ res_r, int r_stride, const uint8_t* __restrict__ active,
                                                        float* __restrict__ partials, int32_t* __restrict__ block_items) {
  __shared__ __align__(16) float sJ[kRedItems][kJRow];
  __shared__ float sW[kRedItems];
  __shared__ int s_cnt;
  const RedBlock B = blocks[blockIdx.x];
  float* out = partials + (size_t)blockIdx.x * kPartial;
  if (!active[B.fruit]) return;
  if (threadIdx.x == 0) s_cnt = 0;
  const float* J = (B.term == 0) ? J_d : (B.term == 1) ? J_m : J_r;
  const float* R = (B.term == 0) ? res_d : (B.term == 1) ? res_m : res_r;
  const int stride = (B.term == 2) ? r_stride : kE;      // the latent-only loop reads the decoder's Jacobian rows in place
  const bool robust = iter >= P.robust_iter;
  const int t = threadIdx.x;
  for (int e = t; e < kRedItems * kJRow; e += 128) (&sJ[0][0])[e] = 0.f;
  __syncthreads();
  // coalesced copy of the block's rows: element e of the contiguous range [start * stride, (start + count) * stride)
  const float* src = J + (size_t)B.start * stride;
  for (int e = t; e < B.count * stride; e += 128) {
    const int row = e / stride, col = e - row * stride;
    if (col < est) sJ[row][col] = src[e];
  }
  __syncthreads();
  {
    float w = 0.f;
    if (t < B.count) {
      const int64_t it = B.start + t;
      const bool ok = (B.term == 2) ? true : (ray_k[it] > 0);
      if (ok) {
        const float rr = R[it];
        w = 1.f;
        if (robust && B.term == 0) w = huber_w2(rr, P.t_depth);      // optimizer.py:145-149
        if (robust && B.term == 2) w = huber_w2(rr, P.t_recon);      // :183-187
        sJ[t][est] = rr;
        atomicAdd(&s_cnt, 1);
      } else {
        for (int c = 0; c < est; ++c) sJ[t][c] = 0.f;                 // a ray without surviving samples contributes nothing
      }
    }
    sW[t] = w;
  }
  __syncthreads();
  const int na = (est + 3) / 4, nb = (est + 4) / 4;      // 4-wide blocks of the rows a (0 .. est-1) and columns b (0 .. est, column est = r)
  if (t < na * nb) {
    const int ta = t / nb, tb = t - ta * nb;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (tb >= ta) {                                        // tiles entirely below the diagonal are never written
      for (int i = 0; i < kRedItems; ++i) {
        const float4 ja = *reinterpret_cast<const float4*>(&sJ[i][4 * ta]);
        const float4 jb = *reinterpret_cast<const float4*>(&sJ[i][4 * tb]);
        const float w = sW[i];
        const float wa[4] = {w * ja.x, w * ja.y, w * ja.z, w * ja.w};
        const float bb[4] = {jb.x, jb.y, jb.z, jb.w};
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(wa[x], bb[y], acc[x][y]);
      }
      const int tri = est * (est + 1) / 2;
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const int ra = 4 * ta + x, cb = 4 * tb + y;
          if (ra >= est || cb > est) continue;
          if (cb == est) out[tri + ra] = acc[x][y];                                         // J^T W r
          else if (cb >= ra) out[ra * est - ra * (ra - 1) / 2 + (cb - ra)] = acc[x][y];      // upper triangle, row-major
        }
    }
  }
  if (threadIdx.x == 0) block_items[blockIdx.x] = s_cnt;
}

// ---- small fp32 helpers mirroring utils.py:220-324 ----
__device__ void mat3_mul(const float* A, const float* B, float* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s = fmaf(A[i * 3 + k], B[k * 3 + j], s);
      C[i * 3 + j] = s;
    }
}

__device__ void exp_pose(const float* x, int pose_dim, float* T /*16*/) {
  const float v[3] = {x[0], x[1], x[2]};
  const float w[3] = {x[3], x[4], x[5]};
  const float W[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
  float W2[9];
  mat3_mul(W, W, W2);
  const float theta = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const float theta_2 = theta * theta;
  const float st = sinf(theta), ct = cosf(theta);
  float R[9], Jm[9];
  const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float e_s = 1.f;
  if (pose_dim == 7) {
    // exp_sim3 (utils.py:279-324)
    const float s = x[6];
    e_s = expf(s);
    const float s_2 = s * s;
    if (theta <= 1e-8f) {
      for (int i = 0; i < 9; ++i) R[i] = I[i];
      if (s == 0.f) {
        for (int i = 0; i < 9; ++i) Jm[i] = I[i];
      } else {
        const float c = (e_s - 1.f) / s;
        for (int i = 0; i < 9; ++i) Jm[i] = c * I[i];
      }
    } else {
      for (int i = 0; i < 9; ++i) R[i] = I[i] + W[i] * st / theta + W2[i] * (1.f - ct) / theta_2;
      const float a = e_s * st, b = e_s * ct;
      const float c = (s <= 1e-8f) ? 0.f : (e_s - 1.f) / s;     // utils.py:314: c = 0 for every non-positive s
      const float k1 = (a * s + (1.f - b) * theta) / (s_2 + theta_2);
      const float k2 = c - ((b - 1.f) * s + a * theta) / (s_2 + theta_2);
      for (int i = 0; i < 9; ++i) Jm[i] = c * I[i] + k1 * W[i] / theta + k2 * W2[i] / theta_2;
    }
  } else {
    // exp_se3 (utils.py:220-254)
    if (theta <= 1e-8f) {
      for (int i = 0; i < 9; ++i) { R[i] = I[i]; Jm[i] = I[i]; }
    } else {
      const float theta_3 = theta * theta * theta;
      const float k1 = (1.f - ct) / theta_2, k2 = (theta - st) / theta_3;
      for (int i = 0; i < 9; ++i) {
        R[i] = I[i] + W[i] * st / theta + W2[i] * (1.f - ct) / theta_2;
        Jm[i] = I[i] + k1 * W[i] + k2 * W2[i];
      }
    }
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[i * 4 + j] = e_s * R[i * 3 + j];
    float s = 0.f;
    for (int k = 0; k < 3; ++k) s = fmaf(Jm[i * 3 + k], v[k], s);
    T[i * 4 + 3] = s;
  }
  T[12] = T[13] = T[14] = 0.f;
  T[15] = 1.f;
}

// ---------------------------------------------------------------------------------------------
// optimizer.py:200-291: assemble H, b from the partials, LM damping, solve, exp-map update, stop tests.
// One block per fruit; the 39x39 solve runs in fp64 (Gauss-Jordan elimination with partial pivoting, one matrix element per thread).
// ---------------------------------------------------------------------------------------------
struct SolveArgs {
  const int32_t* fruit_block_begin;   // [n_fruits + 1] blocks of fruit f (sorted by fruit, then term)
  const RedBlock* blocks;
  const float* partials;
  const int32_t* block_items;
  const float* cube_radius;
  const uint8_t* pose_known;
  float* latents;
  float* T_ow;
  uint8_t* active;
  int32_t* iter_count;
  int32_t* status;
  float* last_H;
  float* last_b;
  float* last_dx;
  int joint, iter, iter_first, iter_last;
};

constexpr int kSolveThreads = 1024;      // 32 x 32: thread (ty, tx) owns the elements (ty + 32 i, tx + 32 j) of the augmented system

__global__ void __launch_bounds__(kSolveThreads) solve_kernel(SolveArgs a, DevParams P) {
  const int f = blockIdx.x;
  if (!a.active[f]) return;
  const int est = a.joint ? P.est : HM_LATENT;
  const int pd = a.joint ? P.pose_dim : 0;
  const int tri = est * (est + 1) / 2;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  __shared__ double sH[kE][kE + 2];      // augmented [H | b]; the odd row stride keeps column accesses off a single bank
  __shared__ double sAcc[3][kPartial];
  __shared__ int s_n[3];
  __shared__ int s_piv, s_fail;
  const int b0 = a.fruit_block_begin[f], b1 = a.fruit_block_begin[f + 1];
  // The blocks of a fruit are sorted by term: find the three ranges once, so that the sums below are plain loops whose loads
  // are independent of each other.
  __shared__ int s_tb[4];                // blocks [s_tb[t], s_tb[t + 1]) carry term t
  if (tid < 4) s_tb[tid] = (tid == 0) ? b0 : b1;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  for (int b = b0 + 1 + tid; b < b1; b += kSolveThreads) {
    const int t0 = a.blocks[b - 1].term, t1 = a.blocks[b].term;
    for (int t = t0 + 1; t <= t1; ++t) s_tb[t] = b;          // first block of term t (and of any empty term before it)
  }
  __syncthreads();
  if (tid == 0 && b1 > b0) {                                  // terms missing at the front / back of the range
    const int first = a.blocks[b0].term, last = a.blocks[b1 - 1].term;
    for (int t = 1; t <= first; ++t) s_tb[t] = b0;
    for (int t = last + 1; t < 3; ++t) s_tb[t] = b1;
  }
  __syncthreads();
  // fixed-order fp64 sum of the block partials per term (the latent-only loop has the recon term only)
  const int term_lo = a.joint ? 0 : 2;
  for (int e = tid + term_lo * (tri + est); e < 3 * (tri + est); e += kSolveThreads) {
    const int term = e / (tri + est), idx = e % (tri + est);
    const int tb0 = s_tb[term], tb1 = s_tb[term + 1];
    double acc = 0.0;
#pragma unroll 4
    for (int b = tb0; b < tb1; ++b) acc += (double)a.partials[(size_t)b * kPartial + idx];
    sAcc[term][idx] = acc;
  }
  if (tid < 3) {
    int n = 0;
    for (int b = s_tb[tid]; b < s_tb[tid + 1]; ++b) n += a.block_items[b];
    s_n[tid] = n;
  }
  __syncthreads();
  const int n_d = s_n[0], n_r = s_n[2];
  if ((a.joint && n_d == 0) || n_r == 0) {
    // optimizer.py:139-141 "This submap is not valid" (no ray survived) / no surface points: nothing to optimise (the reference
    // would fail on the empty tensor)
    if (tid == 0) { a.active[f] = 0; atomicOr(&a.status[f], HM_STATUS_SUBMAP_INVALID); }
    return;
  }
  const double wd = a.joint ? P.w_depth / (double)n_d : 0.0, wm = a.joint ? P.w_mask / (double)n_d : 0.0, wr = P.w_recon / (double)n_r;
  float* lat = a.latents + (size_t)f * HM_LATENT;
  for (int e = tid; e < tri + est; e += kSolveThreads) {
    double v = a.joint ? wd * sAcc[0][e] + wm * sAcc[1][e] + wr * sAcc[2][e] : wr * sAcc[2][e];   // (terms 0, 1 are not summed when !joint)
    if (e < tri) {
      int r = 0, rem = e;
      while (rem >= est - r) { rem -= est - r; ++r; }
      const int c = r + rem;
      sH[r][c] = v;
      sH[c][r] = v;
    } else {
      sH[e - tri][est] = -v;                        // b = -w J^T r (:153)
    }
  }
  __syncthreads();
  if (tid < HM_LATENT) {                            // code regulariser (:200-203)
    sH[pd + tid][pd + tid] += P.w_codereg;
    sH[pd + tid][est] += -P.w_codereg * (double)lat[tid];
  }
  if (tid == 32 && a.joint && P.scale_on) sH[pd - 1][pd - 1] += P.s_damp;     // :217-218
  __syncthreads();
  if (P.lm_on) {                                                              // :220-225
    if (P.lm_eye) {
      double mx = sH[0][0];
      for (int i = 1; i < est; ++i) mx = fmax(mx, sH[i][i]);
      __syncthreads();
      if (tid < est) sH[tid][tid] += P.lm_lambda_0 * mx;
    } else if (tid < est) {
      sH[tid][tid] += P.lm_lambda_0 * sH[tid][tid];
    }
    __syncthreads();
  }
  if (a.last_H) {
    for (int e = tid; e < est * est; e += kSolveThreads) a.last_H[(size_t)f * kE * kE + e] = (float)sH[e / est][e % est];
    if (tid < est) a.last_b[(size_t)f * kE + tid] = (float)sH[tid][est];
  }
  // max |b| for the gradient stop test (:276), by warp 0
  __shared__ float s_bmax, s_cmax;
  __shared__ int s_cnan, s_st_pose;
  __shared__ float s_dx[kE];
  if (tid < 32) {
    float m = 0.f;
    for (int i = tid; i < est; i += 32) m = fmaxf(m, fabsf((float)sH[i][est]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (tid == 0) s_bmax = m;
  }
  // delta_x = H^-1 b (:234) by Gauss-Jordan elimination on [H | b] in fp64, one matrix element per thread.  H is symmetric
  // positive definite here (J^T W J with non-negative weights + code regulariser + LM damping), for which elimination without
  // row exchanges is backward stable; so the fast path takes the diagonal as pivot: step k reads row k and column k and writes
  // neither (column k of the other rows is simply never read again), ONE barrier per step.  A pivot that is not a positive finite
  // number (a semi-definite system, e.g. Gauss-Newton with an unobservable pose) sends the fruit to the partial-pivoting path on a
  // saved copy of the system -- what torch.inverse's LU does.
  __shared__ double sH0[kE][kE + 2];
  for (int e = tid; e < est * (est + 1); e += kSolveThreads) sH0[e / (est + 1)][e % (est + 1)] = sH[e / (est + 1)][e % (est + 1)];
  __syncthreads();
  for (int k = 0; k < est; ++k) {
    const double piv = sH[k][k];
    if (!(piv > 0.0) || !isfinite(piv)) { if (tid == 0) s_fail = 1; }
    const double rp = 1.0 / piv;
    for (int i = ty; i < est; i += 32) {
      if (i == k) continue;
      const double m = sH[i][k] * rp;
      for (int c = tx; c <= est; c += 32)
        if (c > k) sH[i][c] -= m * sH[k][c];
    }
    __syncthreads();
    if (s_fail) break;
  }
  if (s_fail) {
    // ---- partial pivoting (first maximum on ties), three barriers per step
    __syncthreads();
    for (int e = tid; e < est * (est + 1); e += kSolveThreads) sH[e / (est + 1)][e % (est + 1)] = sH0[e / (est + 1)][e % (est + 1)];
    if (tid == 0) s_fail = 0;
    __syncthreads();
    for (int k = 0; k < est; ++k) {
      if (tid < 32) {
        double best = -1.0;
        int p = est;
        for (int i = k + tid; i < est; i += 32) {
          const double v = fabs(sH[i][k]);
          if (v > best) { best = v; p = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, off);
          const int op = __shfl_xor_sync(0xffffffffu, p, off);
          if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
        }
        if (tid == 0) {
          s_piv = p;
          if (!(best > 0.0) || !isfinite(best)) s_fail = 1;          // singular or non-finite system
        }
      }
      __syncthreads();
      if (s_fail) break;
      const int p = s_piv;
      if (p != k && ty == 0)
        for (int c = tx; c <= est; c += 32) { const double t = sH[k][c]; sH[k][c] = sH[p][c]; sH[p][c] = t; }
      __syncthreads();
      const double piv = sH[k][k];
      for (int i = ty; i < est; i += 32) {
        if (i == k) continue;
        const double m = sH[i][k] / piv;
        for (int c = tx; c <= est; c += 32)
          if (c > k) sH[i][c] -= m * sH[k][c];
      }
      __syncthreads();
    }
  }
  if (tid < est) {
    const float d = s_fail ? 0.f : (float)(sH[tid][est] / sH[tid][tid]);
    if (!isfinite(d)) s_fail = 1;
    s_dx[tid] = d;
  }
  __syncthreads();
  if (s_fail) {
    // torch.inverse raises on a singular matrix; here the fruit stops with its state untouched and a status bit
    if (tid == 0) { atomicOr(&a.status[f], HM_STATUS_SOLVE_FAILED); a.active[f] = 0; }
    return;
  }
  if (a.last_dx && tid < est) a.last_dx[(size_t)f * kE + tid] = s_dx[tid];
  // ---- update (:235-253) and stop tests (:276-291): the latent by warp 0, the pose by one lane of warp 1, concurrently
  if (tid < 32) {
    const float dc = s_dx[pd + tid];
    const float ln = lat[tid] + dc;                                        // :248 / :401
    lat[tid] = ln;
    const float q = fabsf(dc / (ln + 1e-12f));                             // :280 uses the updated latent
    const unsigned nan_any = __ballot_sync(0xffffffffu, q != q);
    float m = (q != q) ? 0.f : q;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (tid == 0) { s_cmax = m; s_cnan = nan_any != 0u; }
  } else if (tid == 32) {
    int pose_conv = 0;
    if (a.joint) {
      float dx[kE];
      for (int i = 0; i < pd; ++i) dx[i] = s_dx[i];
      if (a.pose_known[f]) for (int i = 0; i < 6; ++i) dx[i] = 0.f;        // :237-238 (scale is still optimised)
      float dT[16], Tn[16];
      exp_pose(dx, pd, dT);                                                // :242-245
      float* T = a.T_ow + (size_t)f * 16;
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          float s = 0.f;
          for (int k = 0; k < 4; ++k) s = fmaf(dT[i * 4 + k], T[k * 4 + j], s);
          Tn[i * 4 + j] = s;
        }
      for (int i = 0; i < 16; ++i) T[i] = Tn[i];                           // :247
      const float cur_scale = powf(det3(Tn), (float)(-1.0 / 3.0));         // :250
      const float delta_scale = powf(det3(dT), (float)(1.0 / 3.0));        // :251
      const float delta_tran = sqrtf(dT[3] * dT[3] + dT[7] * dT[7] + dT[11] * dT[11]) * cur_scale;   // :252
      const float trace = (dT[0] + dT[5] + dT[10]) * cur_scale;
      const float delta_rot = fabsf(acosf((trace - 1.f) / 2.f)) * (float)(180.0 / 3.14159265358979323846);   // :253 (NaN when |arg| > 1)
      pose_conv = !a.pose_known[f] && delta_tran < (float)P.eps_t && delta_rot < (float)P.eps_r && delta_scale < (float)P.eps_s;   // :285
    }
    s_st_pose = pose_conv;
  }
  __syncthreads();
  if (tid == 0) {
    int st = 0;
    a.iter_count[f] = a.iter + 1 - a.iter_first;
    const bool guard = a.iter > 1;
    if (s_bmax < (float)P.eps_g && guard) st |= HM_STATUS_CONV_GRADIENT;                         // :276
    else if (!s_cnan && s_cmax < (float)P.eps_c && guard) st |= HM_STATUS_CONV_CODE;             // :280
    else if (a.joint && s_st_pose && guard) st |= HM_STATUS_CONV_POSE;                           // :285
    if (!st && a.iter == a.iter_last) st |= HM_STATUS_MAX_ITER;                                  // :289
    if (st) { atomicOr(&a.status[f], st); a.active[f] = 0; }
  }
}


// The tensor-core decoder marks the fruits (latent-table rows) one of whose rows left the calibrated fp16 range during the loop;
// after the loop the marks become HM_STATUS_F16_SATURATED on exactly those fruits.
__global__ void saturation_status_kernel(const int32_t* __restrict__ fruit_sat, int n_fruits, int32_t* __restrict__ status) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < n_fruits && fruit_sat[f] != 0) atomicOr(&status[f], HM_STATUS_F16_SATURATED);
}

// host helpers ---------------------------------------------------------------------------------
DevParams make_dev_params(const hm_opt_params* p, bool joint) {
  DevParams d;
  d.M = p->n_depth_samples;
  d.log_occ = p->log_sdf_occ;
  d.occlusion_on = p->occlusion_on;
  d.scale_on = p->scale_on;
  d.pose_dim = p->scale_on ? 7 : 6;
  d.est = d.pose_dim + HM_LATENT;
  d.robust_iter = p->robust_iter;
  d.min_valid = p->min_valid_sample;
  d.lm_on = p->lm_on;
  d.lm_eye = p->lm_eye;
  d.th = (float)p->occ_cutoff_m;
  d.sigma = (float)(p->occ_cutoff_m / 3 * 0.55);            // loss.py:59-60 (Python doubles)
  d.inv_2th = (float)(2 * p->occ_cutoff_m);                 // utils.py:132 `/ (2 * th)`
  d.do_ds_lin = (float)(-1. / (2 * p->occ_cutoff_m));       // loss.py:123
  d.occl_th = (float)p->occlusion_th;
  d.min_grad = (float)p->min_grad_thre;
  d.t_depth = (float)p->robust_th_depth;
  d.t_recon = (float)p->robust_th_recon;
  d.eps_g = p->epsilon_g; d.eps_c = p->epsilon_c; d.eps_t = p->epsilon_t; d.eps_r = p->epsilon_r; d.eps_s = p->epsilon_s;
  d.w_recon = p->w_recon; d.w_depth = p->w_depth; d.w_mask = p->w_mask; d.w_codereg = p->w_codereg;
  d.lm_lambda_0 = p->lm_lambda_0; d.s_damp = p->s_damp;
  (void)joint;
  return d;
}

struct Carver {            // bump allocator over the context workspace
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* b) : base((uint8_t*)b) {}
  template <class T> T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? (T*)(base + off) : nullptr;
    off += sizeof(T) * n;
    return p;
  }
};

inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

}  // namespace

// per-call tables built on the device (O(rays) / O(points) work that used to be host loops + uploads)
__global__ void ray_table_kernel(const FrameState* __restrict__ fs, int32_t* __restrict__ ray_frame) {
  const int g = blockIdx.x;
  const int64_t r0 = fs[g].ray_begin;
  const int n = fs[g].n_rays;
  for (int j = threadIdx.x; j < n; j += blockDim.x) ray_frame[r0 + j] = g;
}
__global__ void point_table_kernel(const int64_t* __restrict__ point_offsets, int32_t* __restrict__ point_fruit) {
  const int f = blockIdx.x;
  const int64_t i1 = point_offsets[f + 1];
  for (int64_t i = point_offsets[f] + threadIdx.x; i < i1; i += blockDim.x) point_fruit[i] = f;
}

// Pinned staging arena of the per-call host tables: the uploads are asynchronous on the caller's stream and the arena is only
// rewritten after the previous call's uploads have completed (event), so hm_optimize_* never waits for queued GPU work.
static int stage_reserve(hm_context* ctx, size_t bytes) {
  if (ctx->stage_event) HM_CUDA(cudaEventSynchronize(ctx->stage_event));
  else HM_CUDA(cudaEventCreateWithFlags(&ctx->stage_event, cudaEventDisableTiming));
  if (bytes <= ctx->stage_bytes) return HM_OK;
  if (ctx->h_stage) HM_CUDA(cudaFreeHost(ctx->h_stage));
  ctx->h_stage = nullptr;
  ctx->stage_bytes = 0;
  bytes = (bytes + 65535) & ~size_t(65535);
  HM_CUDA(cudaMallocHost(&ctx->h_stage, bytes));
  ctx->stage_bytes = bytes;
  return HM_OK;
}

int hm_optimize_impl(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* b, bool joint, cudaStream_t st) {
  HM_CHECK(ctx && p && b, "hm_optimize: null argument");
  HM_CHECK(b->n_fruits > 0 && b->d_latents && b->d_T_ow && b->d_points_w && b->h_point_offsets && b->d_iter_count && b->d_status,
           "hm_optimize: incomplete fruit batch");
  HM_CHECK(p->max_iter >= 0, "hm_optimize: max_iter < 0");
  HM_CUDA(cudaSetDevice(ctx->device));
  const int nf = b->n_fruits;
  const int64_t n_points = b->h_point_offsets[nf];
  HM_CHECK(n_points >= 0 && b->h_point_offsets[0] == 0, "hm_optimize: point offsets must start at 0 and be non-decreasing");
  int n_frames = 0;
  int64_t n_rays = 0;
  const int M = p->n_depth_samples;
  if (joint) {
    HM_CHECK(b->h_frame_offsets && b->d_T_wc && b->h_ray_offsets && b->h_n_fg && b->d_rays && b->d_depth_obs && b->h_cube_radius && b->h_pose_known,
             "hm_optimize_joint: incomplete render data");
    HM_CHECK(M >= 2 && M <= kMaxM, "hm_optimize_joint: n_depth_samples must be in [2, %d]", kMaxM);
    n_frames = b->h_frame_offsets[nf];
    n_rays = n_frames ? b->h_ray_offsets[n_frames] : 0;
  }
  const DevParams P = make_dev_params(p, joint);
  const int pose_dim = joint ? P.pose_dim : 0;
  const int est = pose_dim + HM_LATENT;
  const int64_t S = n_rays * M;
  const int64_t Gmax = n_points + S;
  HM_CHECK(Gmax < (int64_t)INT32_MAX - 64, "hm_optimize: %lld decoder rows per iteration exceed the int32 row index of one call; split the batch",
           (long long)Gmax);
  const int n_scan_blocks = (int)((n_rays + 255) / 256);

  // ---- small host-side tables (per frame / per normal-equation block), staged in pinned memory
  int n_blocks = 0;
  for (int f = 0; f < nf; ++f) {
    if (joint && b->h_frame_offsets[f + 1] > b->h_frame_offsets[f]) {
      const int64_t nr = b->h_ray_offsets[b->h_frame_offsets[f + 1]] - b->h_ray_offsets[b->h_frame_offsets[f]];
      n_blocks += 2 * (int)((nr + kRedItems - 1) / kRedItems);
    }
    n_blocks += (int)((b->h_point_offsets[f + 1] - b->h_point_offsets[f] + kRedItems - 1) / kRedItems);
  }
  Carver hs_size(nullptr);
  hs_size.take<FrameState>(n_frames); hs_size.take<RedBlock>(n_blocks); hs_size.take<int32_t>(nf + 1); hs_size.take<int64_t>(nf + 1);
  hs_size.take<float>(nf); hs_size.take<uint8_t>(nf);
  int rc = stage_reserve(ctx, hs_size.off + 256);
  if (rc) return rc;
  Carver hs(ctx->h_stage);
  FrameState* h_fs = hs.take<FrameState>(n_frames);
  RedBlock* h_blocks = hs.take<RedBlock>(n_blocks);
  int32_t* h_fbb = hs.take<int32_t>(nf + 1);
  int64_t* h_po = hs.take<int64_t>(nf + 1);
  float* h_cr = hs.take<float>(nf);
  uint8_t* h_pk = hs.take<uint8_t>(nf);
  int nb = 0;
  for (int f = 0; f < nf; ++f) {
    h_fbb[f] = nb;
    h_po[f] = b->h_point_offsets[f];
    h_cr[f] = joint ? b->h_cube_radius[f] : 0.f;
    h_pk[f] = joint ? b->h_pose_known[f] : 0;
    if (joint) {
      int64_t rb = -1, re = -1;
      for (int g = b->h_frame_offsets[f]; g < b->h_frame_offsets[f + 1]; ++g) {
        FrameState& F = h_fs[g];
        memset(&F, 0, sizeof(F));
        F.fruit = f;
        F.n_fg = b->h_n_fg[g];
        F.ray_begin = b->h_ray_offsets[g];
        F.n_rays = (int32_t)(b->h_ray_offsets[g + 1] - b->h_ray_offsets[g]);
        if (rb < 0) rb = b->h_ray_offsets[g];
        re = b->h_ray_offsets[g + 1];
      }
      for (int term = 0; term < 2; ++term)
        for (int64_t s0 = rb; rb >= 0 && s0 < re; s0 += kRedItems)
          h_blocks[nb++] = {f, term, s0, (int)std::min<int64_t>(kRedItems, re - s0)};
    }
    for (int64_t s0 = b->h_point_offsets[f]; s0 < b->h_point_offsets[f + 1]; s0 += kRedItems)
      h_blocks[nb++] = {f, 2, s0, (int)std::min<int64_t>(kRedItems, b->h_point_offsets[f + 1] - s0)};
  }
  h_fbb[nf] = nb;
  h_po[nf] = n_points;
  HM_CHECK(nb == n_blocks, "hm_optimize: internal block count mismatch");

  // ---- carve the workspace (two passes: size, then pointers)
  struct Ptrs {
    FrameState* fs; int32_t *ray_frame, *point_fruit, *fbb; int64_t* point_offsets; RedBlock* blocks; float* cube_radius; uint8_t *pose_known, *active;
    float *xyz_s, *sdf_s, *coef_e, *coef_m; uint8_t* valid; unsigned long long* ray_mask; float *res_d, *res_m;
    float* xyz_c; int32_t *idx_c, *row_latent_c, *n_valid;
    int32_t *ray_k, *ray_off, *ray_slot, *block_sum, *block_base, *n_rows_dyn;
    float *xyz_g, *sdf_g, *jac_g; int32_t* row_latent_g; float *J_d, *J_m, *J_r, *res_r, *partials; int32_t *block_items, *fruit_sat;
    int32_t *cidx_of, *src_g; uint32_t* masks;
  } w;
  // the forward pass over the ray samples keeps its ReLU bits (512 B per row) so that the gradient of the in-band samples does not
  // need a second forward evaluation (hm_set_mask_reuse)
  const bool reuse = joint && n_rays > 0 && ctx->engine == HM_ENGINE_TC && ctx->mask_reuse;
  auto carve = [&](void* base) {
    Carver c(base);
    w.fs = c.take<FrameState>(n_frames); w.ray_frame = c.take<int32_t>(n_rays); w.point_fruit = c.take<int32_t>(n_points);
    w.fbb = c.take<int32_t>(nf + 1); w.point_offsets = c.take<int64_t>(nf + 1); w.blocks = c.take<RedBlock>(n_blocks);
    w.cube_radius = c.take<float>(nf); w.pose_known = c.take<uint8_t>(nf); w.active = c.take<uint8_t>(nf);
    w.xyz_s = c.take<float>(S * 3); w.sdf_s = c.take<float>(S); w.coef_e = c.take<float>(S); w.coef_m = c.take<float>(S);
    w.valid = c.take<uint8_t>(S); w.ray_mask = c.take<unsigned long long>(n_rays); w.res_d = c.take<float>(n_rays); w.res_m = c.take<float>(n_rays);
    w.xyz_c = c.take<float>(S * 3); w.idx_c = c.take<int32_t>(S); w.row_latent_c = c.take<int32_t>(S); w.n_valid = c.take<int32_t>(4);
    w.ray_k = c.take<int32_t>(n_rays); w.ray_off = c.take<int32_t>(n_rays); w.ray_slot = c.take<int32_t>(n_rays);
    w.block_sum = c.take<int32_t>(n_scan_blocks + 1); w.block_base = c.take<int32_t>(n_scan_blocks + 1); w.n_rows_dyn = c.take<int32_t>(4);
    w.xyz_g = c.take<float>(Gmax * 3); w.sdf_g = c.take<float>(Gmax); w.jac_g = c.take<float>(Gmax * HM_IN); w.row_latent_g = c.take<int32_t>(Gmax);
    w.J_d = c.take<float>(n_rays * kE); w.J_m = c.take<float>(n_rays * kE); w.J_r = c.take<float>(n_points * kE); w.res_r = c.take<float>(n_points);
    w.partials = c.take<float>((size_t)n_blocks * kPartial); w.block_items = c.take<int32_t>(n_blocks); w.fruit_sat = c.take<int32_t>(nf);
    w.cidx_of = c.take<int32_t>(reuse ? S : 0); w.src_g = c.take<int32_t>(reuse ? S : 0);
    w.masks = c.take<uint32_t>(reuse ? (size_t)((S + HM_TC_TILE_M - 1) / HM_TC_TILE_M) * 8 * 1024 : 0);
    return c.off + 256;
  };
  const size_t need = carve(nullptr);
  rc = hm_ws2_reserve(ctx, need);
  if (rc) return rc;
  carve(ctx->ws2);

  // last-system buffers (test hook)
  if (ctx->last_n_fruits < nf) {
    if (ctx->d_last_H) { HM_CUDA(cudaDeviceSynchronize()); cudaFree(ctx->d_last_H); cudaFree(ctx->d_last_b); cudaFree(ctx->d_last_dx); }
    HM_CUDA(cudaMalloc(&ctx->d_last_H, sizeof(float) * (size_t)nf * kE * kE));
    HM_CUDA(cudaMalloc(&ctx->d_last_b, sizeof(float) * (size_t)nf * kE));
    HM_CUDA(cudaMalloc(&ctx->d_last_dx, sizeof(float) * (size_t)nf * kE));
    ctx->last_n_fruits = nf;
  }
  ctx->last_est = est;
  ctx->last_call_fruits = nf;

  // ---- upload the staged tables, build the O(rays) / O(points) tables on the device
  auto up = [&](void* d, const void* h, size_t bytes) -> cudaError_t { return bytes ? cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess; };
  HM_CUDA(up(w.fs, h_fs, sizeof(FrameState) * n_frames));
  HM_CUDA(up(w.blocks, h_blocks, sizeof(RedBlock) * n_blocks));
  HM_CUDA(up(w.fbb, h_fbb, sizeof(int32_t) * (nf + 1)));
  HM_CUDA(up(w.point_offsets, h_po, sizeof(int64_t) * (nf + 1)));
  HM_CUDA(up(w.cube_radius, h_cr, sizeof(float) * nf));
  HM_CUDA(up(w.pose_known, h_pk, nf));
  HM_CUDA(cudaEventRecord(ctx->stage_event, st));
  HM_CUDA(cudaMemsetAsync(w.active, 1, nf, st));
  HM_CUDA(cudaMemsetAsync(w.fruit_sat, 0, sizeof(int32_t) * nf, st));
  HM_CUDA(cudaMemsetAsync(b->d_iter_count, 0, sizeof(int32_t) * nf, st));
  HM_CUDA(cudaMemsetAsync(b->d_status, 0, sizeof(int32_t) * nf, st));
  int64_t launches = 0;
  if (n_frames > 0) { ray_table_kernel<<<n_frames, 128, 0, st>>>(w.fs, w.ray_frame); ++launches; }
  if (n_points > 0) { point_table_kernel<<<nf, 256, 0, st>>>(w.point_offsets, w.point_fruit); ++launches; }

  SolveArgs sa;
  sa.fruit_block_begin = w.fbb; sa.blocks = w.blocks; sa.partials = w.partials; sa.block_items = w.block_items;
  sa.cube_radius = w.cube_radius; sa.pose_known = w.pose_known; sa.latents = b->d_latents; sa.T_ow = b->d_T_ow;
  sa.active = w.active; sa.iter_count = b->d_iter_count; sa.status = b->d_status;
  sa.last_H = ctx->d_last_H; sa.last_b = ctx->d_last_b; sa.last_dx = ctx->d_last_dx;
  sa.joint = joint ? 1 : 0; sa.iter_first = p->iter_offset; sa.iter_last = p->iter_offset + p->max_iter - 1;

  for (int it = p->iter_offset; it < p->iter_offset + p->max_iter; ++it) {
    hm_rows grows = {nullptr, w.xyz_g, b->d_latents, w.row_latent_g, n_points, nullptr};
    grows.d_latent_sat = w.fruit_sat;
    if (joint && n_rays > 0) {
      frame_setup_kernel<<<nblk(n_frames, 64), 64, 0, st>>>(n_frames, w.fs, b->d_T_ow, b->d_T_wc, w.cube_radius, w.active, M, w.n_valid);
      sample_kernel<<<nblk(S, 256), 256, 0, st>>>(S, M, w.fs, w.ray_frame, b->d_rays, w.active, w.xyz_s, w.valid, w.xyz_c, w.idx_c,
                                                  w.row_latent_c, w.n_valid, reuse ? w.cidx_of : nullptr);
      launches += 2;
      // forward pass over the in-sphere samples only (loss.py:47-49); the SDF of compact row j lands at sample idx_c[j]
      hm_rows srows = {nullptr, w.xyz_c, b->d_latents, w.row_latent_c, S, w.n_valid};
      srows.d_out_index = w.idx_c;
      srows.d_latent_sat = w.fruit_sat;
      srows.d_mask_out = reuse ? w.masks : nullptr;
      rc = hm_decode(ctx, srows, w.sdf_s, nullptr, st);
      if (rc) return rc;
      composite_kernel<<<n_scan_blocks, 256, 0, st>>>(n_rays, P, w.fs, w.ray_frame, b->d_depth_obs, w.valid, w.sdf_s, w.active, w.coef_e,
                                                      w.coef_m, w.ray_mask, w.res_d, w.res_m, w.ray_k, w.ray_off, w.block_sum, b->d_status);
      scan_blocks_kernel<<<1, 1024, 0, st>>>(n_scan_blocks, w.block_sum, w.block_base, w.n_rows_dyn, (int32_t)n_points);
      scatter_kernel<<<nblk(n_rays, 256), 256, 0, st>>>(n_rays, M, w.ray_frame, w.fs, w.ray_mask, w.ray_off, w.block_base, w.xyz_s,
                                                        (int32_t)n_points, w.xyz_g, w.row_latent_g, w.ray_slot, w.cidx_of, w.sdf_s,
                                                        reuse ? w.src_g : nullptr, w.sdf_g);
      launches += 3;
      if (!reuse) {
        grows.n = Gmax;
        grows.d_n_dynamic = w.n_rows_dyn;
      }
    }
    if (n_points > 0 && (joint || it == p->iter_offset)) {
      // (the latent-only loop never changes T_ow, optimizer.py:343: its object-frame points are the same in every iteration)
      transform_points_kernel<<<nblk(n_points, 256), 256, 0, st>>>(n_points, w.point_fruit, b->d_points_w, b->d_T_ow, w.xyz_g, w.row_latent_g);
      ++launches;
    }
    rc = hm_decode(ctx, grows, w.sdf_g, w.jac_g, st);       // observed points (+ the in-band samples when the forward pass is not reused)
    if (rc) return rc;
    if (reuse) {
      // in-band samples: gradient only, from the ReLU bits and SDF values of the forward launch above
      hm_rows brows = {nullptr, w.xyz_g + n_points * 3, b->d_latents, w.row_latent_g + n_points, S, w.n_rows_dyn + 1};
      brows.d_latent_sat = w.fruit_sat;
      brows.d_mask_in = w.masks;
      brows.d_src_row = w.src_g;
      rc = hm_decode(ctx, brows, w.sdf_g + n_points, w.jac_g + n_points * HM_IN, st);
      if (rc) return rc;
    }
    if (joint && n_rays > 0) {
      ray_jacobian_kernel<<<nblk(n_rays, 128), 128, 0, st>>>(n_rays, M, pose_dim, w.ray_mask, w.ray_slot, w.coef_e, w.coef_m, w.xyz_g, w.jac_g, w.J_d, w.J_m);
      ++launches;
    }
    if (joint) {
      if (n_points > 0) {
        point_jacobian_kernel<<<nblk(n_points, 128), 128, 0, st>>>(n_points, pose_dim, w.xyz_g, w.sdf_g, w.jac_g, w.res_r, w.J_r);
        ++launches;
      }
      if (n_blocks > 0)
        normal_eq_kernel<<<n_blocks, 128, 0, st>>>(w.blocks, est, it, P, w.J_d, w.J_m, w.res_d, w.res_m, w.ray_k, w.J_r, w.res_r, kE, w.active,
                                                   w.partials, w.block_items);
    } else if (n_blocks > 0) {
      // latent only (optimizer.py:306-429): J = d sdf / d latent = the first 32 columns of the decoder's Jacobian rows and the
      // residual is the SDF itself (loss.py:219-243 with the pose fixed) -- read both where the decoder wrote them
      normal_eq_kernel<<<n_blocks, 128, 0, st>>>(w.blocks, est, it, P, w.J_d, w.J_m, w.res_d, w.res_m, w.ray_k, w.jac_g, w.sdf_g, HM_IN, w.active,
                                                 w.partials, w.block_items);
    }
    sa.iter = it;
    solve_kernel<<<nf, kSolveThreads, 0, st>>>(sa, P);
    launches += 2;
  }
  if (ctx->engine == HM_ENGINE_TC && ctx->d_tc_flags) {
    saturation_status_kernel<<<nblk(nf, 256), 256, 0, st>>>(w.fruit_sat, nf, b->d_status);
    ++launches;
  }
  HM_CUDA(cudaGetLastError());
  ctx->counters.kernel_launches += launches;
  ctx->counters.iterations += p->max_iter;
  return HM_OK;
}

extern "C" int hm_optimize_shape(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* batch, void* stream) {
  HM_CHECK(ctx, "hm_optimize_shape: null context");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  return hm_optimize_impl(ctx, p, batch, false, (cudaStream_t)stream);
}
extern "C" int hm_optimize_joint(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* batch, void* stream) {
  HM_CHECK(ctx, "hm_optimize_joint: null context");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  return hm_optimize_impl(ctx, p, batch, true, (cudaStream_t)stream);
}

extern "C" int hm_get_last_system(hm_context* ctx, int32_t n_fruits, float* d_H, float* d_b, float* d_dx, void* stream) {
  HM_CHECK(ctx && ctx->d_last_H && ctx->last_est > 0, "hm_get_last_system: no optimisation has run");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CHECK(n_fruits > 0 && n_fruits <= ctx->last_call_fruits, "hm_get_last_system: the last call optimised %d fruits, %d requested",
           ctx->last_call_fruits, n_fruits);
  const int est = ctx->last_est, nf = n_fruits;
  cudaStream_t st = (cudaStream_t)stream;
  if (d_H) HM_CUDA(cudaMemcpy2DAsync(d_H, sizeof(float) * est * est, ctx->d_last_H, sizeof(float) * kE * kE, sizeof(float) * est * est, nf, cudaMemcpyDeviceToDevice, st));
  if (d_b) HM_CUDA(cudaMemcpy2DAsync(d_b, sizeof(float) * est, ctx->d_last_b, sizeof(float) * kE, sizeof(float) * est, nf, cudaMemcpyDeviceToDevice, st));
  if (d_dx) HM_CUDA(cudaMemcpy2DAsync(d_dx, sizeof(float) * est, ctx->d_last_dx, sizeof(float) * kE, sizeof(float) * est, nf, cudaMemcpyDeviceToDevice, st));
  return HM_OK;
}

// host-buffer variant: the call bench.py times end to end (H2D of every input, D2H of the results)
static int optimize_host(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* hb, bool joint) {
  HM_CHECK(ctx && p && hb && hb->n_fruits > 0, "hm_optimize_*_host: bad argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)0);
  HM_CUDA(cudaSetDevice(ctx->device));
  const int nf = hb->n_fruits;
  const int64_t n_points = hb->h_point_offsets[nf];
  const int n_frames = joint ? hb->h_frame_offsets[nf] : 0;
  const int64_t n_rays = (joint && n_frames) ? hb->h_ray_offsets[n_frames] : 0;
  size_t bytes = sizeof(float) * ((size_t)nf * 48 + n_points * 3 + (size_t)n_frames * 16 + n_rays * 4) + sizeof(int32_t) * 2 * nf + 4096;
  if (bytes > ctx->io_arena_bytes) {
    if (ctx->io_arena) { HM_CUDA(cudaDeviceSynchronize()); cudaFree(ctx->io_arena); }
    ctx->io_arena = nullptr;
    ctx->io_arena_bytes = 0;
    HM_CUDA(cudaMalloc(&ctx->io_arena, bytes));
    ctx->io_arena_bytes = bytes;
  }
  Carver c(ctx->io_arena);
  hm_fruit_batch db = *hb;
  db.d_latents = c.take<float>((size_t)nf * 32);
  db.d_T_ow = c.take<float>((size_t)nf * 16);
  float* d_pts = c.take<float>(n_points * 3);
  float* d_Twc = c.take<float>((size_t)n_frames * 16);
  float* d_rays = c.take<float>(n_rays * 3);
  float* d_dobs = c.take<float>(n_rays);
  db.d_iter_count = c.take<int32_t>(nf);
  db.d_status = c.take<int32_t>(nf);
  db.d_points_w = d_pts; db.d_T_wc = d_Twc; db.d_rays = d_rays; db.d_depth_obs = d_dobs;
  cudaStream_t st = 0;
  HM_CUDA(cudaMemcpyAsync(db.d_latents, hb->d_latents, sizeof(float) * nf * 32, cudaMemcpyHostToDevice, st));
  HM_CUDA(cudaMemcpyAsync(db.d_T_ow, hb->d_T_ow, sizeof(float) * nf * 16, cudaMemcpyHostToDevice, st));
  HM_CUDA(cudaMemcpyAsync(d_pts, hb->d_points_w, sizeof(float) * n_points * 3, cudaMemcpyHostToDevice, st));
  if (joint && n_frames) {
    HM_CUDA(cudaMemcpyAsync(d_Twc, hb->d_T_wc, sizeof(float) * n_frames * 16, cudaMemcpyHostToDevice, st));
    HM_CUDA(cudaMemcpyAsync(d_rays, hb->d_rays, sizeof(float) * n_rays * 3, cudaMemcpyHostToDevice, st));
    HM_CUDA(cudaMemcpyAsync(d_dobs, hb->d_depth_obs, sizeof(float) * n_rays, cudaMemcpyHostToDevice, st));
  }
  int rc = hm_optimize_impl(ctx, p, &db, joint, st);
  if (rc) return rc;
  HM_CUDA(cudaMemcpyAsync(hb->d_latents, db.d_latents, sizeof(float) * nf * 32, cudaMemcpyDeviceToHost, st));
  HM_CUDA(cudaMemcpyAsync(hb->d_T_ow, db.d_T_ow, sizeof(float) * nf * 16, cudaMemcpyDeviceToHost, st));
  HM_CUDA(cudaMemcpyAsync(hb->d_iter_count, db.d_iter_count, sizeof(int32_t) * nf, cudaMemcpyDeviceToHost, st));
  HM_CUDA(cudaMemcpyAsync(hb->d_status, db.d_status, sizeof(int32_t) * nf, cudaMemcpyDeviceToHost, st));
  HM_CUDA(cudaStreamSynchronize(st));
  return HM_OK;
}

extern "C" int hm_optimize_shape_host(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* host_batch) {
  return optimize_host(ctx, p, host_batch, false);
}
extern "C" int hm_optimize_joint_host(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* host_batch) {
  return optimize_host(ctx, p, host_batch, true);
}

// ---------------------------------------------------------------------------------------------
// Single-call loss terms with the reference's own signatures (used by the parity tests and by a
// host that keeps the reference's Python loop)
// ---------------------------------------------------------------------------------------------
extern "C" int hm_sdf_loss(hm_context* ctx, const float* d_latent, const float* d_pts_obj, int64_t n, int32_t scale_on,
                           float* d_res, float* d_J_pose, float* d_J_code, void* stream) {
  HM_CHECK(ctx && d_latent && d_pts_obj && d_res && d_J_pose && d_J_code && n > 0, "hm_sdf_loss: bad argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int pose_dim = scale_on ? 7 : 6;
  Carver sz(nullptr);
  sz.take<float>(n); sz.take<float>(n * HM_IN); sz.take<float>(n * kE);
  int rc = hm_ws2_reserve(ctx, sz.off + 1024);
  if (rc) return rc;
  Carver c(ctx->ws2);
  float* sdf = c.take<float>(n);
  float* jac = c.take<float>(n * HM_IN);
  float* J = c.take<float>(n * kE);
  hm_rows rows = {nullptr, d_pts_obj, d_latent, nullptr, n, nullptr};
  rc = hm_decode(ctx, rows, sdf, jac, st);
  if (rc) return rc;
  point_jacobian_kernel<<<nblk(n, 128), 128, 0, st>>>(n, pose_dim, d_pts_obj, sdf, jac, d_res, J);
  ctx->counters.kernel_launches += 1;
  HM_CUDA(cudaMemcpy2DAsync(d_J_pose, sizeof(float) * pose_dim, J, sizeof(float) * kE, sizeof(float) * pose_dim, n, cudaMemcpyDeviceToDevice, st));
  HM_CUDA(cudaMemcpy2DAsync(d_J_code, sizeof(float) * HM_LATENT, J + pose_dim, sizeof(float) * kE, sizeof(float) * HM_LATENT, n, cudaMemcpyDeviceToDevice, st));
  HM_CUDA(cudaGetLastError());
  return HM_OK;
}

extern "C" int hm_render_loss(hm_context* ctx, const hm_opt_params* p, const float* d_latent, const float* d_rays, int32_t n_rays,
                              int32_t n_fg, const float* d_depth_obs, const float* h_T_oc, const float* h_depths,
                              float bbx_radius, int32_t* d_ray_valid, float* d_res_d, float* d_J_d, float* d_res_m,
                              float* d_J_m, int32_t* h_n_valid_samples, void* stream) {
  HM_CHECK(ctx && p && d_latent && d_rays && d_depth_obs && h_T_oc && h_depths && d_ray_valid && d_res_d && d_J_d && d_res_m && d_J_m,
           "hm_render_loss: null argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  const int M = p->n_depth_samples;
  HM_CHECK(n_rays > 0 && n_fg >= 0 && n_fg <= n_rays && M >= 2 && M <= kMaxM, "hm_render_loss: bad sizes");
  HM_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const DevParams P = make_dev_params(p, true);
  const int est = P.est;
  const int64_t S = (int64_t)n_rays * M;
  const int n_scan_blocks = (n_rays + 255) / 256;
  struct { FrameState* fs; int32_t *ray_frame, *status, *ray_k, *ray_off, *ray_slot, *block_sum, *block_base, *n_rows_dyn, *row_latent_g;
           int32_t *idx_c, *row_latent_c, *n_valid; uint8_t *active, *valid;
           float *xyz_s, *xyz_c, *sdf_s, *coef_e, *coef_m, *xyz_g, *sdf_g, *jac_g, *Jd, *Jm; unsigned long long* ray_mask; } w;
  auto carve = [&](void* base) {
    Carver c(base);
    w.fs = c.take<FrameState>(1); w.ray_frame = c.take<int32_t>(n_rays); w.status = c.take<int32_t>(4); w.active = c.take<uint8_t>(4);
    w.xyz_s = c.take<float>(S * 3); w.valid = c.take<uint8_t>(S); w.sdf_s = c.take<float>(S); w.coef_e = c.take<float>(S); w.coef_m = c.take<float>(S);
    w.xyz_c = c.take<float>(S * 3); w.idx_c = c.take<int32_t>(S); w.row_latent_c = c.take<int32_t>(S); w.n_valid = c.take<int32_t>(4);
    w.ray_mask = c.take<unsigned long long>(n_rays); w.ray_k = c.take<int32_t>(n_rays); w.ray_off = c.take<int32_t>(n_rays); w.ray_slot = c.take<int32_t>(n_rays);
    w.block_sum = c.take<int32_t>(n_scan_blocks + 1); w.block_base = c.take<int32_t>(n_scan_blocks + 1); w.n_rows_dyn = c.take<int32_t>(4);
    w.xyz_g = c.take<float>(S * 3); w.sdf_g = c.take<float>(S); w.jac_g = c.take<float>(S * HM_IN); w.row_latent_g = c.take<int32_t>(S);
    w.Jd = c.take<float>((size_t)n_rays * kE); w.Jm = c.take<float>((size_t)n_rays * kE);
    return c.off + 256;
  };
  int rc = hm_ws2_reserve(ctx, carve(nullptr));
  if (rc) return rc;
  carve(ctx->ws2);
  FrameState F;
  memset(&F, 0, sizeof(F));
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) F.A[i * 3 + j] = h_T_oc[i * 4 + j]; F.t[i] = h_T_oc[i * 4 + 3]; }
  F.rho = bbx_radius;
  for (int m = 0; m < M; ++m) F.depths[m] = h_depths[m];
  F.delta_d = (F.depths[M - 1] - F.depths[0]) / (float)(M - 1);      // loss.py:73-75
  F.d_term = F.depths[M - 1] + F.delta_d;                             // loss.py:78
  F.fruit = 0; F.n_fg = n_fg; F.ray_begin = 0; F.n_rays = n_rays; F.valid_count = 0;
  HM_CUDA(cudaMemcpyAsync(w.fs, &F, sizeof(F), cudaMemcpyHostToDevice, st));
  HM_CUDA(cudaMemsetAsync(w.ray_frame, 0, sizeof(int32_t) * n_rays, st));
  HM_CUDA(cudaMemsetAsync(w.status, 0, 16, st));
  HM_CUDA(cudaMemsetAsync(w.active, 1, 4, st));
  HM_CUDA(cudaMemsetAsync(w.n_valid, 0, 16, st));
  HM_CUDA(cudaStreamSynchronize(st));
  sample_kernel<<<nblk(S, 256), 256, 0, st>>>(S, M, w.fs, w.ray_frame, d_rays, w.active, w.xyz_s, w.valid, w.xyz_c, w.idx_c, w.row_latent_c, w.n_valid, nullptr);
  hm_rows srows = {nullptr, w.xyz_c, d_latent, nullptr, S, w.n_valid};      // in-sphere samples only (loss.py:47-49)
  srows.d_out_index = w.idx_c;
  rc = hm_decode(ctx, srows, w.sdf_s, nullptr, st);
  if (rc) return rc;
  composite_kernel<<<n_scan_blocks, 256, 0, st>>>(n_rays, P, w.fs, w.ray_frame, d_depth_obs, w.valid, w.sdf_s, w.active, w.coef_e, w.coef_m,
                                                  w.ray_mask, d_res_d, d_res_m, w.ray_k, w.ray_off, w.block_sum, w.status);
  scan_blocks_kernel<<<1, 1024, 0, st>>>(n_scan_blocks, w.block_sum, w.block_base, w.n_rows_dyn, 0);
  scatter_kernel<<<nblk(n_rays, 256), 256, 0, st>>>(n_rays, M, w.ray_frame, w.fs, w.ray_mask, w.ray_off, w.block_base, w.xyz_s, 0, w.xyz_g,
                                                    w.row_latent_g, w.ray_slot, nullptr, nullptr, nullptr, nullptr);
  hm_rows grows = {nullptr, w.xyz_g, d_latent, nullptr, S, w.n_rows_dyn};
  rc = hm_decode(ctx, grows, w.sdf_g, w.jac_g, st);
  if (rc) return rc;
  ray_jacobian_kernel<<<nblk(n_rays, 128), 128, 0, st>>>(n_rays, M, P.pose_dim, w.ray_mask, w.ray_slot, w.coef_e, w.coef_m, w.xyz_g, w.jac_g, w.Jd, w.Jm);
  ctx->counters.kernel_launches += 5;
  HM_CUDA(cudaMemcpy2DAsync(d_J_d, sizeof(float) * est, w.Jd, sizeof(float) * kE, sizeof(float) * est, n_rays, cudaMemcpyDeviceToDevice, st));
  HM_CUDA(cudaMemcpy2DAsync(d_J_m, sizeof(float) * est, w.Jm, sizeof(float) * kE, sizeof(float) * est, n_rays, cudaMemcpyDeviceToDevice, st));
  HM_CUDA(cudaMemcpyAsync(d_ray_valid, w.ray_k, sizeof(int32_t) * n_rays, cudaMemcpyDeviceToDevice, st));
  FrameState Fo;
  HM_CUDA(cudaMemcpyAsync(&Fo, w.fs, sizeof(Fo), cudaMemcpyDeviceToHost, st));
  HM_CUDA(cudaStreamSynchronize(st));
  HM_CUDA(cudaGetLastError());
  if (h_n_valid_samples) *h_n_valid_samples = Fo.valid_count;
  return HM_OK;
}

#ifdef HM_TESTING
// ---------------------------------------------------------------------------------------------
// TEST-ONLY exports (libhortimapping_b200_testing.so): the device functions of the LM step evaluated on their own, so that the
// tests can hold them against the reference's golden vectors directly (tests/golden/misc.npz: exp_sim3 / exp_se3 incl. the
// quirk branches of utils.py:279-324, Huber weights incl. the exact-zero case of utils.py:327-358).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void debug_exp_pose_kernel(const float* __restrict__ x, int n, int pose_dim, float* __restrict__ T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) exp_pose(x + (size_t)i * 7, pose_dim, T + (size_t)i * 16);
}
__global__ void debug_huber_kernel(const float* __restrict__ r, int n, float b, float* __restrict__ w2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w2[i] = huber_w2(r[i], b);
}
}  // namespace

// h_x [n][7] (translation, rotation, log-scale; the last entry is ignored for pose_dim 6) -> h_T [n][16]
extern "C" int hm_debug_exp_pose(hm_context* ctx, const float* h_x, int n, int pose_dim, float* h_T) {
  HM_CHECK(ctx && h_x && h_T && n > 0 && (pose_dim == 6 || pose_dim == 7), "hm_debug_exp_pose: bad argument");
  HM_CUDA(cudaSetDevice(ctx->device));
  float *dx = nullptr, *dT = nullptr;
  HM_CUDA(cudaMalloc(&dx, sizeof(float) * 7 * n));
  HM_CUDA(cudaMalloc(&dT, sizeof(float) * 16 * n));
  HM_CUDA(cudaMemcpy(dx, h_x, sizeof(float) * 7 * n, cudaMemcpyHostToDevice));
  debug_exp_pose_kernel<<<(n + 63) / 64, 64>>>(dx, n, pose_dim, dT);
  HM_CUDA(cudaGetLastError());
  HM_CUDA(cudaMemcpy(h_T, dT, sizeof(float) * 16 * n, cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dT);
  return HM_OK;
}

// squared Huber weights w^2 of residuals h_r [n] with threshold b (what optimizer.py:145-149 multiplies J^T J and J^T r with)
extern "C" int hm_debug_huber_w2(hm_context* ctx, const float* h_r, int n, float b, float* h_w2) {
  HM_CHECK(ctx && h_r && h_w2 && n > 0, "hm_debug_huber_w2: bad argument");
  HM_CUDA(cudaSetDevice(ctx->device));
  float *dr = nullptr, *dw = nullptr;
  HM_CUDA(cudaMalloc(&dr, sizeof(float) * n));
  HM_CUDA(cudaMalloc(&dw, sizeof(float) * n));
  HM_CUDA(cudaMemcpy(dr, h_r, sizeof(float) * n, cudaMemcpyHostToDevice));
  debug_huber_kernel<<<(n + 63) / 64, 64>>>(dr, n, b, dw);
  HM_CUDA(cudaGetLastError());
  HM_CUDA(cudaMemcpy(h_w2, dw, sizeof(float) * n, cudaMemcpyDeviceToHost));
  cudaFree(dr); cudaFree(dw);
  return HM_OK;
}
#endif  // HM_TESTING
