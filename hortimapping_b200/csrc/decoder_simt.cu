// fp32 CUDA-core decoder: layer-by-layer tiled SGEMM with fused epilogues.
//
// This is the VALIDATION / CALIBRATION engine (HM_ENGINE_SIMT): plain fp32 FMA arithmetic like the
// reference's cuBLAS sgemm path, used (a) to calibrate the fp16 operand scales of the tensor-core
// engine and (b) in the GPU tests as an independent device implementation.  The product engine is
// decoder_tc.cu.
//
// Restates deepsdf/networks/deep_sdf_decoder.py:75-110 (forward) and the input gradient that
// wild_completion/utils.py:112-122 obtains through autograd.
#include "common.cuh"

namespace {

enum Epi {
  EPI_RELU_BIAS = 1,      // C = relu(acc + bias)
  EPI_RELU_BIAS_SKIP = 2, // lin3: col < 477 relu(acc + bias); col >= 477 copy of the raw input row
  EPI_MASK = 3,           // backward: C = acc * (aux > 0)
  EPI_MASK_SKIP = 4,      // backward through lin4: col < 477 masked, col >= 477 kept (skip gradient)
  EPI_FINAL = 5           // backward through lin0: C[row][col] = acc + aux[row][477 + col], col < 35
};

constexpr int BM = 64, BN = 64, BK = 16;

// C[n x N] = epi(A[n x K] * B), B given as W[N][K] (B_NK, forward) or W[K][N] (backward).
template <bool B_NK>
__global__ void __launch_bounds__(256) simt_gemm(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                 int ldb, const float* __restrict__ bias, float* __restrict__ C,
                                                 int ldc, int64_t n_rows, int N, int K, int epi,
                                                 const float* __restrict__ aux, int ld_aux) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int64_t row0 = (int64_t)blockIdx.y * BM;
  const int col0 = blockIdx.x * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = threadIdx.x; i < BM * BK; i += 256) {
      int r = i / BK, k = i % BK;
      int64_t gr = row0 + r;
      As[k][r] = (gr < n_rows && k0 + k < K) ? A[gr * lda + k0 + k] : 0.f;
    }
    if (B_NK) {
      for (int i = threadIdx.x; i < BN * BK; i += 256) {
        int c = i / BK, k = i % BK;
        Bs[k][c] = (col0 + c < N && k0 + k < K) ? B[(int64_t)(col0 + c) * ldb + k0 + k] : 0.f;
      }
    } else {
      for (int i = threadIdx.x; i < BN * BK; i += 256) {
        int k = i / BN, c = i % BN;
        Bs[k][c] = (col0 + c < N && k0 + k < K) ? B[(int64_t)(k0 + k) * ldb + col0 + c] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = row0 + ty * 4 + i;
    if (r >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = col0 + tx * 4 + j;
      if (c >= N) continue;
      float v = acc[i][j];
      switch (epi) {
        case EPI_RELU_BIAS: v = fmaxf(v + bias[c], 0.f); break;
        case EPI_RELU_BIAS_SKIP:
          v = (c < HM_SKIP_COL) ? fmaxf(v + bias[c], 0.f) : aux[r * ld_aux + (c - HM_SKIP_COL)];
          break;
        case EPI_MASK: v = (aux[r * ld_aux + c] > 0.f) ? v : 0.f; break;
        case EPI_MASK_SKIP: v = (c < HM_SKIP_COL) ? ((aux[r * ld_aux + c] > 0.f) ? v : 0.f) : v; break;
        case EPI_FINAL: v = v + aux[r * ld_aux + HM_SKIP_COL + c]; break;
      }
      C[r * ldc + c] = v;
    }
  }
}

// rows35[n][35] from xyz + latent table
__global__ void build_rows(const float* __restrict__ xyz, const float* __restrict__ latents,
                           const int32_t* __restrict__ row_latent, int64_t n, float* __restrict__ rows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * HM_IN) return;
  int64_t r = i / HM_IN;
  int c = (int)(i % HM_IN);
  int li = row_latent ? row_latent[r] : 0;
  rows[i] = (c < HM_LATENT) ? latents[(int64_t)li * HM_LATENT + c] : xyz[r * 3 + (c - HM_LATENT)];
}

// sdf = tanh(h7 . w8 + b8) (deep_sdf_decoder.py:107-108); optionally d7 = (1 - sdf^2) * w8 * (h7 > 0)
// out_index (optional): the SDF of row r goes to sdf[out_index[r]].  unit_seed: d7 without the (1 - sdf^2) factor -- the operand the
// tensor-core engine feeds to its backward GEMMs (it applies the factor to the finished gradient), used when calibrating its scales.
__global__ void head_kernel(const float* __restrict__ h7, const float* __restrict__ w8, const float* __restrict__ b8,
                            int64_t n, float* __restrict__ sdf, const int32_t* __restrict__ out_index, float* __restrict__ d7, int unit_seed) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= n) return;
  const float* h = h7 + row * HM_HIDDEN;
  float s = 0.f;
  for (int c = lane; c < HM_HIDDEN; c += 32) s = fmaf(h[c], w8[c], s);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float f = tanhf(s + b8[0]);
  if (lane == 0) sdf[out_index ? (int64_t)out_index[row] : row] = f;
  if (d7) {
    float coef = unit_seed ? 1.f : 1.f - f * f;
    for (int c = lane; c < HM_HIDDEN; c += 32) d7[row * HM_HIDDEN + c] = (h[c] > 0.f) ? coef * w8[c] : 0.f;
  }
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (threadIdx.x % 32 == 0) atomicMax((int*)out, __float_as_int(m));   // non-negative floats order as ints
}

// per-column maximum of a row-major [n][ncol] matrix of non-negative values (post-ReLU activations): which hidden units were ever alive
__global__ void colmax_kernel(const float* __restrict__ x, int64_t n, int ncol, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  float m = 0.f;
  for (int64_t r = blockIdx.y; r < n; r += gridDim.y) m = fmaxf(m, x[r * ncol + c]);
  atomicMax((int*)(out + c), __float_as_int(m));   // non-negative floats order as ints
}

template <bool B_NK>
void launch_gemm(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int64_t n,
                 int N, int K, int epi, const float* aux, int ld_aux, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (unsigned)((n + BM - 1) / BM));
  simt_gemm<B_NK><<<grid, 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, n, N, K, epi, aux, ld_aux);
}

}  // namespace

int hm_simt_decode(hm_context* ctx, const hm_rows& rows, float* d_sdf, float* d_jac, cudaStream_t st,
                   float* h_absmax_out, float* h_unit_max_out) {
  const int64_t CH = 16384;
  const bool want_jac = d_jac != nullptr || h_absmax_out != nullptr;
  // workspace: rows35 [CH][35] (padded to 36), h[8][CH][512], d ping-pong [2][CH][512], absmax[16]
  size_t bytes = sizeof(float) * (CH * 36 + (size_t)8 * CH * HM_HIDDEN + (size_t)2 * CH * HM_HIDDEN + 64 + CH * HM_IN + 8 * HM_HIDDEN);
  int rc = hm_ws_reserve(ctx, bytes);
  if (rc) return rc;
  float* w = (float*)ctx->ws;
  float* rows35 = w; w += CH * 36;
  float* h[8];
  for (int l = 0; l < 8; ++l) { h[l] = w; w += CH * HM_HIDDEN; }
  float* dA = w; w += CH * HM_HIDDEN;
  float* dB = w; w += CH * HM_HIDDEN;
  float* amax = w; w += 64;
  float* jac_tmp = w; w += CH * HM_IN;
  float* unit_max = w;                   // [8][512]: per hidden unit, the largest activation seen (calibration only)
  if (h_absmax_out) HM_CUDA(cudaMemsetAsync(amax, 0, 64 * sizeof(float), st));
  if (h_unit_max_out) HM_CUDA(cudaMemsetAsync(unit_max, 0, 8 * HM_HIDDEN * sizeof(float), st));
  int64_t n_total = rows.n;
  if (rows.d_n_dynamic) {   // validation engine: a host read-back of the device-side row count is acceptable here
    int32_t nd = 0;
    HM_CUDA(cudaMemcpyAsync(&nd, rows.d_n_dynamic, sizeof(nd), cudaMemcpyDeviceToHost, st));
    HM_CUDA(cudaStreamSynchronize(st));
    n_total = nd < rows.n ? nd : rows.n;
  }
  if (!h_absmax_out) { if (d_jac) ctx->counters.rows_jacobian += n_total; else ctx->counters.rows_forward += n_total; }

  for (int64_t r0 = 0; r0 < n_total; r0 += CH) {
    int64_t n = n_total - r0 < CH ? n_total - r0 : CH;
    const float* x0;
    if (rows.d_rows) {
      x0 = rows.d_rows + r0 * HM_IN;
    } else {
      int64_t tot = n * HM_IN;
      build_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(rows.d_xyz + r0 * 3, rows.d_latents,
                                                               rows.d_row_latent ? rows.d_row_latent + r0 : nullptr, n, rows35);
      x0 = rows35;
    }
    auto amx = [&](const float* p, int64_t cnt, int slot) {
      if (h_absmax_out) absmax_kernel<<<256, 256, 0, st>>>(p, cnt, amax + slot);
    };
    amx(x0, n * HM_IN, 0);
    // forward (deep_sdf_decoder.py:85-105)
    launch_gemm<true>(x0, HM_IN, ctx->d_W[0], HM_IN, ctx->d_b[0], h[0], HM_HIDDEN, n, HM_HIDDEN, HM_IN, EPI_RELU_BIAS, nullptr, 0, st);
    for (int l = 1; l < 8; ++l) {
      int epi = (l == 3) ? EPI_RELU_BIAS_SKIP : EPI_RELU_BIAS;
      amx(h[l - 1], n * HM_HIDDEN, l);
      launch_gemm<true>(h[l - 1], HM_HIDDEN, ctx->d_W[l], HM_HIDDEN, ctx->d_b[l], h[l], HM_HIDDEN, n, HM_HIDDEN, HM_HIDDEN,
                        epi, x0, HM_IN, st);
    }
    if (h_unit_max_out)
      for (int l = 0; l < 8; ++l) colmax_kernel<<<dim3(HM_HIDDEN / 128, 64), 128, 0, st>>>(h[l], n, HM_HIDDEN, unit_max + l * HM_HIDDEN);
    head_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(h[7], ctx->d_W[8], ctx->d_b[8], n, rows.d_out_index ? d_sdf : d_sdf + r0,
                                                         rows.d_out_index ? rows.d_out_index + r0 : nullptr, want_jac ? dA : nullptr,
                                                         h_absmax_out ? 1 : 0);
    if (!want_jac) continue;
    // backward to the input (what autograd.grad computes for utils.py:112-122)
    float* cur = dA;
    float* nxt = dB;
    for (int l = 7; l >= 1; --l) {
      amx(cur, n * HM_HIDDEN, 8 + (7 - l));
      int epi = (l == 4) ? EPI_MASK_SKIP : EPI_MASK;
      launch_gemm<false>(cur, HM_HIDDEN, ctx->d_W[l], HM_HIDDEN, nullptr, nxt, HM_HIDDEN, n, HM_HIDDEN, HM_HIDDEN, epi,
                         h[l - 1], HM_HIDDEN, st);
      if (l == 4) {
        // keep the skip gradient (cols 477..511 of d3c) alive in h[7] (no longer needed) for EPI_FINAL
        HM_CUDA(cudaMemcpyAsync(h[7], nxt, sizeof(float) * n * HM_HIDDEN, cudaMemcpyDeviceToDevice, st));
      }
      float* t = cur; cur = nxt; nxt = t;
    }
    amx(cur, n * HM_HIDDEN, 15);
    float* jout = d_jac ? d_jac + r0 * HM_IN : jac_tmp;
    launch_gemm<false>(cur, HM_HIDDEN, ctx->d_W[0], HM_IN, nullptr, jout, HM_IN, n, HM_IN, HM_HIDDEN, EPI_FINAL, h[7],
                       HM_HIDDEN, st);
  }
  HM_CUDA(cudaGetLastError());
  if (h_unit_max_out) HM_CUDA(cudaMemcpyAsync(h_unit_max_out, unit_max, 8 * HM_HIDDEN * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (h_absmax_out) {
    HM_CUDA(cudaMemcpyAsync(h_absmax_out, amax, 16 * sizeof(float), cudaMemcpyDeviceToHost, st));
    HM_CUDA(cudaStreamSynchronize(st));
  }
  return HM_OK;
}
