// TEST-ONLY translation unit: bring-up probes and micro-benchmarks of the tensor-core building blocks, and the read-out of the
// instrumentation of the decoder kernel (timeline trace, wait-cycle counters).  Compiled only into
// libhortimapping_b200_testing.so (hortimapping_b200/build.py: the product sources with -DHM_TESTING plus csrc/testing/*.cu);
// the product library exports none of this.  scripts/probe_*.py are the drivers.
#include <algorithm>
#include <vector>

#include "../common.cuh"
#include "../tc_ptx.cuh"

using namespace hm_tc;

namespace {

// One 64 x 128 x 64 GEMM through exactly the building blocks above (SW128 K-major descriptors, M = 64
// accumulator layout with an optional +16 lane offset, 32x32b TMEM loads); dumps all 128 lanes x 256
// columns so the host can check the layout assumptions.
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ B,
                                                             float* __restrict__ out, int lane_off, int col_off, int repeats) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t done_bar;
  const int warp = threadIdx.x >> 5;
  const uint32_t sa = smem_u32(smem), sb = sa + 8192;
  for (int i = threadIdx.x; i < 64 * 64; i += 128) *reinterpret_cast<__half*>(smem + sw128_offset(i / 64, i % 64)) = A[i];
  for (int i = threadIdx.x; i < 128 * 64; i += 128) *reinterpret_cast<__half*>(smem + 8192 + sw128_offset(i / 64, i % 64)) = B[i];
  if (threadIdx.x == 0) { mbar_init(smem_u32(&done_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc<1>(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  // clear the accumulator region we are going to dump
  {
    uint32_t z = 0;
    for (int c = 0; c < 256; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tb + ((uint32_t)(32 * warp) << 16) + c), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(64, 128);
    const uint32_t d = tb + ((uint32_t)lane_off << 16) + col_off;
    for (int rep = 0; rep < repeats; ++rep)
      for (int ks = 0; ks < 4; ++ks) umma_f16<1>(d, make_desc(sa + ks * 32), make_desc(sb + ks * 32), idesc, (rep | ks) ? 1u : 0u);
    umma_commit<1>(smem_u32(&done_bar));
  }
  mbar_wait(smem_u32(&done_bar), 0);
  tc_fence_after();
  for (int q = 0; q < 8; ++q) {
    float v[32];
    tmem_ld32(tb + ((uint32_t)(32 * warp) << 16) + q * 32, v);
    for (int i = 0; i < 32; ++i) out[(size_t)threadIdx.x * 256 + q * 32 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tb, 512); }
}

// MMA issue-rate microbenchmark: `reps` x 4 chained MMAs of shape M x N x 16 from shared memory; returns cycles.
__global__ void __launch_bounds__(128, 1) tc_mma_rate_kernel(int M, int N, int reps, int n_acc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t done_bar;
  const int warp = threadIdx.x >> 5;
  const uint32_t sa = smem_u32(smem), sb = sa + 16384;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(smem_u32(&done_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc<1>(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(M, N);
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      const uint32_t d = tb + (uint32_t)((rep % n_acc) * N) % 512;
      for (int ks = 0; ks < 4; ++ks) umma_f16<1>(d, make_desc(sa + ks * 32), make_desc(sb + ks * 32), idesc, (rep >= n_acc || ks) ? 1u : 0u);
    }
    long long t1 = clock64();
    umma_commit<1>(smem_u32(&done_bar));
    mbar_wait(smem_u32(&done_bar), 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tb, 512); }
}

// CTA-pair probe (cta_group::2): each CTA fills `m_rows` rows of A (64-wide K, SW128) and N/2 rows of B from global memory,
// the leader issues `reps` x 4 chained MMAs of shape M x N x 16 (M = 2 * m_rows) and both CTAs dump their 128 lanes x 256
// TMEM columns.  Answers: where does the accumulator of an M = 128 pair MMA (64 rows per CTA) live, and how long does it take?
__global__ void __launch_bounds__(128, 1) tc_pair_probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ out,
                                                               long long* __restrict__ cycles, int m_rows, int N, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t done_bar;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const uint32_t sa = smem_u32(smem), sb = sa + 16384;
  const int nb = N / 2;
  for (int i = threadIdx.x; i < m_rows * 64; i += 128) *reinterpret_cast<__half*>(smem + sw128_offset(i / 64, i % 64)) = A[(size_t)rank * m_rows * 64 + i];
  for (int i = threadIdx.x; i < nb * 64; i += 128) *reinterpret_cast<__half*>(smem + 16384 + sw128_offset(i / 64, i % 64)) = B[(size_t)rank * nb * 64 + i];
  if (threadIdx.x == 0) { mbar_init(smem_u32(&done_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc<2>(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  {
    uint32_t z = 0x7fc00000u;          // NaN marker: untouched cells stay recognisable
    for (int c = 0; c < 256; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tb + ((uint32_t)(32 * warp) << 16) + c), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(2 * m_rows, N);
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int ks = 0; ks < 4; ++ks) umma_f16<2>(tb, make_desc(sa + ks * 32), make_desc(sb + ks * 32), idesc, (rep | ks) ? 1u : 0u);
    const long long t1 = clock64();
    umma_commit<2>(smem_u32(&done_bar));
    mbar_wait(smem_u32(&done_bar), 0);
    cycles[0] = t1 - t0;
    cycles[1] = clock64() - t0;
  } else {
    mbar_wait(smem_u32(&done_bar), 0);
  }
  __syncthreads();
  tc_fence_after();
  for (int q = 0; q < 8; ++q) {
    float v[32];
    tmem_ld32(tb + ((uint32_t)(32 * warp) << 16) + q * 32, v);
    for (int i = 0; i < 32; ++i) out[((size_t)rank * 128 + threadIdx.x) * 256 + q * 32 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<2>(tb, 512); }
}

// L2 -> shared-memory ingest microbenchmark: every CTA streams `n_stages` stages of `bytes` through a 3-slot ring
// (no MMAs; the consumer frees a slot as soon as it is full), unicast or multicast over the cluster.
__global__ void __launch_bounds__(64, 1) tc_ingest_kernel(const uint8_t* blob, int64_t blob_bytes, int n_stages, uint32_t bytes, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars_[6];
  const uint32_t csize = cluster_nctarank(), crank = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  const uint32_t sbase = smem_u32(smem), b0 = smem_u32(bars_);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(b0 + 8 * i, 1); mbar_init(b0 + 8 * (3 + i), csize); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    uint32_t slot = 0, phase = 0;
    const uint32_t part = bytes / csize;
    int64_t off = ((int64_t)(blockIdx.x / csize) * 7919 * bytes) % (blob_bytes - bytes);
    off &= ~int64_t(1023);
    for (int s = 0; s < n_stages; ++s) {
      mbar_wait(b0 + 8 * (3 + slot), phase ^ 1);
      mbar_expect_tx(b0 + 8 * slot, bytes);
      if (csize == 1) bulk_g2s(sbase + slot * bytes, blob + off, bytes, b0 + 8 * slot);
      else bulk_g2s_multicast(sbase + slot * bytes + crank * part, blob + off + crank * part, part, b0 + 8 * slot, cmask);
      off += bytes;
      if (off + bytes > blob_bytes) off = 0;
      if (++slot == 3) { slot = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    uint32_t slot = 0, phase = 0;
    for (int s = 0; s < n_stages; ++s) {
      mbar_wait(b0 + 8 * slot, phase);
      if (csize == 1) mbar_arrive(b0 + 8 * (3 + slot));
      else {
        for (uint32_t r = 0; r < csize; ++r) {
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(b0 + 8 * (3 + slot)), "r"(r));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
        }
      }
      if (++slot == 3) { slot = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

}  // namespace

// Debug export (not part of the public header): A [64][64] and B [128][64] fp16 bit patterns (host),
// out [128][256] fp32 (host) = raw TMEM dump after D = A * B^T was issued at (lane_off, col_off).
extern "C" int hm_debug_tc_selftest(hm_context* ctx, const uint16_t* h_A, const uint16_t* h_B, float* h_out,
                                    int lane_off, int col_off, int repeats) {
  HM_CHECK(ctx && h_A && h_B && h_out, "hm_debug_tc_selftest: null argument");
  HM_CUDA(cudaSetDevice(ctx->device));
  __half *dA = nullptr, *dB = nullptr;
  float* dO = nullptr;
  HM_CUDA(cudaMalloc(&dA, 64 * 64 * 2));
  HM_CUDA(cudaMalloc(&dB, 128 * 64 * 2));
  HM_CUDA(cudaMalloc(&dO, 128 * 256 * 4));
  HM_CUDA(cudaMemcpy(dA, h_A, 64 * 64 * 2, cudaMemcpyHostToDevice));
  HM_CUDA(cudaMemcpy(dB, h_B, 128 * 64 * 2, cudaMemcpyHostToDevice));
  HM_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  tc_selftest_kernel<<<1, 128, 32768>>>(dA, dB, dO, lane_off, col_off, repeats < 1 ? 1 : repeats);
  HM_CUDA(cudaGetLastError());
  HM_CUDA(cudaDeviceSynchronize());
  HM_CUDA(cudaMemcpy(h_out, dO, 128 * 256 * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return HM_OK;
}

// Debug export: cumulative wait-cycle counters of the producer / MMA threads (summed over CTAs and launches):
// out[0] producer waiting for a free slot, out[1..3] MMA thread waiting for the A operand / a free partial
// buffer / a full weight stage, out[4] MMA thread total cycles; out[5..8] first epilogue warp of the leader (or single)
// CTAs: waiting for a partial accumulator, adding it, finalizing, total; out[9..12] the same for the peer CTAs.
// Resets the counters.
extern "C" int hm_debug_tc_wait_cycles(hm_context* ctx, unsigned long long* out) {
  HM_CHECK(ctx && out && ctx->d_tc_flags, "hm_debug_tc_wait_cycles: bad argument");
  HM_CUDA(cudaDeviceSynchronize());
  HM_CUDA(cudaMemcpy(out, ctx->d_tc_flags + HM_TC_FLAG_DEBUG, sizeof(unsigned long long) * 13, cudaMemcpyDeviceToHost));
  HM_CUDA(cudaMemset(ctx->d_tc_flags + HM_TC_FLAG_DEBUG, 0, sizeof(unsigned long long) * 13));
  return HM_OK;
}

// Debug export: with `enable` != 0 allocates (and clears) the timeline buffer so that the following decoder launches record
// into it; with h_out != NULL copies 3 regions x 8192 (code, clock) pairs to the host.  enable == 0 frees the buffer.
extern "C" int hm_debug_tc_trace(hm_context* ctx, int enable, uint32_t* h_out) {
  HM_CHECK(ctx, "hm_debug_tc_trace: null context");
  HM_CUDA(cudaSetDevice(ctx->device));
  HM_CUDA(cudaDeviceSynchronize());
  const size_t bytes = sizeof(uint32_t) * 3 * 8192 * 2;
  if (h_out && ctx->d_tc_trace) HM_CUDA(cudaMemcpy(h_out, ctx->d_tc_trace, bytes, cudaMemcpyDeviceToHost));
  if (enable) {
    if (!ctx->d_tc_trace) HM_CUDA(cudaMalloc(&ctx->d_tc_trace, bytes));
    HM_CUDA(cudaMemset(ctx->d_tc_trace, 0, bytes));
  } else if (ctx->d_tc_trace) {
    cudaFree(ctx->d_tc_trace);
    ctx->d_tc_trace = nullptr;
  }
  return HM_OK;
}

extern "C" int hm_debug_tc_mma_rate(hm_context* ctx, int M, int N, int reps, int n_acc, long long* h_out) {
  HM_CHECK(ctx && h_out, "hm_debug_tc_mma_rate: bad argument");
  HM_CUDA(cudaSetDevice(ctx->device));
  long long* d = nullptr;
  HM_CUDA(cudaMalloc(&d, 16));
  HM_CUDA(cudaFuncSetAttribute(tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  tc_mma_rate_kernel<<<1, 128, 65536>>>(M, N, reps, n_acc, d);
  HM_CUDA(cudaGetLastError());
  HM_CUDA(cudaDeviceSynchronize());
  HM_CUDA(cudaMemcpy(h_out, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return HM_OK;
}

// Debug export: h_A [2][m_rows][64], h_B [N][64] fp16 bit patterns, h_out [2][128][256] fp32 TMEM dumps of the two CTAs,
// h_cycles [2] = (issue, issue + completion) cycles of reps x 4 MMAs.
extern "C" int hm_debug_tc_pair_probe(hm_context* ctx, const uint16_t* h_A, const uint16_t* h_B, float* h_out, long long* h_cycles, int m_rows,
                                      int N, int reps) {
  HM_CHECK(ctx && h_A && h_B && h_out && h_cycles && (m_rows == 64 || m_rows == 128) && N >= 32 && N <= 256 && N % 32 == 0, "hm_debug_tc_pair_probe: bad argument");
  HM_CUDA(cudaSetDevice(ctx->device));
  __half *dA = nullptr, *dB = nullptr;
  float* dO = nullptr;
  long long* dC = nullptr;
  HM_CUDA(cudaMalloc(&dA, 2 * m_rows * 64 * 2));
  HM_CUDA(cudaMalloc(&dB, N * 64 * 2));
  HM_CUDA(cudaMalloc(&dO, 2 * 128 * 256 * 4));
  HM_CUDA(cudaMalloc(&dC, 16));
  HM_CUDA(cudaMemcpy(dA, h_A, 2 * m_rows * 64 * 2, cudaMemcpyHostToDevice));
  HM_CUDA(cudaMemcpy(dB, h_B, N * 64 * 2, cudaMemcpyHostToDevice));
  HM_CUDA(cudaFuncSetAttribute(tc_pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 49152; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  HM_CUDA(cudaLaunchKernelEx(&cfg, tc_pair_probe_kernel, (const __half*)dA, (const __half*)dB, dO, dC, m_rows, N, reps < 1 ? 1 : reps));
  HM_CUDA(cudaGetLastError());
  HM_CUDA(cudaDeviceSynchronize());
  HM_CUDA(cudaMemcpy(h_out, dO, 2 * 128 * 256 * 4, cudaMemcpyDeviceToHost));
  HM_CUDA(cudaMemcpy(h_cycles, dC, 16, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
  return HM_OK;
}

extern "C" int hm_debug_tc_ingest(hm_context* ctx, int cluster, int n_stages, int bytes, double* h_bytes_per_clk_per_sm, double* h_ms) {
  HM_CHECK(ctx && ctx->d_tc_blob, "hm_debug_tc_ingest: bad argument");
  HM_CUDA(cudaSetDevice(ctx->device));
  const int grid = (ctx->sm_count / cluster) * cluster;
  long long* d = nullptr;
  HM_CUDA(cudaMalloc(&d, sizeof(long long) * grid));
  HM_CUDA(cudaFuncSetAttribute(tc_ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * bytes + 1024));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 3 * bytes + 1024; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    HM_CUDA(cudaLaunchKernelEx(&cfg, tc_ingest_kernel, (const uint8_t*)ctx->d_tc_blob, (int64_t)ctx->tc_blob_bytes, n_stages, (uint32_t)bytes, d));
    cudaEventRecord(e1);
    HM_CUDA(cudaDeviceSynchronize());
  }
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(grid);
  HM_CUDA(cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double mx = 0; for (long long v : h) mx = std::max(mx, (double)v);
  *h_bytes_per_clk_per_sm = (double)n_stages * bytes / mx;
  *h_ms = ms;
  cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
  return HM_OK;
}

