// PTX wrappers and shared-memory layout helpers of the tensor-core decoder (sm_100a: tcgen05 / TMEM / mbarrier / bulk copies /
// thread-block clusters).  Internal header, included by decoder_tc.cu and by the test-only probes (testing/tc_probes.cu).
// The wrappers follow the naming of the CUTLASS / CuTe sources they were checked against (cited per function); they are
// hand-written inline PTX.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#if defined(HM_TESTING) && !defined(HM_TC_LIGHT)
#define HM_TC_COUNTERS 1      // the testing build keeps per-role wait-cycle counters (hm_debug_tc_wait_cycles)
#endif                        // (-DHM_TC_LIGHT: a testing build whose only instrumentation is one timeline event per op)

namespace hm_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity);
// HM_TC_COUNTERS (build flag) keeps per-role wait-cycle counters for hm_debug_tc_wait_cycles; off in the product build
template <bool kCluster>
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, long long& acc) {
#ifdef HM_TC_COUNTERS
  long long t0 = clock64();
#endif
  if constexpr (kCluster) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity);
#ifdef HM_TC_COUNTERS
  acc += clock64() - t0;
#endif
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// weight stages are multicast to every CTA of the cluster: this CTA fetches 1/C of the stage and the copy
// lands at the same shared-memory offset in all C CTAs, signalling the same-offset mbarrier in each of them
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM allocation for a single CTA (CG = 1) or a CTA pair (CG = 2: the same warp of BOTH CTAs executes it and both
// receive the same column address, cute::TMEM::Allocator2Sm)
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16, issued by one thread.  CG = 2: issued by the leader CTA of a pair; each CTA
// supplies its own 128 rows of A and one half of B's N rows from the same shared-memory offsets, and receives its own
// 128 rows of D in its own TMEM.
template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrive when all MMAs issued so far by this thread have completed.  CG = 2: the arrive is multicast to the
// same-offset barrier of both CTAs of the pair (cutlass::arch::umma_arrive_multicast_2x1SM).
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
// true in exactly one lane of a fully active warp (cute::elect_one_sync).  The warp-specialised roles run their loops with
// all 32 lanes (warp-uniform control flow keeps descriptors and addresses in uniform registers) and issue the asynchronous
// operations -- bulk copies, MMAs, commits -- from the elected lane only.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}
// address of `addr` (a shared::cta address of this CTA) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier of another CTA of the cluster (address from mapa_rank).  Default semantics, as
// cutlass::arch::ClusterBarrier::arrive(cta_id): a cluster-scope release would cost MEMBAR.ALL.GPU per arrive (measured:
// it serialised the weight ring), and nothing this thread wrote is read through the generic proxy on the other side --
// the A operand is published with fence.proxy.async and read by each CTA's own tensor core.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local mbarrier whose arrivals may come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "DONE_C:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
// ---- flag exchange between the two CTAs of a pair (zero-operand shortcut of the decoder kernel)
// plain store into the shared memory of another CTA of the cluster (address from mapa_rank)
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// arrive on another CTA's mbarrier with cluster-scope release: orders this thread's earlier (remote) stores before the
// waiters that observe the phase with cluster-scope acquire.  Used once per tile, never in the weight ring (it is a MEMBAR).
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local mbarrier with cluster-scope acquire (the data it guards was written by the peer CTA)
__device__ __forceinline__ void mbar_wait_acq_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_A:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_A;\n\t"
      "bra WAIT_LOOP_A;\n\t"
      "DONE_A:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzle UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start address
// bits [0,14) (>>4), LBO [16,30) = 1 (unused for swizzled K-major), SBO [32,46) = 1024 B (8 rows x 128 B),
// version [46,48) = 1 (Blackwell), layout_type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format F16 (0),
// K-major A and B, n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row, k) of a K-major SW128 tile whose rows are 64 fp16 (128 B) wide
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// packed fp32x2 arithmetic (FADD2 / FFMA2 on sm_100) and saturating fp16x2 pack
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
      "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&r);
}
// fp16x2 {lo half = a, hi half = b}, round-to-nearest, saturated to +-65504
__device__ __forceinline__ uint32_t pack_h2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// split a scaled fp32 pair into fp16 hi / lo words; `sat` keeps the running per-half maximum of |hi| (0x7bff = the
// conversion saturated: the kernel reports HM_STATUS_F16_SATURATED)
__device__ __forceinline__ void split2(float2 v, uint32_t& hi, uint32_t& lo, uint32_t& sat) {
  hi = pack_h2_sat(v.x, v.y);
  sat = __vmaxu2(sat, hi & 0x7fff7fffu);
  const float2 f = __half22float2(*reinterpret_cast<__half2*>(&hi));
  const float2 r = fma2(f, make_float2(-1.f, -1.f), v);
  lo = pack_h2_sat(r.x, r.y);
}

// A operand of one k-chunk (16 KB): the fp16 hi parts of the tile's 64 points as one K-major SW128 tile of 64 rows x 64 k
// (8 KB), followed by the lo parts as a second tile.  A pair MMA of M = 128 takes 64 rows from each CTA, so the hi tile
// and the lo tile are separate M operands; their products accumulate into the SAME TMEM rows.
constexpr int kAChunkBytes = 16384;               // one k-chunk of the A operand: 64 points x 64 k x 2 B, hi tile then lo tile
constexpr int kALoOffset = 8192;

// split two scaled fp32 values into packed fp16 hi and lo words (hi = rn(x) saturated to the finite fp16 range,
// lo = rn(x - hi)) and store them at columns (k, k+1) of point p in chunk `chunk` of the A operand
__device__ __forceinline__ void store_pair(uint8_t* smem, int chunk, int p, int k, float a, float b, int& sat) {
  const float ac = fminf(fmaxf(a, -65504.f), 65504.f), bc = fminf(fmaxf(b, -65504.f), 65504.f);
  sat |= (ac != a) | (bc != b);
  const __half2 hh = __floats2half2_rn(ac, bc);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(ac - hf.x, bc - hf.y);
  uint8_t* base = smem + (uint32_t)chunk * kAChunkBytes;
  *reinterpret_cast<__half2*>(base + sw128_offset(p, k)) = hh;
  *reinterpret_cast<__half2*>(base + kALoOffset + sw128_offset(p, k)) = ll;
}

// 16-column variants (the epilogue promotes a 32-column partial in two halves: 16 fewer live registers next to the 64 accumulators)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep16(float (&a)[16]) {
  uint32_t* x = reinterpret_cast<uint32_t*>(a);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8]), "+r"(x[9]),
                 "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15])
               :
               : "memory");
}

}  // namespace hm_tc
