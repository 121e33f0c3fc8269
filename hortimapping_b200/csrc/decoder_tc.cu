// Tensor-core decoder engine (HM_ENGINE_TC): the whole DeepSDF MLP -- forward and, optionally, the input gradient -- for
// tiles of 64 query rows per CTA in ONE persistent, warp-specialised sm_100a kernel launched as clusters of two CTAs.
//
// Restates deepsdf/networks/deep_sdf_decoder.py:75-110 (forward) and the autograd input gradient of
// wild_completion/utils.py:112-122,175-193.  DESIGN.md section 4.1 has the full description and the measurements; in short:
//
//   * every layer is a tcgen05.mma GEMM with fp32 accumulators in TMEM.  fp32 parity needs more than one fp16/bf16 MMA
//     (SURVEY.md 7.3): operands are split x*s = hi + lo (fp16 each, s a calibrated power of two); the A tile stacks the hi
//     and lo rows of the same 64 points as 128 MMA rows and every weight tile exists as a lo and a hi copy, so that
//     (hi + lo) x (hi + lo) runs at the full-rate shape M = 128 per CTA x N = 256 -- measured 3e-8 abs SDF error, fp32 grade.
//   * the pair issues cta_group::2 MMAs (M = 256): the B operand is split between the two CTAs' shared memories, each CTA
//     streams half of every weight tile (cp.async.bulk + mbarrier complete_tx, 6 x 16 KB ring; stages are stored in the blob
//     pre-swizzled and in consumption order).
//   * activations never leave the SM: the epilogue warps read partial accumulators from TMEM, apply bias/ReLU (or the ReLU
//     mask in the backward pass), re-split to fp16 hi/lo and write the next layer's A operand straight into shared memory
//     in the 128-byte-swizzled K-major UMMA layout, in place.
//   * the tensor core accumulates fp32 with round-toward-zero (measured: -9e-6 relative after the 96 chained MMAs of one
//     layer), so MMAs are chained only inside an accumulation group (2 k-chunks x one 256-column output half, 16 MMAs) into
//     one of two TMEM buffers, and the group partials are summed in registers in fp32 round-to-nearest.
//
// Warp roles (640 threads): warpgroup 0 = control (warp 0 bulk-copy producer, warp 1 MMA issuer in the leader CTA / weight
// arrival forwarder in the peer CTA + TMEM alloc, warps 2-3 idle), warpgroups 1-4 = 16 epilogue warps (4 per TMEM
// sub-partition, 64 output columns of each half per warp); setmaxnreg moves registers from the control warpgroup to them.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kWeightRing = 98304;                 // shared memory of the weight ring: 3 x 32 KB (single CTA) or 6 x 16 KB (CTA pair)
constexpr int kMaxStages = 6;
constexpr int kAChunkBytes = 16384;                // 128 stacked rows (64 points x {hi, lo}) x 64 k x 2 B
constexpr int kSmemA = 0;
constexpr int kSmemStages = 131072;
constexpr int kSmemBars = kSmemStages + kWeightRing;             // 229376
constexpr int kSmemTotal = kSmemBars + 256 + 2048;
constexpr int kBufs = 4;                           // TMEM accumulator buffers of 128 columns
constexpr int kEpiWarps = 16;
constexpr int kCtrlWarps = 4;                      // one full warpgroup: producer, MMA issuer, two idle warps (setmaxnreg is per warpgroup)
constexpr int kThreads = (kCtrlWarps + kEpiWarps) * 32;
// setmaxnreg moves registers inside the CTA's OWN allocation (threads x launch registers = 640 x 96): the control warpgroup
// gives up 128 x (96 - 24) = 9216 registers and the 512 epilogue threads can take at most that many: 96 + 16 = 112.  (A request
// the pool cannot satisfy does not fail, it spins forever -- measured.)
constexpr int kCtrlRegs = 24, kEpiRegs = 112;
constexpr int kMaskWordsPerOp = 64 * 16;           // 64 points x 512 bits per forward layer

// barrier slots (8 bytes each) inside the barrier block
enum { BAR_W_FULL = 0, BAR_W_EMPTY = kMaxStages, BAR_A_READY = 2 * kMaxStages, BAR_PART_FULL = BAR_A_READY + 4,
       BAR_PART_EMPTY = BAR_PART_FULL + 4, BAR_COUNT = BAR_PART_EMPTY + 4 };
static_assert(8 * BAR_COUNT + 4 <= 256, "barrier block overflows into the dot-product scratch");

struct TcParams {
  hm_tc_plan plan;
  const uint8_t* blob;
  int64_t blob_bytes;         // size of one copy of the weight blob
  int32_t blob_copies;        // copies laid out back to back (CTAs spread over them)
  const float* bias;          // [8][512]
  const float* w8;            // [512]
  const float* b8;            // [1]
  const float* rows;          // [n][35] or null
  const float* xyz;           // [n][3]
  const float* latents;       // [L][32]
  const int32_t* row_latent;  // [n] or null
  const int32_t* n_dynamic;   // device row count or null
  int64_t n;
  float* sdf;
  float* jac;
  uint32_t* masks;            // [grid][8][64][16]
  int32_t* flags;             // [0] = saturation count
  uint32_t* trace;            // debug timeline of CTA 0 / 1 (hm_debug_tc_trace) or null
  int32_t grid_n;             // > 0: xyz of row i = voxel grid point i (fused mesher grid, hm_rows)
  float grid_voxel, grid_radius;
};

// ------------------------------------------------------------------ PTX wrappers
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

// wait for all outstanding TMEM loads; the 32 loaded registers pass through the statement so that no consumer can be
// scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait_dep32(float (&a)[32]) {
  uint32_t* x = reinterpret_cast<uint32_t*>(a);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8]), "+r"(x[9]),
                 "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15])
               :
               : "memory");
  asm volatile(""
               : "+r"(x[16]), "+r"(x[17]), "+r"(x[18]), "+r"(x[19]), "+r"(x[20]), "+r"(x[21]), "+r"(x[22]), "+r"(x[23]), "+r"(x[24]),
                 "+r"(x[25]), "+r"(x[26]), "+r"(x[27]), "+r"(x[28]), "+r"(x[29]), "+r"(x[30]), "+r"(x[31])
               :
               : "memory");
}
// TMEM -> registers, 32 lanes x 32 consecutive columns (thread = lane = accumulator row), without the wait
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
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
}

// k-chunk order of an 8-chunk op: k-step s multiplies chunks {0,2}, {1,3}, {4,6}, {5,7}.  Steps 0,1 read the chunks that the
// previous op's output half 0 becomes (0..3), steps 2,3 those of its half 1 (4..7).
__host__ __device__ __forceinline__ int chunk_of(int step, int which) { return (step & 1) + 4 * (step >> 1) + 2 * which; }

// Accumulation groups of one op, in issue order.  A group = (k-step, 256-column output half nh).  For an 8-chunk op with
// two output halves the order is (0,0) (0,1) (1,0) (1,1) (2,0) (3,0) (2,1) (3,1): output half 0 is complete TWO groups
// before the op ends, so the epilogue warps turn it into the next op's A chunks 0..3 while the tensor core is still busy
// with half 1, and the next op (whose first four groups only read chunks 0..3) starts without a bubble.
__host__ __device__ __forceinline__ int groups_of(int n_kchunks, int n_nblocks) { return n_kchunks == 1 ? n_nblocks : 4 * n_nblocks; }
__host__ __device__ __forceinline__ void group_of(int n_kchunks, int n_nblocks, int g, int& step, int& nh) {
  if (n_kchunks == 1) { step = 0; nh = g; }
  else if (n_nblocks == 1) { step = g; nh = 0; }
  else { step = (0x32321100u >> (4 * g)) & 0xF; nh = (0xCAu >> g) & 1; }
}

// The kernel runs as clusters of two CTAs (one TPC).  Each CTA owns a tile of 64 points; the pair issues cta_group::2 MMAs
// of M = 128 (64 rows per CTA -- measured: full rate, 64 cycles for N = 256, scratch/dbg_pair_probe.py) whose B operand is
// split between the two CTAs' shared memories, so each CTA streams only HALF of every weight tile from L2.  Per weight
// tile pair the THREE needed products are issued -- A_hi x W_lo, A_lo x W_hi, A_hi x W_hi -- into the same accumulator
// rows (the 64 x N accumulator of each CTA is folded onto 128 lanes x N/2 columns: lanes 0..63 hold output columns
// [0, N/2), lanes 64..127 columns [N/2, N)).  CTA rank 0 (the leader) issues all MMAs.
template <bool kJac>
__global__ void __launch_bounds__(kThreads, 1) tc_decoder_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int CG = 2;
  constexpr int kStages = kMaxStages;
  constexpr int kStageBytes = kWeightRing / kStages;      // 128 output features x 64 k x 2 B (hi OR lo) = this CTA's half of a tile
  constexpr bool kPair = true;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;    // warp index as a uniform value
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + kSmemBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kSmemBars + 8 * BAR_COUNT);
  float* dot_scratch = reinterpret_cast<float*>(smem + kSmemBars + 256);      // [64 points][8 column groups] (lin8 tail) ...
  float* bias_s = dot_scratch;                                                // ... and, before it, the current forward op's 512 biases
  constexpr int kOps = kJac ? HM_TC_NOPS_ALL : HM_TC_NOPS_FWD;
  const uint32_t rank = cluster_ctarank();                                    // 0 = leader (MMA issuer) of the pair
  const uint32_t lead_bars = mapa_rank(bars, 0);                              // the leader's barrier block (cluster address)
  auto lead_bar = [&](int i) { return lead_bars + 8u * i; };
  const int64_t unit0 = blockIdx.x >> 1, unit_stride = gridDim.x >> 1;
  // debug timeline: (code << 24 | op << 16 | index, clock) pairs of the first CTA pair; region 0 = MMA issuer,
  // 1 / 2 = first epilogue warp of the leader / peer CTA
  constexpr uint32_t kTraceCap = 8192;
  const bool tracing = P.trace != nullptr && blockIdx.x < 2;
  uint32_t trace_n = 0;
  auto trace = [&](int region, uint32_t code, uint32_t op, uint32_t idx) {
    if (tracing && trace_n < kTraceCap) {
      uint32_t* t = P.trace + (size_t)region * kTraceCap * 2 + 2 * trace_n;
      t[0] = (code << 24) | (op << 16) | idx;
      t[1] = (uint32_t)clock64();
      ++trace_n;
    }
  };

  if (threadIdx.x == 0) {
    // W_FULL of the leader also collects the peer's "my half has landed" arrive
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(BAR_W_FULL + s), rank == 0 ? 2 : 1); mbar_init(bar(BAR_W_EMPTY + s), 1); }
    for (int p = 0; p < 4; ++p) mbar_init(bar(BAR_A_READY + p), kEpiWarps * CG / 2);      // a k-step's chunks come from half of the warps
    for (int b = 0; b < kBufs; ++b) { mbar_init(bar(BAR_PART_FULL + b), 1); mbar_init(bar(BAR_PART_EMPTY + b), kEpiWarps * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<CG>(smem_u32((const void*)tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (lane == 0 && (warp == 1 || warp == kCtrlWarps)) trace(warp == 1 ? 0 : 1 + rank, 0, 0, 0);     // common time origin

  const int64_t n_rows = P.n_dynamic ? (int64_t)min((int64_t)*P.n_dynamic, P.n) : P.n;
  const int64_t n_tiles = (n_rows + HM_TC_TILE_M - 1) / HM_TC_TILE_M;
  const int64_t n_units = (n_tiles + 1) / 2;               // a unit = one tile per CTA of the pair

  if (warp < kCtrlWarps) {
  // the control warpgroup hands most of its registers to the four epilogue warpgroups (64 accumulators per thread)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtrlRegs));
  if (warp == 0) {
    // ===================== weight producer (every CTA fetches its half of each stage) =====================
    uint32_t slot = 0, phase = 0;
    long long t_empty = 0;
    for (int64_t unit = unit0; unit < n_units; unit += unit_stride) {
      for (int op = 0; op < kOps; ++op) {
        const hm_tc_op& o = P.plan.ops[op];
        const uint32_t bytes = (uint32_t)o.stage_rows * 128u / CG;     // this CTA's rows of one 64-k fp16 tile
        const int nst = o.n_kchunks * o.n_nblocks * 2;                 // (chunk, n-half) x {lo, hi}
        const uint8_t* src = P.blob + o.blob_offset + (size_t)rank * bytes;
        for (int s = 0; s < nst; ++s) {
          mbar_wait_timed<false>(bar(BAR_W_EMPTY + slot), phase ^ 1, t_empty);
          if (elect_one()) {
            mbar_expect_tx(bar(BAR_W_FULL + slot), bytes);
            bulk_g2s(smem_base + kSmemStages + slot * kStageBytes, src + (size_t)s * bytes * CG, bytes, bar(BAR_W_FULL + slot));
          }
          __syncwarp();
          if (++slot == kStages) { slot = 0; phase ^= 1; }
        }
      }
    }
    if (P.flags && rank == 0 && lane == 0) atomicAdd((unsigned long long*)(P.flags + 8), (unsigned long long)t_empty);
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA; whole warp walks the loops, one elected lane issues) =====================
      // One accumulation group = (k-step = 2 k-chunks, 256-column output half): per chunk A_hi x W_lo (4 MMAs), then A_lo x W_hi
      // and A_hi x W_hi (8 MMAs), each M = 64 per CTA x N = 256 x K = 16, into a FRESH 128-column TMEM buffer (four buffers).
      // The tensor core accumulates fp32 with round-toward-zero (measured: -1e-7 relative per chained MMA), so chains are
      // kept to one group and the epilogue warps add the group partials in fp32 round-to-nearest.
      uint32_t slot = 0, phase = 0, op_seq = 0, gseq = 0;
      long long t_a = 0, t_part = 0, t_w = 0;
#ifdef HM_TC_COUNTERS
      const long long t_begin = clock64();
#endif
      for (int64_t unit = unit0; unit < n_units; unit += unit_stride) {
        for (int op = 0; op < kOps; ++op, ++op_seq) {
          const hm_tc_op& o = P.plan.ops[op];
          const uint32_t idesc = make_idesc(64 * CG, o.stage_rows);
          const int ng = groups_of(o.n_kchunks, o.n_nblocks);
          const int nwhich = (o.n_kchunks == 1) ? 1 : 2;
          int steps_ready = 0;
          if (o.n_kchunks == 1) {                  // F0 reads chunk 0 only; its four A_READY phases are consumed up front
            for (; steps_ready < 4; ++steps_ready) mbar_wait_timed<kPair>(bar(BAR_A_READY + steps_ready), op_seq & 1, t_a);
            tc_fence_after();
          }
          for (int g = 0; g < ng; ++g, ++gseq) {
            int step, nh;
            group_of(o.n_kchunks, o.n_nblocks, g, step, nh);
            if (steps_ready <= step) {
              for (; steps_ready <= step; ++steps_ready) mbar_wait_timed<kPair>(bar(BAR_A_READY + steps_ready), op_seq & 1, t_a);
              tc_fence_after();
            }
            const uint32_t buf = gseq & (kBufs - 1);
            if (lane == 0) trace(0, 1, op, g);
            mbar_wait_timed<kPair>(bar(BAR_PART_EMPTY + buf), ((gseq / kBufs) & 1) ^ 1, t_part);
            tc_fence_after();
            if (lane == 0) trace(0, 2, op, g);
            const uint32_t d = tmem_base + buf * 128;
#pragma unroll
            for (int part = 0; part < 2; ++part) {                   // weight tiles: 0 = lo (small terms first), 1 = hi
              for (int which = 0; which < nwhich; ++which) {
                const int chunk = (o.n_kchunks == 1) ? 0 : chunk_of(step, which);
                const uint64_t a_hi = make_desc(smem_base + kSmemA + chunk * kAChunkBytes);
                const uint64_t a_lo = a_hi + (kALoOffset >> 4);
                const uint64_t w_desc = make_desc(smem_base + kSmemStages + slot * kStageBytes);
                mbar_wait_timed<kPair>(bar(BAR_W_FULL + slot), phase, t_w);
                tc_fence_after();
                if (elect_one()) {
                  if (part == 0) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)                   // +32 B per 16-wide k step = +2 in the descriptor's address field
                      umma_f16<CG>(d, a_hi + 2 * ks, w_desc + 2 * ks, idesc, (which | ks) ? 1u : 0u);
                  } else {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_f16<CG>(d, a_lo + 2 * ks, w_desc + 2 * ks, idesc, 1u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_f16<CG>(d, a_hi + 2 * ks, w_desc + 2 * ks, idesc, 1u);
                  }
                  umma_commit<CG>(bar(BAR_W_EMPTY + slot));          // frees the slot in both CTAs of the pair
                  if (part == 1 && which == nwhich - 1) umma_commit<CG>(bar(BAR_PART_FULL + buf));
                }
                __syncwarp();
                if (++slot == kStages) { slot = 0; phase ^= 1; }
              }
            }
            if (lane == 0) trace(0, 3, op, g);
          }
        }
      }
#ifdef HM_TC_COUNTERS
      if (P.flags && lane == 0) {
        atomicAdd((unsigned long long*)(P.flags + 10), (unsigned long long)t_a);
        atomicAdd((unsigned long long*)(P.flags + 12), (unsigned long long)t_part);
        atomicAdd((unsigned long long*)(P.flags + 14), (unsigned long long)t_w);
        atomicAdd((unsigned long long*)(P.flags + 16), (unsigned long long)(clock64() - t_begin));
      }
#endif
    } else {
      // ===================== peer CTA: tell the leader that this CTA's half of a weight stage has landed =====================
      uint32_t slot = 0, phase = 0;
      for (int64_t unit = unit0; unit < n_units; unit += unit_stride) {
        for (int op = 0; op < kOps; ++op) {
          const hm_tc_op& o = P.plan.ops[op];
          const int nst = o.n_kchunks * o.n_nblocks * 2;
          for (int s = 0; s < nst; ++s) {
            mbar_wait(bar(BAR_W_FULL + slot), phase);
            if (lane == 0) mbar_arrive_cluster(lead_bar(BAR_W_FULL + slot));
            __syncwarp();
            if (++slot == kStages) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    // ===================== epilogue warps (16) =====================
    // TMEM sub-partition sp = warp % 4 (lanes 32*sp ..): lanes 0..63 hold the tile's 64 points for output columns [0, 128) of
    // a 256-column half, lanes 64..127 the same points for columns [128, 256).  So a thread is ONE point (p = 32*(sp & 1) +
    // lane), hq = sp / 2 selects the 128-column quarter and the four warps of a sub-partition split its 128 TMEM columns
    // (cq = 0..3, 32 columns each): a thread owns 32 CONSECUTIVE output columns of each half = half a k-chunk row of the next
    // op's A operand, which it writes with four 16-byte stores for the hi and four for the lo part.
    const int e_w = warp - kCtrlWarps;
    const int sp = warp & 3;
    const int hq = sp >> 1;
    const int cq = e_w >> 2;
    const int g8 = 4 * hq + cq;              // this thread's column group among the 8 of a half
    const int p = 32 * (sp & 1) + lane;
    const uint32_t t_addr = tmem_base + ((uint32_t)(32 * sp) << 16) + 32 * cq;
    uint32_t* my_masks = P.masks + ((size_t)blockIdx.x * 8 * kEpiWarps * 32 + (size_t)(e_w * 32 + lane)) * 2;   // [op][thread][2 words]
    constexpr size_t kMaskStride = (size_t)kEpiWarps * 32 * 2;
    uint32_t op_seq = 0, gseq = 0;
    int sat = 0;
#ifdef HM_TC_COUNTERS
    long long t_pfull = 0, t_pbody = 0, t_fin = 0;       // debug counters (hm_debug_tc_wait_cycles)
    const long long t_epi_begin = clock64();
#endif
    auto publish = [&](int j) {              // this warp's part of k-step j of the next A operand is written
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_bar(BAR_A_READY + j));
    };
    // First of this thread's 32 consecutive columns of output half nh (within the 512-wide layer), and where they go in the
    // next A operand: chunk 4*nh + 2*hq + cq/2, k = 32*(cq & 1) + j, i.e. 16-byte units 4*(cq & 1) .. +3 of row p.
    auto col0_of = [&](int nh) { return 256 * nh + 128 * hq + 32 * cq; };
    uint8_t* const st_row = smem + kSmemA + (uint32_t)(2 * hq + (cq >> 1)) * kAChunkBytes + (uint32_t)(p >> 3) * 1024 + (p & 7) * 128;
    const uint32_t u_base = 4 * (cq & 1), r7 = (uint32_t)p & 7u;
    uint32_t sat2 = 0;
    // split 8 consecutive values (one 16-byte unit u = 0..3 of this thread's 32 columns) and store the hi and lo units
    auto emit_unit = [&](int nh, int u, const float (&y)[8]) {
      uint4 hi, lo;
      split2(make_float2(y[0], y[1]), hi.x, lo.x, sat2);
      split2(make_float2(y[2], y[3]), hi.y, lo.y, sat2);
      split2(make_float2(y[4], y[5]), hi.z, lo.z, sat2);
      split2(make_float2(y[6], y[7]), hi.w, lo.w, sat2);
      uint8_t* dst = st_row + (uint32_t)(4 * nh) * kAChunkBytes + (((u_base + (uint32_t)u) ^ r7) << 4);
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + kALoOffset) = lo;
    };
    for (int64_t unit = unit0; unit < n_units; unit += unit_stride) {
      const int64_t tile = 2 * unit + rank;       // the odd CTA of the last pair may get an all-padding tile
      const int64_t grow = tile * HM_TC_TILE_M + p;
      const bool ok = grow < n_rows;
      // raw input x0 = [latent(32), xyz(3)] of this thread's point (deep_sdf_decoder.py:76-88)
      const int64_t lr = ok ? grow : (n_rows - 1);
      const int32_t li = (!P.rows && P.row_latent) ? __ldg(P.row_latent + lr) : 0;      // latent-table row of the point
      const float* lat_ptr = P.rows ? P.rows + lr * HM_IN : P.latents + (size_t)li * HM_LATENT;   // the 32 latent values are contiguous in both input modes
      auto x0_xyz = [&](int c) -> float {      // xyz coordinate c of the point
        if (P.rows) return __ldg(lat_ptr + HM_LATENT + c);
        if (P.grid_n > 0) return hm_grid_coord(lr, c, P.grid_n, P.grid_voxel, P.grid_radius);
        return __ldg(P.xyz + lr * 3 + c);
      };
      // ---- A operand of F0: chunk 0 = [x0 * s, 0 ...] (K padded 35 -> 64); column group g8 writes k in [8*g8, +8)
      {
        const float s0 = P.plan.ops[0].in_scale;
        float xin[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) xin[i] = 0.f;
        if (g8 < 4) {                            // 8 independent loads: one memory latency, not eight
#pragma unroll
          for (int i = 0; i < 8; ++i) xin[i] = __ldg(lat_ptr + 8 * g8 + i);
        } else if (g8 == 4) {
#pragma unroll
          for (int c = 0; c < 3; ++c) xin[c] = x0_xyz(c);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) store_pair(smem, 0, p, 8 * g8 + 2 * e, xin[2 * e] * s0, xin[2 * e + 1] * s0, sat);
        publish(cq >> 1);                        // every warp arrives once per k-step of its own chunks' parity (see finalize)
        publish(2 + (cq >> 1));
      }
      float f_out = 0.f;
      // Accumulators of this thread: acc[nh][i] = columns col0_of(nh) + 2i + {0, 1}.
      float2 acc[2][16];
      uint32_t m0 = 0u, m1 = 0u;             // ReLU bits of the current op: output half 0 / half 1, bit j = column col0 + j
      float dot = 0.f;
      const std::integral_constant<int, 0> I0{};
      const std::integral_constant<int, 1> I1{};
#pragma unroll 1
      for (int op = 0; op < kOps; ++op, ++op_seq) {
        const hm_tc_op& o = P.plan.ops[op];
        const float unscale = o.out_unscale;
        const float s_next = (op + 1 < kOps) ? P.plan.ops[op + 1].in_scale : 1.f;
        const float k_mul = unscale * s_next;
        const bool narrow = (o.stage_rows == 64);        // B0: 64 output columns (32 TMEM columns), one group per step
        const bool wide = (o.n_kchunks != 1);
        m0 = m1 = 0u;
        if (kJac && op >= 8 && op < 15) {
          const uint2 mw = *reinterpret_cast<const uint2*>(my_masks + (size_t)(14 - op) * kMaskStride);   // ReLU mask of h_{l-1}, l = 15 - op
          m0 = mw.x; m1 = mw.y;
        }
        // collect the partial accumulator of one (step, n-half) group.  FIRST: the group opens the op for this output half
        // (overwrite instead of accumulate).
        auto promote = [&](auto NH, auto FIRST) {
          constexpr int nh = decltype(NH)::value;
          constexpr bool first = decltype(FIRST)::value != 0;
          const uint32_t buf = gseq & (kBufs - 1);
#ifdef HM_TC_COUNTERS
          const long long tp0 = clock64();
#endif
          mbar_wait(bar(BAR_PART_FULL + buf), (gseq / kBufs) & 1);
#ifdef HM_TC_COUNTERS
          const long long tp1 = clock64();
          t_pfull += tp1 - tp0;
#endif
          tc_fence_after();
          if (e_w == 0 && lane == 0) trace(1 + rank, 10, op, gseq & 0xffff);
          auto release = [&]() {               // all TMEM reads of this buffer are complete
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_bar(BAR_PART_EMPTY + buf));
          };
          if (narrow && cq != 0) {             // B0's 64 output columns occupy TMEM columns 0..31 only
            release();
          } else {
            float v[32];
            tmem_ld32_nowait(t_addr + buf * 128, v);
            tmem_ld_wait_dep32(v);
            release();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 w = make_float2(v[2 * i], v[2 * i + 1]);
              if constexpr (first) acc[nh][i] = w; else acc[nh][i] = add2(acc[nh][i], w);
            }
          }
          ++gseq;
#ifdef HM_TC_COUNTERS
          t_pbody += clock64() - tp1;
#endif
          if (e_w == 0 && lane == 0) trace(1 + rank, 12, op, gseq & 0xffff);
        };
        // Turn the finished output half nh of op `opx` into the next op's A chunks 4*nh .. 4*nh+3 (k-steps 2*nh, 2*nh+1) or the
        // final outputs.  kClass: 0 = any op (1 / 2 restrict the compiled branches to hidden layers / lin7).
        auto finalize = [&](auto NH, auto CLASS, const int opx, const float k_mul_x, const float unscale_x, const float s_next_x, uint32_t& m_) {
          constexpr int nh = decltype(NH)::value, kClass = decltype(CLASS)::value;
#ifdef HM_TC_COUNTERS
          const long long tf0 = clock64();
#endif
          if (e_w == 0 && lane == 0) trace(1 + rank, 13, opx, nh);
          const int col0 = col0_of(nh);
          if (kClass != 2 && opx < 7) {
            // ---------------- forward hidden layer: h = relu(acc + b); next A = h * s_next (bias pre-scaled by s_next)
            const float* bias = bias_s + col0;
            const float2 kk = make_float2(k_mul_x, k_mul_x);
            if (nh == 0 && cq >= 2) asm volatile("bar.sync 2, %0;" ::"n"(kEpiWarps * 32) : "memory");      // k-step 0 first (see below)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 b0 = *reinterpret_cast<const float4*>(bias + 8 * u), b1 = *reinterpret_cast<const float4*>(bias + 8 * u + 4);
              float2 y0 = fma2(acc[nh][4 * u + 0], kk, make_float2(b0.x, b0.y)), y1 = fma2(acc[nh][4 * u + 1], kk, make_float2(b0.z, b0.w));
              float2 y2 = fma2(acc[nh][4 * u + 2], kk, make_float2(b1.x, b1.y)), y3 = fma2(acc[nh][4 * u + 3], kk, make_float2(b1.z, b1.w));
              const float y[8] = {y0.x, y0.y, y1.x, y1.y, y2.x, y2.y, y3.x, y3.y};
              float r[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                m_ |= ((y[i] > 0.f) ? 1u : 0u) << (8 * u + i);
                r[i] = fmaxf(y[i], 0.f);
              }
              emit_unit(nh, u, r);
            }
            if (nh == 1 && opx == 3 && hq == 1 && cq >= 2) {
              // lin3 has 477 outputs; columns 477..511 of the next input are the raw x0 (skip concat, deep_sdf_decoder.py:87-88).
              // Kept out of the loop above (a branch per element would end its instruction-level parallelism): the threads that
              // own columns >= 477 rewrite those units and clear their ReLU bits.
              if (cq == 3) {                           // columns 480..511 = x0[3..34]: 29 latent values + xyz, loaded as one batch
                float xv[32];
#pragma unroll
                for (int j = 0; j < 29; ++j) xv[j] = __ldg(lat_ptr + 3 + j);
#pragma unroll
                for (int c = 0; c < 3; ++c) xv[29 + c] = x0_xyz(c);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float r[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) r[i] = xv[8 * u + i] * s_next_x;
                  emit_unit(nh, u, r);
                }
                m_ = 0u;
              } else {                                 // columns 448..479: 472..476 are the last lin3 outputs, 477..479 = x0[0..2]
                float r[8];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                  const int j = 24 + i;
                  const float a = (j & 1) ? acc[nh][j >> 1].y : acc[nh][j >> 1].x;
                  r[i] = fmaxf(fmaf(a, k_mul_x, bias[j]), 0.f);
                }
#pragma unroll
                for (int i = 5; i < 8; ++i) r[i] = __ldg(lat_ptr + (i - 5)) * s_next_x;
                emit_unit(nh, 3, r);
                m_ &= 0x1fffffffu;
              }
            }
            // Warps cq = 0,1 own the chunks of k-step 2*nh, warps cq = 2,3 those of k-step 2*nh + 1.  For output half 0 the
            // second pair waits for the first: the next op's first group needs k-step 0 only, and two warps per scheduler
            // deliver it in half the time four would need for both steps.
            publish(2 * nh + (cq >> 1));
            if (nh == 0 && cq < 2) asm volatile("bar.arrive 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
          } else if (kClass != 1 && opx == 7) {
            // ---------------- lin7 epilogue + the lin8 dot product (deep_sdf_decoder.py:107-108)
            const float* bias = bias_s + col0;
            const float* w8 = P.w8 + col0;
            const float2 uu = make_float2(unscale_x, unscale_x);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 bz = *reinterpret_cast<const float2*>(bias + 2 * i);
              const float2 wz = __ldg(reinterpret_cast<const float2*>(w8 + 2 * i));
              const float2 y = fma2(acc[nh][i], uu, bz);
              m_ |= ((y.x > 0.f) ? 1u : 0u) << (2 * i) | ((y.y > 0.f) ? 1u : 0u) << (2 * i + 1);
              dot = fmaf(fmaxf(y.x, 0.f), wz.x, dot);
              dot = fmaf(fmaxf(y.y, 0.f), wz.y, dot);
            }
          } else if (kClass != 2 && opx > 7 && opx < 15) {
            // ---------------- backward through lin_l (l = 15 - opx = 7..1): d_{l-1} = (d_l W_l) * relu'(h_{l-1})
            if (nh == 0 && cq >= 2) asm volatile("bar.sync 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float r[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int j = 8 * u + i;
                const float a = (j & 1) ? acc[nh][j >> 1].y : acc[nh][j >> 1].x;
                r[i] = a * (((m_ >> j) & 1u) ? k_mul_x : 0.f);
              }
              emit_unit(nh, u, r);
            }
            if (nh == 1 && opx == 11 && hq == 1 && cq >= 2) {
              // columns 477..511 of d(lin4 input) are the gradient w.r.t. the concatenated raw input x0 (deep_sdf_decoder.py:87-88).
              // They are parked in the output Jacobian row; B0 adds the rest.  (Their ReLU bits are clear, so the loop above wrote
              // zeros for them into the next A operand.)
              if (ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float a = (j & 1) ? acc[nh][j >> 1].y : acc[nh][j >> 1].x;
                  if (col0 + j >= HM_SKIP_COL) __stcg(P.jac + grow * HM_IN + (col0 + j - HM_SKIP_COL), a * unscale_x);
                }
              }
              __threadfence_block();
            }
            publish(2 * nh + (cq >> 1));
            if (nh == 0 && cq < 2) asm volatile("bar.arrive 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
          }
#ifdef HM_TC_COUNTERS
          t_fin += clock64() - tf0;
#endif
          if (e_w == 0 && lane == 0) trace(1 + rank, 14, opx, nh);
        };
        // ---- Schedule of one op (groups in the issue order of group_of(); P = promote, F(nh) = finalize an output half):
        //        P(0,0) P(0,1) P(1,0) P(1,1) P(2,0) P(3,0) F(0) P(2,1) P(3,1) F(1)
        //      Output half 0 is complete two groups before the op ends and is turned into the next op's A chunks 0..3 while the
        //      tensor core still works on half 1; F(1) runs under the next op's first four groups (they read chunks 0..3 only) --
        //      with four TMEM buffers the tensor core can be that far ahead of the promotions.
        if (op <= 7) {
          // stage this op's biases in shared memory (L1 is ~0 KB next to 226 KB of shared memory: a global load in the finalize
          // loop is an exposed L2 round trip)
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          bias_s[e_w * 32 + lane] = __ldg(P.bias + op * HM_HIDDEN + e_w * 32 + lane);
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        }
        promote(I0, I1);
        if (narrow) {
#pragma unroll 1
          for (int st = 0; st < 3; ++st) promote(I0, I0);
        } else {
          promote(I1, I1);
          if (wide) { promote(I0, I0); promote(I1, I0); promote(I0, I0); promote(I0, I0); }
          finalize(I0, I0, op, k_mul, unscale, s_next, m0);
          if (wide) { promote(I1, I0); promote(I1, I0); }
          finalize(I1, I0, op, k_mul, unscale, s_next, m1);
          if (kJac && op < 7) *reinterpret_cast<uint2*>(my_masks + (size_t)op * kMaskStride) = make_uint2(m0, m1);
        }
        if (op == 7) {
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");     // every warp is done with bias_s (same memory)
          dot_scratch[p * 8 + g8] = dot;
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          const float b8 = __ldg(P.b8);
          const float4 d0 = *reinterpret_cast<const float4*>(dot_scratch + p * 8), d1 = *reinterpret_cast<const float4*>(dot_scratch + p * 8 + 4);
          f_out = tanhf((((d0.x + d0.y) + (d0.z + d0.w)) + ((d1.x + d1.y) + (d1.z + d1.w))) + b8);
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          dot = 0.f;
          if (g8 == 0 && ok) P.sdf[grow] = f_out;
          if (kJac) {
            // d7 = (1 - f^2) * w8 * relu'(h7): A operand of B7
            const float c7 = (1.f - f_out * f_out) * s_next;
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
              const uint32_t m_ = nh ? m1 : m0;
              const float* w8 = P.w8 + col0_of(nh);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(w8 + 8 * u)), w1 = __ldg(reinterpret_cast<const float4*>(w8 + 8 * u + 4));
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                float r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = ((m_ >> (8 * u + i)) & 1u) ? c7 * w[i] : 0.f;
                emit_unit(nh, u, r);
              }
              publish(2 * nh + (cq >> 1));
            }
          }
        } else if (op == 15) {
          // ---------------- B0: g = d0 W0 (35 valid of 64 columns: TMEM columns 0..31 of lanes 0..63 hold columns 0..31, of
          //                  lanes 64..127 columns 32..63) + the parked skip gradient
          if (cq == 0 && ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = 32 * hq + j;
              if (col < HM_IN) {
                const float a = (j & 1) ? acc[0][j >> 1].y : acc[0][j >> 1].x;
                float* q = P.jac + grow * HM_IN + col;
                *q = fmaf(a, unscale, __ldcg(q));
              }
            }
          }
        }
      }
    }
    if (sat | (int)((sat2 & 0xffffu) >= 0x7bffu) | (int)((sat2 >> 16) >= 0x7bffu)) atomicAdd(P.flags, 1);
#ifdef HM_TC_COUNTERS
    if (P.flags && e_w == 0 && lane == 0) {
      atomicAdd((unsigned long long*)(P.flags + 18 + 8 * rank), (unsigned long long)t_pfull);
      atomicAdd((unsigned long long*)(P.flags + 20 + 8 * rank), (unsigned long long)t_pbody);
      atomicAdd((unsigned long long*)(P.flags + 22 + 8 * rank), (unsigned long long)t_fin);
      atomicAdd((unsigned long long*)(P.flags + 24 + 8 * rank), (unsigned long long)(clock64() - t_epi_begin));
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // the peer's TMEM / shared memory stay valid until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ bring-up self test
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

// ------------------------------------------------------------------ host side: plan + weight blob
float pow2_floor(float x) { return std::exp2(std::floor(std::log2(x))); }

__half f2h(float x) { return __float2half_rn(x); }

// Fill one K-major SW128 tile of `rows` x 64 with the fp16 hi (part = 1) or lo (part = 0) half of B(n, k) * scale.
template <class F>
void fill_tile(uint8_t* dst, int rows, float scale, int part, F&& get) {
  for (int n = 0; n < rows; ++n)
    for (int k = 0; k < 64; ++k) {
      float x = get(n, k) * scale;
      __half h = f2h(x);
      __half l = f2h(x - __half2float(h));
      memcpy(dst + sw128_offset(n, k), part ? &h : &l, 2);
    }
}

}  // namespace

static int hm_tc_blob_copies() {
  const char* e = getenv("HM_TC_BLOB_COPIES");
  int c = e ? atoi(e) : 1;
  return c < 1 ? 1 : (c > 8 ? 8 : c);
}

int hm_tc_init(hm_context* ctx) {
  // op list: F0..F7 (lin0..lin7), B7..B1, B0
  hm_tc_plan& plan = ctx->tc_plan;
  std::vector<uint8_t> blob;
  auto wmax = [&](int l) {
    float m = 0.f;
    for (float v : ctx->h_W[l]) m = std::max(m, std::fabs(v));
    return std::max(m, 1e-20f);
  };
  for (int op = 0; op < HM_TC_NOPS_ALL; ++op) {
    const bool fwd = op < 8;
    const int l = fwd ? op : 15 - op;
    hm_tc_op& o = plan.ops[op];
    o.n_kchunks = (op == 0) ? 1 : 8;
    o.n_nblocks = (op == 15) ? 1 : 2;                // 256-column output halves (B0: one 64-column block)
    o.stage_rows = (op == 15) ? 64 : 256;
    o.pad_ = 0;
    const float amax = std::max(ctx->act_absmax[op], 1e-20f);
    o.in_scale = pow2_floor(1024.f / amax);          // 64x headroom below the fp16 maximum
    const float w_scale = pow2_floor(8192.f / wmax(l));
    o.out_unscale = 1.f / (o.in_scale * w_scale);
    o.blob_offset = (int64_t)blob.size();
    const std::vector<float>& W = ctx->h_W[l];
    const int in_dim = ctx->in_dim[l];
    const size_t tile_bytes = (size_t)o.stage_rows * 128;
    for (int g = 0; g < groups_of(o.n_kchunks, o.n_nblocks); ++g) {    // stages in the MMA warp's consumption order
      int step, nh;
      group_of(o.n_kchunks, o.n_nblocks, g, step, nh);
        for (int part = 0; part < 2; ++part)           // lo tiles of the step's chunks first, then the hi tiles
          for (int which = 0; which < ((o.n_kchunks == 1) ? 1 : 2); ++which) {
            const int chunk = (o.n_kchunks == 1) ? 0 : chunk_of(step, which);
            size_t at = blob.size();
            blob.resize(at + tile_bytes, 0);
            fill_tile(blob.data() + at, o.stage_rows, w_scale, part, [&](int n, int k) -> float {
              const int gn = nh * o.stage_rows + n, gk = chunk * 64 + k;
              if (fwd) {                       // B[n][k] = W_l[n][k]
                if (gk >= in_dim) return 0.f;
                return W[(size_t)gn * in_dim + gk];
              }
              // backward: D[row][i] = sum_o d[row][o] W_l[o][i]  ->  B[n = i][k = o] = W_l[o][i]
              if (gn >= in_dim) return 0.f;
              return W[(size_t)gk * in_dim + gn];
            });
          }
    }
  }
  const int copies = hm_tc_blob_copies();
  if (ctx->d_tc_blob && (ctx->tc_blob_bytes != blob.size() || ctx->tc_blob_copies != copies)) { cudaFree(ctx->d_tc_blob); ctx->d_tc_blob = nullptr; }
  if (!ctx->d_tc_blob) HM_CUDA(cudaMalloc(&ctx->d_tc_blob, blob.size() * copies));
  ctx->tc_blob_bytes = blob.size();
  ctx->tc_blob_copies = copies;
  for (int c = 0; c < copies; ++c)
    HM_CUDA(cudaMemcpy(ctx->d_tc_blob + (size_t)c * blob.size(), blob.data(), blob.size(), cudaMemcpyHostToDevice));
  if (!ctx->d_tc_bias) {
    HM_CUDA(cudaMalloc(&ctx->d_tc_bias, sizeof(float) * 8 * HM_HIDDEN));
    HM_CUDA(cudaMalloc(&ctx->d_tc_masks, sizeof(uint32_t) * (size_t)ctx->sm_count * 8 * kMaskWordsPerOp));
    HM_CUDA(cudaMalloc(&ctx->d_tc_flags, sizeof(int32_t) * 64));
    HM_CUDA(cudaMemset(ctx->d_tc_flags, 0, sizeof(int32_t) * 64));
    HM_CUDA(cudaFuncSetAttribute(tc_decoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    HM_CUDA(cudaFuncSetAttribute(tc_decoder_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  }
  std::vector<float> bias(8 * HM_HIDDEN, 0.f);
  for (int l = 0; l < 8; ++l) {
    memcpy(bias.data() + l * HM_HIDDEN, ctx->h_b[l].data(), sizeof(float) * ctx->h_b[l].size());
    if (l < 7)                                       // the F_l epilogue emits h_l * in_scale(F_{l+1}): fold the scale into the bias
      for (int c = 0; c < HM_HIDDEN; ++c) bias[l * HM_HIDDEN + c] *= plan.ops[l + 1].in_scale;
  }
  HM_CUDA(cudaMemcpy(ctx->d_tc_bias, bias.data(), sizeof(float) * bias.size(), cudaMemcpyHostToDevice));
  return HM_OK;
}

void hm_tc_free(hm_context* ctx) {
  if (ctx->d_tc_blob) cudaFree(ctx->d_tc_blob);
  if (ctx->d_tc_bias) cudaFree(ctx->d_tc_bias);
  if (ctx->d_tc_masks) cudaFree(ctx->d_tc_masks);
  if (ctx->d_tc_flags) cudaFree(ctx->d_tc_flags);
  if (ctx->d_tc_trace) cudaFree(ctx->d_tc_trace);
  ctx->d_tc_trace = nullptr;
  ctx->d_tc_blob = nullptr;
  ctx->d_tc_bias = nullptr;
  ctx->d_tc_masks = nullptr;
  ctx->d_tc_flags = nullptr;
}

int hm_tc_decode(hm_context* ctx, const hm_rows& rows, float* d_sdf, float* d_jac, cudaStream_t st) {
  HM_CHECK(ctx->d_tc_blob, "tensor-core engine not initialised");
  TcParams P;
  P.plan = ctx->tc_plan;
  P.blob = ctx->d_tc_blob;
  P.blob_bytes = (int64_t)ctx->tc_blob_bytes;
  P.blob_copies = ctx->tc_blob_copies;
  P.bias = ctx->d_tc_bias;
  P.w8 = ctx->d_W[8];
  P.b8 = ctx->d_b[8];
  P.rows = rows.d_rows;
  P.xyz = rows.d_xyz;
  P.latents = rows.d_latents;
  P.row_latent = rows.d_row_latent;
  P.n_dynamic = rows.d_n_dynamic;
  P.n = rows.n;
  P.sdf = d_sdf;
  P.jac = d_jac;
  P.masks = reinterpret_cast<uint32_t*>(ctx->d_tc_masks);
  P.flags = ctx->d_tc_flags;
  P.trace = ctx->d_tc_trace;
  P.grid_n = rows.grid_n;
  P.grid_voxel = rows.grid_voxel;
  P.grid_radius = rows.grid_radius;
  const int64_t n_tiles = (rows.n + HM_TC_TILE_M - 1) / HM_TC_TILE_M;
  const int csize = 2;                                              // the kernel is a CTA-pair kernel
  const int64_t n_units = (n_tiles + csize - 1) / csize;             // a unit = one 64-row tile per CTA of the cluster
  const int grid = (int)std::min<int64_t>(n_units, ctx->sm_count / csize) * csize;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (d_jac) HM_CUDA(cudaLaunchKernelEx(&cfg, tc_decoder_kernel<true>, P));
  else HM_CUDA(cudaLaunchKernelEx(&cfg, tc_decoder_kernel<false>, P));
  ctx->counters.kernel_launches += 1;
  HM_CUDA(cudaGetLastError());
  return HM_OK;
}

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
  HM_CUDA(cudaMemcpy(out, ctx->d_tc_flags + 8, sizeof(unsigned long long) * 13, cudaMemcpyDeviceToHost));
  HM_CUDA(cudaMemset(ctx->d_tc_flags + 8, 0, sizeof(unsigned long long) * 13));
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
