// Tensor-core decoder engine (HM_ENGINE_TC): the whole DeepSDF MLP -- forward and, optionally, the input gradient -- for
// tiles of 64 query rows per CTA in ONE persistent, warp-specialised sm_100a kernel launched as clusters of two CTAs.
//
// Restates deepsdf/networks/deep_sdf_decoder.py:75-110 (forward) and the autograd input gradient of
// wild_completion/utils.py:112-122,175-193.  DESIGN.md section 4.1 has the full description and the measurements; in short:
//
//   * every layer is a tcgen05.mma GEMM with fp32 accumulators in TMEM.  fp32 parity needs more than one fp16/bf16 MMA
//     (SURVEY.md 7.3): operands are split x*s = hi + lo (fp16 each, s a calibrated power of two) and exactly the three needed
//     products A_hi x W_lo + A_lo x W_hi + A_hi x W_hi are issued -- measured 3e-8 abs SDF error, fp32 grade.
//   * a CTA owns a tile of 64 points; the CTA pair issues cta_group::2 MMAs of M = 128 (64 rows per CTA, full rate) x N = 256:
//     the hi and the lo parts of the points are two separate 64-row A tiles whose products accumulate into the SAME TMEM rows,
//     the B operand is split between the two CTAs' shared memories, each CTA streams half of every weight tile
//     (cp.async.bulk + mbarrier complete_tx, 6 x 16 KB ring; stages are stored in the blob pre-swizzled, in consumption order).
//   * activations never leave the SM: the epilogue warps read partial accumulators from TMEM, apply bias/ReLU (or the ReLU
//     mask in the backward pass), re-split to fp16 hi/lo and write the next layer's A operand straight into shared memory
//     in the 128-byte-swizzled K-major UMMA layout, in place.
//   * the tensor core accumulates fp32 with round-toward-zero (measured: -9e-6 relative after the 96 chained MMAs of one
//     layer), so MMAs are chained only inside an accumulation group (2 k-chunks x one 256-column output half, 24 MMAs) into
//     one of four TMEM buffers, and the group partials are summed in registers in fp32 round-to-nearest.
//   * the backward pass is seeded with w8 * relu'(h7) as soon as lin7's output half is final; the tanh' factor (1 - sdf^2) is a
//     per-row scalar and multiplies the finished gradient, so lin8 + tanh are off the tensor core's critical path.
//   * sparse plan: the shipped DeepSDF models are extremely sparse after their ReLUs (lin3 is dead outright -- the network lives
//     on the skip connection -- and only 12 .. 320 of the 512 units of the other layers are ever alive).  hm_calibrate records
//     which units were alive, the hidden units are permuted so that those come first, and the plan drops every MMA whose A
//     operand is then an all-zero 64-wide k-chunk (and every output half nobody reads in the gradient pass).  The assumption is
//     CHECKED, not trusted: each forward layer's epilogue looks at the ReLU bits of the columns the plan counts on being zero and
//     a tile that contradicts it is queued and re-evaluated by a second launch of the same kernel with the full plan.  The
//     dropped products are exact zeros, so sparse == full bit for bit.
//   * three modes (template kMode): forward only (optionally storing every row's ReLU bits, 512 B per row), forward + gradient,
//     and gradient ONLY from stored bits + SDF values (the joint loop differentiates a subset of the rows it has just evaluated,
//     loss.py:185-215; the second forward evaluation the reference pays for them is not needed).
//   * the MMA issuer, the weight producer and the peer's arrival forwarder walk a flat STAGE PROGRAM (hm_tc_plan::rec, one 32-bit
//     record per issued weight stage, built on the host): the issuing warp's instruction stream is on the tile's critical path.
//     F0's operand has a barrier of its own (X0_READY) and, when the plan leaves a chunk unread during the tile's last op, is
//     written for the NEXT tile while that op runs, so F0's MMAs follow the last op's directly.
//   * lin8's weight and the eight bias vectors travel in the kernel-parameter constant bank: every lane of an epilogue warp reads
//     the same columns, so they arrive at register speed without shared-memory staging or barriers.
//
// Warp roles (640 threads): warpgroup 0 = control (warp 0 bulk-copy producer, warp 1 MMA issuer in the leader CTA / weight
// arrival forwarder in the peer CTA + TMEM alloc, warps 2-3 idle), warpgroups 1-4 = 16 epilogue warps (4 per TMEM
// sub-partition, 64 output columns of each half per warp); setmaxnreg moves registers from the control warpgroup to them.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

using namespace hm_tc;

namespace {

constexpr int kWeightRing = 98304;                 // shared memory of the weight ring: 6 x 16 KB
constexpr int kMaxStages = 6;
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

// barrier slots (8 bytes each) inside the 256-byte control block at kSmemBars, followed by a few control words
enum { BAR_W_FULL = 0, BAR_W_EMPTY = kMaxStages, BAR_A_READY = 2 * kMaxStages, BAR_PART_FULL = BAR_A_READY + 4,
       BAR_PART_EMPTY = BAR_PART_FULL + 4, BAR_X0_READY = BAR_PART_EMPTY + 4, BAR_COUNT = BAR_X0_READY + 1 };
constexpr int kCtlTmemPtr = 8 * BAR_COUNT;         // TMEM base column written by tcgen05.alloc
constexpr int kCtlViolated = kCtlTmemPtr + 4;      // set by an epilogue warp whose tile contradicted the sparse plan
static_assert(kCtlViolated + 4 <= 256, "control block overflows into the bias / dot-product scratch");

struct TcParams {
  hm_tc_plan plan;
  const uint8_t* blob;
  float b8;                   // lin8 bias
  const float* rows;          // [n][35] or null
  const float* xyz;           // [n][3]
  const float* latents;       // [L][32]
  const int32_t* row_latent;  // [n] or null
  const int32_t* n_dynamic;   // device row count or null
  int64_t n;
  float* sdf;
  const int32_t* out_index;   // optional: sdf of row i is written to sdf[out_index[i]] (compacted ray samples) or null
  float* jac;
  uint32_t* masks;            // [grid][8][64][16]
  int32_t* flags;             // device counters, HM_TC_FLAG_* slots (common.cuh)
  int32_t* latent_sat;        // optional [L]: set to 1 for the latent-table rows (fruits) one of whose rows saturated an fp16 operand
  uint32_t* trace;            // timeline of the first CTA pair (testing build only) or null
  int32_t grid_n;             // > 0: xyz of row i = voxel grid point i (fused mesher grid, hm_rows)
  float grid_voxel, grid_radius;
  uint32_t* mask_out;         // forward-only pass: optional per-TILE store of the 8 layers' ReLU masks, [tile][8][512 threads][2 words]
  const uint32_t* mask_in;    // backward-only pass (kMode 2): the masks a forward-only pass stored, and ...
  const int32_t* src_row;     // ... [n] the row index each row had in THAT pass (its tile = src / 64, its point = src % 64)
  int32_t* redo;              // [0] = number of queued tiles, [4 ..] = tile indices: appended by the sparse pass, read by the redo pass
  // lin8 weight (permuted unit order).  Every lane of an epilogue warp reads the SAME columns, so the kernel-parameter constant
  // bank serves it at register speed; a global load in the finalize loop is an exposed L2 round trip (the L1 is ~0 KB here).
  alignas(16) float w8[HM_HIDDEN];
  // ... and so do the biases of lin0..7 ([8][512], pre-scaled for the next op): staging them in shared memory cost two CTA-wide
  // barriers and one exposed L2 round trip per forward op
  alignas(16) float bias[8 * HM_HIDDEN];
};
static_assert(sizeof(TcParams) <= 32764, "kernel parameters exceed the 32 KB limit of sm_70+ (CUDA >= 12.1)");

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
// of M = 128 (64 rows per CTA -- measured: full rate, 64 cycles for N = 256, scripts/probe_decoder.py pair) whose B operand is
// split between the two CTAs' shared memories, so each CTA streams only HALF of every weight tile from L2.  Per weight
// tile pair the THREE needed products are issued -- A_hi x W_lo, A_lo x W_hi, A_hi x W_hi -- into the same accumulator
// rows (the 64 x N accumulator of each CTA is folded onto 128 lanes x N/2 columns: lanes 0..63 hold output columns
// [0, N/2), lanes 64..127 columns [N/2, N)).  CTA rank 0 (the leader) issues all MMAs.
//
// kRedo = false: tiles 2 * unit + rank, evaluated with P.plan (the sparse plan); a tile whose activations contradict the
// plan's zero-chunk assumptions is appended to P.redo.  kRedo = true: the tiles listed in P.redo, evaluated with P.plan = the
// full plan (launched right behind the first kernel; exits at once when the list is empty).
//
// kMode 0: forward only (optionally storing the ReLU masks per tile, P.mask_out).  kMode 1: forward + input gradient.
// kMode 2: input gradient ONLY, for rows a forward-only pass has already evaluated (the joint loop's in-band ray samples,
// loss.py:185-215): the backward pass needs nothing of the forward pass but the ReLU bits and the SDF value, so the tile starts at
// B7 with d7 = w8 * relu'(h7) rebuilt from the stored bits -- the same operand F7's epilogue writes in kMode 1, hence the same bits
// out -- and P.sdf is an input.
template <int kMode, bool kRedo>
__global__ void __launch_bounds__(kThreads, 1) tc_decoder_kernel(const __grid_constant__ TcParams P) {
  constexpr bool kJac = kMode != 0;
  constexpr bool kBwd = kMode == 2;
  constexpr int kOpBegin = kBwd ? 8 : 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int CG = 2;
  constexpr int kStages = kMaxStages;
  constexpr int kStageBytes = kWeightRing / kStages;      // 128 output features x 64 k x 2 B (hi OR lo) = this CTA's half of a tile
  constexpr bool kPair = true;
  const int64_t n_rows = P.n_dynamic ? (int64_t)min((int64_t)*P.n_dynamic, P.n) : P.n;
  int64_t n_tiles = (n_rows + HM_TC_TILE_M - 1) / HM_TC_TILE_M;
  if (kRedo) {
    n_tiles = P.redo[0];                                  // tiles queued by the sparse pass (uniform over the grid)
    if (n_tiles == 0) return;
  }
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;    // warp index as a uniform value
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + kSmemBars;
  auto bar = [&](int i) { return bars + 8u * i; };
  volatile uint32_t* const ctl = reinterpret_cast<volatile uint32_t*>(smem + kSmemBars);        // control words (byte offsets kCtl*)
  volatile uint32_t* tmem_ptr_smem = ctl + kCtlTmemPtr / 4;
  float* dot_scratch = reinterpret_cast<float*>(smem + kSmemBars + 256);      // [64 points][8 column groups]: partial lin8 dot products
  constexpr int kOps = kJac ? HM_TC_NOPS_ALL : HM_TC_NOPS_FWD;
  const int last_op = kJac ? P.plan.last_op_jac : P.plan.last_op_fwd;         // last executed op of a tile
  const uint32_t rank = cluster_ctarank();                                    // 0 = leader (MMA issuer) of the pair
  const uint32_t lead_bars = mapa_rank(bars, 0);                              // the leader's barrier block (cluster address)
  auto lead_bar = [&](int i) { return lead_bars + 8u * i; };
  const int64_t unit0 = blockIdx.x >> 1, unit_stride = gridDim.x >> 1;
  const int rec_begin = kBwd ? P.plan.n_rec_fwd : 0, rec_end = kJac ? P.plan.n_rec_all : P.plan.n_rec_fwd;     // this mode's part of the stage program
#ifdef HM_TESTING
  // timeline of the first CTA pair: (code << 24 | op << 16 | index, clock) pairs; region 0 = MMA issuer, 1 / 2 = first
  // epilogue warp of the leader / peer CTA (scripts/probe_decoder.py trace)
  constexpr uint32_t kTraceCap = 8192;
  const bool tracing = P.trace != nullptr && blockIdx.x < 2 && !kRedo;
  uint32_t trace_n = 0;
  auto trace = [&](int region, uint32_t code, uint32_t op, uint32_t idx) {
    if (tracing && trace_n < kTraceCap) {
      uint32_t* t = P.trace + (size_t)region * kTraceCap * 2 + 2 * trace_n;
      t[0] = (code << 24) | (op << 16) | idx;
      t[1] = (uint32_t)clock64();
      ++trace_n;
    }
  };
#define HM_TRACE_OP(...) trace(__VA_ARGS__)          // one event per op (MMA issuer): cheap enough to leave the timing alone
#ifdef HM_TC_LIGHT
#define HM_TRACE(...) ((void)0)
#else
#define HM_TRACE(...) trace(__VA_ARGS__)
#endif
#else
#define HM_TRACE(...) ((void)0)
#define HM_TRACE_OP(...) ((void)0)
#endif

  if (threadIdx.x == 0) {
    // W_FULL of the leader also collects the peer's "my half has landed" arrive
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(BAR_W_FULL + s), rank == 0 ? 2 : 1); mbar_init(bar(BAR_W_EMPTY + s), 1); }
    for (int p = 0; p < 4; ++p) mbar_init(bar(BAR_A_READY + p), kEpiWarps * CG / 2);      // a k-step's chunks come from half of the warps
    mbar_init(bar(BAR_X0_READY), kEpiWarps * CG);                                          // F0's operand: every warp writes 8 of its 64 columns
    for (int b = 0; b < kBufs; ++b) { mbar_init(bar(BAR_PART_FULL + b), 1); mbar_init(bar(BAR_PART_EMPTY + b), kEpiWarps * CG); }
    ctl[kCtlViolated / 4] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<CG>(smem_u32((const void*)tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (lane == 0 && (warp == 1 || warp == kCtrlWarps)) HM_TRACE(warp == 1 ? 0 : 1 + rank, 0, 0, 0);     // common time origin
#ifdef HM_TESTING
  const long long t_cta_begin = clock64();
#endif

  if (!kRedo && blockIdx.x == 0 && threadIdx.x == 0) {     // exact row / tile accounting for the roofline (rows actually evaluated)
    atomicAdd(reinterpret_cast<unsigned long long*>(P.flags + (kBwd ? HM_TC_FLAG_ROWS_BWD : kJac ? HM_TC_FLAG_ROWS_JAC : HM_TC_FLAG_ROWS_FWD)), (unsigned long long)n_rows);
    atomicAdd(reinterpret_cast<unsigned long long*>(P.flags + (kBwd ? HM_TC_FLAG_TILES_BWD : kJac ? HM_TC_FLAG_TILES_JAC : HM_TC_FLAG_TILES_FWD)), (unsigned long long)n_tiles);
  }
  if (kRedo && blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(P.flags + (kBwd ? HM_TC_FLAG_DEAD_BWD : kJac ? HM_TC_FLAG_DEAD_JAC : HM_TC_FLAG_DEAD_FWD)), (unsigned long long)n_tiles);
  const int64_t n_units = (n_tiles + 1) / 2;               // a unit = one tile per CTA of the pair

  if (warp < kCtrlWarps) {
  // the control warpgroup hands most of its registers to the four epilogue warpgroups (64 accumulators per thread)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtrlRegs));
  if (warp == 0 || (warp == 1 && rank != 0)) {
    // ===================== warp 0: weight producer (every CTA fetches its half of each stage) =====================
    // ===================== warp 1 of the peer CTA: tells the leader that this CTA's half of a stage has landed =====================
    // Both walk the same stage sequence: the stages of the groups that the plan issues, in the blob's (consumption) order.
    const bool producer = warp == 0;
    uint32_t slot = 0, phase = 0;
    long long t_empty = 0;
    for (int64_t unit = unit0; unit < n_units; unit += unit_stride) {
      for (int i = rec_begin; i < rec_end; ++i) {
        const uint32_t r = P.plan.rec[i];
        const uint32_t bytes = (r & HM_TC_REC_NARROW) ? 64u * 128u / CG : 256u * 128u / CG;      // this CTA's rows of one 64-k fp16 tile
        if (producer) {
          mbar_wait_timed<false>(bar(BAR_W_EMPTY + slot), phase ^ 1, t_empty);
          if (elect_one()) {
            mbar_expect_tx(bar(BAR_W_FULL + slot), bytes);
            bulk_g2s(smem_base + kSmemStages + slot * kStageBytes, P.blob + (size_t)HM_TC_REC_SRC(r) * 8192u + (size_t)rank * bytes, bytes, bar(BAR_W_FULL + slot));
          }
        } else {
          mbar_wait(bar(BAR_W_FULL + slot), phase);
          if (lane == 0) mbar_arrive_cluster(lead_bar(BAR_W_FULL + slot));
        }
        __syncwarp();
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      }
    }
#ifdef HM_TC_COUNTERS
    if (producer && rank == 0 && lane == 0) atomicAdd((unsigned long long*)(P.flags + HM_TC_FLAG_DEBUG + 0), (unsigned long long)t_empty);
#endif
  } else if (warp == 1) {
      // ===================== MMA issuer (leader CTA; the whole warp walks the stage program, one elected lane issues) =====================
      // One accumulation group = (k-step = 2 k-chunks, 256-column output half): per chunk A_hi x W_lo (4 MMAs), then A_lo x W_hi
      // and A_hi x W_hi (8 MMAs), each M = 64 per CTA x N = 256 x K = 16, into a FRESH 128-column TMEM buffer (four buffers).
      // The tensor core accumulates fp32 with round-toward-zero (measured: -1e-7 relative per chained MMA), so chains are
      // kept to one group and the epilogue warps add the group partials in fp32 round-to-nearest.
      uint32_t slot = 0, phase = 0, a_seq = 0, x_seq = 0, gseq = 0, a_par = 0, steps_ready = 0;
      long long t_a = 0, t_part = 0, t_w = 0;
#ifdef HM_TC_COUNTERS
      const long long t_begin = clock64();
#endif
      // descriptors without their address fields; the A chunk / ring slot address is added per stage (16 KB = 1024 units of 16 B)
      const uint64_t a_desc0 = make_desc(smem_base + kSmemA), w_desc0 = make_desc(smem_base + kSmemStages);
      constexpr uint32_t idesc_wide = make_idesc(64 * CG, 256), idesc_narrow = make_idesc(64 * CG, 64);
      for (int64_t unit = unit0; unit < n_units; unit += unit_stride) {
        uint32_t r_next = P.plan.rec[rec_begin];
        for (int i = rec_begin; i < rec_end; ++i) {
          const uint32_t r = r_next;
          r_next = P.plan.rec[i + 1 < rec_end ? i + 1 : rec_begin];       // (the next record's constant-bank load overlaps this stage)
          if (r & HM_TC_REC_OP_FIRST) {
            if (HM_TC_REC_NEED_READY(r) == 7u) {   // F0: its operand has a barrier of its own, one phase per tile (it may have been
              steps_ready = 7;                     // written while the previous tile's last op ran); no A_READY phase is consumed
              mbar_wait_timed<kPair>(bar(BAR_X0_READY), x_seq & 1u, t_a);
              ++x_seq;
              tc_fence_after();
            } else {                               // A_READY completes one phase per executed op
              a_par = a_seq & 1u;
              ++a_seq;
              steps_ready = 0;
            }
          }
          {
            const uint32_t buf = gseq & (kBufs - 1);
            if (r & HM_TC_REC_GROUP_FIRST) {
              const uint32_t need = HM_TC_REC_NEED_READY(r);
              if (steps_ready < need) {
                for (; steps_ready < need; ++steps_ready) mbar_wait_timed<kPair>(bar(BAR_A_READY + steps_ready), a_par, t_a);
                tc_fence_after();
              }
              if (lane == 0) HM_TRACE(0, 1, HM_TC_REC_OP(r), HM_TC_REC_GROUP(r));
              mbar_wait_timed<kPair>(bar(BAR_PART_EMPTY + buf), ((gseq / kBufs) & 1) ^ 1, t_part);
              tc_fence_after();
              if (lane == 0) HM_TRACE(0, 2, HM_TC_REC_OP(r), HM_TC_REC_GROUP(r));
            }
            const uint32_t d = tmem_base + buf * 128;
            const uint64_t a_hi = a_desc0 + (uint64_t)(HM_TC_REC_CHUNK(r) * (kAChunkBytes >> 4));
            const uint64_t a_lo = a_hi + (kALoOffset >> 4);
            const uint64_t w_desc = w_desc0 + (uint64_t)(slot * (kStageBytes >> 4));
            const uint32_t idesc = (r & HM_TC_REC_NARROW) ? idesc_narrow : idesc_wide;
            mbar_wait_timed<kPair>(bar(BAR_W_FULL + slot), phase, t_w);
            tc_fence_after();
            if (elect_one()) {
              if (!(r & HM_TC_REC_PART)) {           // lo weight tile (small terms first); the group's first MMA overwrites the buffer
                umma_f16<CG>(d, a_hi, w_desc, idesc, (r & HM_TC_REC_GROUP_FIRST) ? 0u : 1u);
#pragma unroll
                for (int ks = 1; ks < 4; ++ks)       // +32 B per 16-wide k step = +2 in the descriptor's address field
                  umma_f16<CG>(d, a_hi + 2 * ks, w_desc + 2 * ks, idesc, 1u);
              } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_f16<CG>(d, a_lo + 2 * ks, w_desc + 2 * ks, idesc, 1u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_f16<CG>(d, a_hi + 2 * ks, w_desc + 2 * ks, idesc, 1u);
              }
              umma_commit<CG>(bar(BAR_W_EMPTY + slot));          // frees the slot in both CTAs of the pair
              if (r & HM_TC_REC_GROUP_LAST) umma_commit<CG>(bar(BAR_PART_FULL + buf));
            }
            __syncwarp();
            if (r & HM_TC_REC_GROUP_LAST) {
              if (lane == 0) HM_TRACE(0, 3, HM_TC_REC_OP(r), HM_TC_REC_GROUP(r));
            }
          }
          if (++slot == kStages) { slot = 0; phase ^= 1; }
          if (r & HM_TC_REC_GROUP_LAST) ++gseq;
          // every executed op completes one phase of all four A_READY barriers (the epilogue warps always publish all k-steps):
          // consume the ones whose groups the plan dropped, so that the phase parity stays in step
          if (r & HM_TC_REC_OP_LAST)
            for (; steps_ready < 4; ++steps_ready) mbar_wait_timed<kPair>(bar(BAR_A_READY + steps_ready), a_par, t_a);
        }
      }
#ifdef HM_TC_COUNTERS
      if (P.flags && lane == 0) {
        atomicAdd((unsigned long long*)(P.flags + HM_TC_FLAG_DEBUG + 2), (unsigned long long)t_a);
        atomicAdd((unsigned long long*)(P.flags + HM_TC_FLAG_DEBUG + 4), (unsigned long long)t_part);
        atomicAdd((unsigned long long*)(P.flags + HM_TC_FLAG_DEBUG + 6), (unsigned long long)t_w);
        atomicAdd((unsigned long long*)(P.flags + HM_TC_FLAG_DEBUG + 8), (unsigned long long)(clock64() - t_begin));
      }
#endif
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
    // this thread's slot in the plan's shared-memory mask chunk: [layer slot][thread] x 8 bytes (forward + gradient pass)
    uint8_t* const mask_smem = smem + kSmemA + (uint32_t)(P.plan.mask_chunk < 0 ? 0 : P.plan.mask_chunk) * kAChunkBytes + (uint32_t)(e_w * 32 + lane) * 8u;
    uint32_t gseq = 0;
    int sat = 0;
#ifdef HM_TC_COUNTERS
    long long t_pfull = 0, t_pbody = 0, t_fin = 0;       // wait-cycle counters (testing build)
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
    const int chunk_lo = 2 * hq + (cq >> 1);                  // this thread's output chunk within a half (add 4 * nh)
    uint8_t* const st_row = smem + kSmemA + (uint32_t)chunk_lo * kAChunkBytes + (uint32_t)(p >> 3) * 1024 + (p & 7) * 128;
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
    // Inputs of a tile: this thread's point has the global row `grow`; its latent-table row (one load) is fetched for the NEXT tile
    // while the tensor core works on the last op of the current one, and the lines its x0 values sit in are pulled into L2, so that
    // only L2 hits remain on the critical path at the tile start (a register prefetch of the values themselves does not fit next
    // to the 64 accumulators).
    auto tile_of = [&](int64_t unit) -> int64_t {             // tile index of this CTA in `unit`, -1 = padding
      const int64_t t = 2 * unit + rank;                      // the odd CTA of the last pair may get an all-padding tile
      if (t >= n_tiles) return -1;
      return kRedo ? (int64_t)__ldg(P.redo + 4 + t) : t;
    };
    auto row_of = [&](int64_t tile) -> int64_t {              // global row of this thread's point (clamped for padding rows)
      const int64_t g = tile * HM_TC_TILE_M + p;
      return (tile >= 0 && g < n_rows) ? g : n_rows - 1;
    };
    auto lat_ptr_of = [&](int64_t lr, int32_t li) -> const float* {   // the 32 latent values are contiguous in both input modes
      return P.rows ? P.rows + lr * HM_IN : P.latents + (size_t)li * HM_LATENT;
    };
    auto xyz_of = [&](int64_t lr, int c) -> float {           // xyz coordinate c of row lr
      if (P.rows) return __ldg(P.rows + lr * HM_IN + HM_LATENT + c);
      if (P.grid_n > 0) return hm_grid_coord(lr, c, P.grid_n, P.grid_voxel, P.grid_radius);
      return __ldg(P.xyz + lr * 3 + c);
    };
    auto latent_row_of = [&](int64_t lr) -> int32_t { return (!P.rows && P.row_latent) ? __ldg(P.row_latent + lr) : 0; };
    // A operand of F0: chunk x0_chunk = [x0 * s, 0 ...] (K padded 35 -> 64) of row lr; column group g8 writes k in [8*g8, +8).
    // Published on X0_READY (all 16 warps of both CTAs arrive once per tile).
    bool x0_written = false;
    int sat_next = 0;                            // saturation seen while writing the NEXT tile's operand (attributed to that tile's fruit)
    auto write_x0 = [&](int64_t lr_x, const float* lat_x, int& sat_x) {
      const float s0 = P.plan.ops[0].in_scale;
      float xin[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xin[i] = 0.f;
      if (g8 < 4) {                              // 8 independent loads: one memory latency, not eight
#pragma unroll
        for (int i = 0; i < 8; ++i) xin[i] = __ldg(lat_x + 8 * g8 + i);
      } else if (g8 == 4) {
#pragma unroll
        for (int c = 0; c < 3; ++c) xin[c] = xyz_of(lr_x, c);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) store_pair(smem, P.plan.x0_chunk, p, 8 * g8 + 2 * e, xin[2 * e] * s0, xin[2 * e + 1] * s0, sat_x);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_bar(BAR_X0_READY));
    };
    int32_t cur_li = (unit0 < n_units) ? latent_row_of(row_of(tile_of(unit0))) : 0, nxt_li = 0;
    for (int64_t unit = unit0; unit < n_units; unit += unit_stride) {
      const int64_t tile = tile_of(unit);
      const int64_t grow = (tile < 0 ? 0 : tile) * HM_TC_TILE_M + p;
      const bool ok = tile >= 0 && grow < n_rows;
      const int64_t lr = ok ? grow : n_rows - 1;
      const float* const lat_ptr = lat_ptr_of(lr, cur_li);
      float f_out = 0.f, c7 = 0.f;
      bool viol = false;                     // a forward op found a live unit where the sparse plan assumes zeros
      // where this tile's ReLU masks are written (forward ops) and read (backward ops): [op][thread][2 words]
      uint32_t* mask_wr = my_masks;
      const uint32_t* mask_rd = my_masks;
      if (kMode == 0 && P.mask_out && tile >= 0) mask_wr = P.mask_out + (size_t)tile * 8 * kMaskStride + (size_t)(e_w * 32 + lane) * 2;
      if constexpr (!kBwd) {
        // ---- A operand of F0 (unless it was written while the previous tile's last op ran, see below)
        if (!x0_written) write_x0(lr, lat_ptr, sat);
        x0_written = false;
      } else {
        // ---- backward-only tile: the masks of this thread's point sit where the forward-only pass stored them -- in the slot of
        //      the thread that owned the SAME column group of that point there (sub-partition 2 * hq + point / 32, lane point % 32)
        const int32_t src = __ldg(P.src_row + lr);
        const int32_t sp_src = 2 * hq + ((src & 63) >> 5);
        mask_rd = P.mask_in + (size_t)(src >> 6) * 8 * kMaskStride + (size_t)((4 * cq + sp_src) * 32 + (src & 31)) * 2;
        uint2 mk[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) mk[l] = __ldg(reinterpret_cast<const uint2*>(mask_rd + (size_t)l * kMaskStride));     // 8 independent loads
        f_out = __ldg(P.sdf + lr);
        c7 = 1.f - f_out * f_out;                // tanh' (deep_sdf_decoder.py:107-108)
#pragma unroll
        for (int l = 0; l < 8; ++l) {            // the sparse plan's assumptions, checked on the stored bits
          const uint32_t va = P.plan.ops[l].verify_alive;
          if ((!((va >> chunk_lo) & 1u) && mk[l].x != 0u) || (!((va >> (4 + chunk_lo)) & 1u) && mk[l].y != 0u)) viol = true;
        }
        // A operand of B7: d7 = w8 * relu'(h7) (what F7's epilogue writes when forward and gradient run in one pass)
        const float s_in = P.plan.ops[8].in_scale;
        const uint32_t need7 = P.plan.ops[7].need_out;
#pragma unroll
        for (int nh = 0; nh < 2; ++nh) {
          if ((need7 >> (4 * nh + chunk_lo)) & 1u) {
            const uint32_t mbits = nh ? mk[7].y : mk[7].x;
            const float* w8 = P.w8 + col0_of(nh);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float r[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = ((mbits >> (8 * u + i)) & 1u) ? w8[8 * u + i] * s_in : 0.f;
              emit_unit(nh, u, r);
            }
          }
          publish(2 * nh + (cq >> 1));
        }
      }
      // Accumulators of this thread: acc[nh][i] = columns col0_of(nh) + 2i + {0, 1}.
      float2 acc[2][16];
      uint32_t m0 = 0u, m1 = 0u;             // ReLU bits of the current op: output half 0 / half 1, bit j = column col0 + j
      float dot = 0.f;
      const std::integral_constant<int, 0> I0{};
      const std::integral_constant<int, 1> I1{};
#pragma unroll 1
      for (int op = kOpBegin; op < kOps; ++op) {
        const hm_tc_op& o = P.plan.ops[op];
        const uint32_t gm = o.group_mask;
        if (gm == 0u) continue;                          // dropped by the plan (its A operand is exactly zero)
        const float unscale = o.out_unscale;
        const float s_next = (op + 1 < kOps) ? P.plan.ops[op + 1].in_scale : 1.f;
        const float k_mul = unscale * s_next;
        const bool narrow = (o.stage_rows == 64);        // B0: 64 output columns (32 TMEM columns), one group per step
        const bool wide = (o.n_kchunks != 1);
        const bool fwd_op = op < 8;
        if (!kBwd && op == last_op && unit + unit_stride < n_units) {
          // next tile.  If F0's operand chunk is free during this op (plan.x0_early) it is written NOW -- every reader of that chunk in
          // this tile has completed (all earlier ops' partials have been promoted) and this op's epilogue writes no A operand -- so
          // that F0's MMAs follow this op's MMAs directly and the tile hand-over (final epilogue, SDF / Jacobian stores, input loads)
          // leaves the op chain's critical path.  Otherwise: latent-table row now, x0 lines into L2.
          const int64_t nr = row_of(tile_of(unit + unit_stride));
          nxt_li = latent_row_of(nr);
          if (P.plan.x0_early) {
            write_x0(nr, lat_ptr_of(nr, nxt_li), sat_next);
            x0_written = true;
          } else {
            if (!P.rows && P.grid_n == 0 && g8 == 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.xyz + nr * 3));
            if (P.rows && g8 <= 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.rows + nr * HM_IN + 8 * g8));
          }
        }
        // which of this thread's two column blocks the epilogue has to produce (all of them for a forward op: its ReLU bits are the
        // check of the plan's assumptions; for a backward op only the columns somebody reads)
        const uint32_t need = o.need_out;
        const bool need0 = fwd_op || ((need >> chunk_lo) & 1u), need1 = fwd_op || ((need >> (4 + chunk_lo)) & 1u);
        const bool emit0 = (need >> chunk_lo) & 1u, emit1 = (need >> (4 + chunk_lo)) & 1u;
        m0 = m1 = 0u;
        if (kJac && op >= 8 && op < 15 && !o.is_last) {
          // ReLU mask of h_{l-1}, l = 15 - op
          const uint2 mw = (kMode == 1 && P.plan.mask_chunk >= 0)
                               ? *reinterpret_cast<const uint2*>(mask_smem + 4096u * (uint32_t)__popc(P.plan.mask_layers & ((1u << (14 - op)) - 1u)))
                               : *reinterpret_cast<const uint2*>(mask_rd + (size_t)(14 - op) * kMaskStride);
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
          if (e_w == 0 && lane == 0) HM_TRACE(1 + rank, 10, op, gseq & 0xffff);
#ifdef HM_TC_LIGHT
          if constexpr (first && nh == 0) { if (e_w == 0 && lane == 0) HM_TRACE_OP(1 + rank, 10, op, gseq & 0xffff); }
#endif
          auto release = [&]() {               // all TMEM reads of this buffer are complete
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_bar(BAR_PART_EMPTY + buf));
          };
          if ((narrow && cq != 0) || !(nh ? need1 : need0)) {      // B0's 64 output columns occupy TMEM columns 0..31 only; unread columns are skipped
            release();
          } else {
            // two 16-column halves: 16 live registers next to the 64 accumulators instead of 32 (the kernel has no L1 to speak of
            // -- 227 KB of the SM's 256 KB are shared memory -- so every spilled register is an L2 round trip on the op chain)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float v[16];
              tmem_ld16_nowait(t_addr + buf * 128 + 16 * h, v);
              tmem_ld_wait_dep16(v);
              if (h == 1) release();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 w = make_float2(v[2 * i], v[2 * i + 1]);
                if constexpr (first) acc[nh][8 * h + i] = w; else acc[nh][8 * h + i] = add2(acc[nh][8 * h + i], w);
              }
            }
          }
          ++gseq;
#ifdef HM_TC_COUNTERS
          t_pbody += clock64() - tp1;
#endif
          if (e_w == 0 && lane == 0) HM_TRACE(1 + rank, 12, op, gseq & 0xffff);
        };
        // Turn the finished output half nh of op `opx` into the next op's A chunks 4*nh .. 4*nh+3 (k-steps 2*nh, 2*nh+1) or the
        // final outputs.
        auto finalize = [&](auto NH, const int opx, const float k_mul_x, const float unscale_x, const float s_next_x, uint32_t& m_) {
          constexpr int nh = decltype(NH)::value;
#ifdef HM_TC_COUNTERS
          const long long tf0 = clock64();
#endif
          if (e_w == 0 && lane == 0) HM_TRACE(1 + rank, 13, opx, nh);
          const int col0 = col0_of(nh);
          // Skip concat (deep_sdf_decoder.py:87-88): columns 480..511 of lin4's input are x0[3..34].  The eight threads that own
          // point p (one per column group g8) each load four of them -- ONE round trip to L2 at the start of lin3's last finalize,
          // hidden under its ReLU loop -- and write them into chunk 7 below; the thread that owns those columns of lin3's (padded,
          // all-zero) output does not emit them.  (A single thread loading all 32 values serialised the loads under register
          // pressure: lin3 + lin4 took 12 k / 21 k cycles of a forward / forward + gradient tile against 3 k of MMA work.)
          const bool skip_cols = nh == 1 && opx == 3;
          float xs[4] = {0.f, 0.f, 0.f, 0.f};
          if (skip_cols) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = 3 + 4 * g8 + i;
              xs[i] = j < HM_LATENT ? __ldg(lat_ptr + j) : xyz_of(lr, j - HM_LATENT);
            }
          }
          const bool emit = (nh ? emit1 : emit0) && !(skip_cols && hq == 1 && cq == 3);      // does the next op multiply this thread's chunk?
          const bool may_live = (o.verify_alive >> (4 * nh + chunk_lo)) & 1u;      // may this thread's chunk hold non-zeros (forward ops)?
          if (opx < 7) {
            // ---------------- forward hidden layer: h = relu(acc + b); next A = h * s_next (bias pre-scaled by s_next)
            const float* bias = P.bias + opx * HM_HIDDEN + col0;
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
              if (emit) emit_unit(nh, u, r);
            }
            if (skip_cols) {
              // columns 480..511 of lin4's input = x0[3..34]: this thread's four values (loaded above), two fp16 pairs of point p
              store_pair(smem, 7, p, 32 + 4 * g8, xs[0] * s_next_x, xs[1] * s_next_x, sat);
              store_pair(smem, 7, p, 34 + 4 * g8, xs[2] * s_next_x, xs[3] * s_next_x, sat);
            }
            if (nh == 1 && opx == 3 && hq == 1 && cq >= 2) {
              // lin3 has 477 outputs; columns 477..511 of the next input are the raw x0 (skip concat, deep_sdf_decoder.py:87-88).
              // Kept out of the loop above (a branch per element would end its instruction-level parallelism): the threads that
              // own columns >= 477 clear their ReLU bits.  (Chunk 7 is always read by lin4.)
              if (cq == 3) {                           // columns 480..511: written by the point's eight threads (skip_cols)
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
            if (!may_live && m_ != 0u) viol = true;     // the sparse plan counts on these columns being zero
            // Warps cq = 0,1 own the chunks of k-step 2*nh, warps cq = 2,3 those of k-step 2*nh + 1.  For output half 0 the
            // second pair waits for the first: the next op's first group needs k-step 0 only, and two warps per scheduler
            // deliver it in half the time four would need for both steps.
            publish(2 * nh + (cq >> 1));
            if (nh == 0 && cq < 2) asm volatile("bar.arrive 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
          } else if (opx == 7) {
            // ---------------- lin7 epilogue + the lin8 dot product (deep_sdf_decoder.py:107-108).  With the gradient requested
            // the A operand of B7 is written right here: d7 = w8 * relu'(h7) WITHOUT the tanh' factor (1 - sdf^2), which is a
            // per-row scalar that needs the whole dot product; it multiplies the finished gradient instead (see B4 / B0 below).
            const float* bias = P.bias + opx * HM_HIDDEN + col0;
            const float* w8 = P.w8 + col0;
            const float2 uu = make_float2(unscale_x, unscale_x);
            if (kJac && nh == 0 && cq >= 2) asm volatile("bar.sync 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 w0 = *reinterpret_cast<const float4*>(w8 + 8 * u), w1 = *reinterpret_cast<const float4*>(w8 + 8 * u + 4);
              const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              float r[8];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 bz = *reinterpret_cast<const float2*>(bias + 8 * u + 2 * i);
                const float2 y = fma2(acc[nh][4 * u + i], uu, bz);
                m_ |= ((y.x > 0.f) ? 1u : 0u) << (8 * u + 2 * i) | ((y.y > 0.f) ? 1u : 0u) << (8 * u + 2 * i + 1);
                dot = fmaf(fmaxf(y.x, 0.f), w[2 * i], dot);
                dot = fmaf(fmaxf(y.y, 0.f), w[2 * i + 1], dot);
                r[2 * i] = (y.x > 0.f) ? w[2 * i] * s_next_x : 0.f;
                r[2 * i + 1] = (y.y > 0.f) ? w[2 * i + 1] * s_next_x : 0.f;
              }
              if (kJac && emit) emit_unit(nh, u, r);
            }
            if (!may_live && m_ != 0u) viol = true;
            if (kJac) {
              publish(2 * nh + (cq >> 1));
              if (nh == 0 && cq < 2) asm volatile("bar.arrive 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
            }
          } else if (o.is_last && opx < 15) {
            // ---------------- B4 as the last op of the gradient pass (lin3 is dead in the plan: what B3..B0 would add is exactly
            // zero): the skip-gradient columns 477..511 of d(lin4 input) ARE the input gradient.  The tanh' factor c7 closes the chain.
            if (nh == 1 && hq == 1 && cq >= 2 && ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float a = (j & 1) ? acc[nh][j >> 1].y : acc[nh][j >> 1].x;
                if (col0 + j >= HM_SKIP_COL) P.jac[grow * HM_IN + (col0 + j - HM_SKIP_COL)] = (a * unscale_x) * c7;
              }
            }
          } else if (opx < 15) {
            // ---------------- backward through lin_l (l = 15 - opx = 7..1): d_{l-1} = (d_l W_l) * relu'(h_{l-1})
            if (nh == 0 && cq >= 2) asm volatile("bar.sync 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
            if (emit) {
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
          if (e_w == 0 && lane == 0) HM_TRACE(1 + rank, 14, opx, nh);
        };
        // ---- Schedule of one op (groups in the issue order of group_of(); P = promote, F(nh) = finalize an output half):
        //        P(0,0) P(0,1) P(1,0) P(1,1) P(2,0) P(3,0) F(0) P(2,1) P(3,1) F(1)
        //      Output half 0 is complete two groups before the op ends and is turned into the next op's A chunks 0..3 while the
        //      tensor core still works on half 1; F(1) runs under the next op's first four groups (they read chunks 0..3 only) --
        //      with four TMEM buffers the tensor core can be that far ahead of the promotions.  Groups the plan dropped (gm) are
        //      skipped; a half whose opening group is among them starts from zero (0 + x = x: the same bits as an overwrite).
        if (!(gm & 1u)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[0][i] = make_float2(0.f, 0.f);
        }
        if (!narrow && !(gm & 2u)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[1][i] = make_float2(0.f, 0.f);
        }
        if (gm & 1u) promote(I0, I1);
        if (narrow) {
          if (gm & 2u) promote(I0, I0);
          if (gm & 4u) promote(I0, I0);
          if (gm & 8u) promote(I0, I0);
        } else {
          if (gm & 2u) promote(I1, I1);
          if (wide) {
            if (gm & 4u) promote(I0, I0);
            if (gm & 8u) promote(I1, I0);
            if (gm & 16u) promote(I0, I0);
            if (gm & 32u) promote(I0, I0);
          }
          finalize(I0, op, k_mul, unscale, s_next, m0);
          if (wide) {
            if (gm & 64u) promote(I1, I0);
            if (gm & 128u) promote(I1, I0);
          }
          finalize(I1, op, k_mul, unscale, s_next, m1);
          // ReLU bits of this layer.  Forward + gradient pass: only the layers the plan's gradient ops read (a plan that ends the
          // gradient pass at lin4 reads three of the seven: tile 116 k -> 108 k cycles together with the two-half promotion), and into
          // the plan's free A chunk when it has one -- the thread that writes them reads them back.  (Measured on one box: the
          // shared-memory home and the global scratch take the same time; the store itself is not what a layer costs.)
          if (kMode == 1 && op < 7 && ((P.plan.mask_layers >> op) & 1)) {
            if (P.plan.mask_chunk >= 0) *reinterpret_cast<uint2*>(mask_smem + 4096u * (uint32_t)__popc(P.plan.mask_layers & ((1u << op) - 1u))) = make_uint2(m0, m1);
            else *reinterpret_cast<uint2*>(mask_wr + (size_t)op * kMaskStride) = make_uint2(m0, m1);
          }
          if (kMode == 0 && P.mask_out && op < 8) *reinterpret_cast<uint2*>(mask_wr + (size_t)op * kMaskStride) = make_uint2(m0, m1);
        }
        // ---- lin8 + tanh (deep_sdf_decoder.py:107-108): every thread sums the 8 column-group partials of its point.  A forward-only
        //      pass does it at the end of F7 (= the end of the tile).  With the gradient requested it is DEFERRED to the end of B7's
        //      epilogue: nothing needs the SDF or tanh' before the last gradient op, and here it would sit between F7's last finalize
        //      and B7's first promotion, i.e. on the critical path of the op chain; there it runs while the tensor core works on B6.
        if ((kMode == 0 && op == 7) || (kMode == 1 && op == 8)) {
          // (no barrier in front of the writes: a warp reads the previous tile's sums before it leaves that tile -- before the
          // tile-end barrier of a sparse pass, or, without one, before it publishes any A operand of lin0's output -- and no warp
          // gets to this point of the next tile earlier: lin1 multiplies every chunk the plan keeps of h0, i.e. waits for all warps)
          dot_scratch[p * 8 + g8] = dot;
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          const float4 d0 = *reinterpret_cast<const float4*>(dot_scratch + p * 8), d1 = *reinterpret_cast<const float4*>(dot_scratch + p * 8 + 4);
          f_out = tanhf((((d0.x + d0.y) + (d0.z + d0.w)) + ((d1.x + d1.y) + (d1.z + d1.w))) + P.b8);
          dot = 0.f;
          c7 = 1.f - f_out * f_out;                                             // tanh' (deep_sdf_decoder.py:107-108)
          if (g8 == 0 && ok) P.sdf[P.out_index ? (int64_t)__ldg(P.out_index + grow) : grow] = f_out;
        }
        if (op == 15) {
          // ---------------- B0: g = c7 * (d0 W0 + the parked skip gradient) (35 valid of 64 columns: TMEM columns 0..31 of lanes
          //                  0..63 hold columns 0..31, of lanes 64..127 columns 32..63)
          if (cq == 0 && ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = 32 * hq + j;
              if (col < HM_IN) {
                const float a = (j & 1) ? acc[0][j >> 1].y : acc[0][j >> 1].x;
                float* q = P.jac + grow * HM_IN + col;
                *q = fmaf(a, unscale, __ldcg(q)) * c7;
              }
            }
          }
        }
      }
      // ---- per-tile bookkeeping: saturation (counted per tile and thread, attributed to the latent-table row = fruit) and the
      // sparse plan's verdict: a tile that contradicted an assumption is queued for the full plan (its outputs are rewritten there)
      if (sat | (int)((sat2 & 0xffffu) >= 0x7bffu) | (int)((sat2 >> 16) >= 0x7bffu)) {
        atomicAdd(P.flags + HM_TC_FLAG_SAT, 1);
        if (P.latent_sat && ok) P.latent_sat[cur_li] = 1;
        sat = 0;
        sat2 = 0u;
      }
      sat = sat_next;
      sat_next = 0;
      if (!kRedo && P.plan.sparse) {
        if (__any_sync(0xffffffffu, viol) && lane == 0) ctl[kCtlViolated / 4] = 1u;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (e_w == 0 && lane == 0 && ctl[kCtlViolated / 4]) {
          ctl[kCtlViolated / 4] = 0u;
          if (tile >= 0) P.redo[4 + atomicAdd(P.redo, 1)] = (int32_t)tile;
        }
      }
      cur_li = nxt_li;
    }
#ifdef HM_TC_COUNTERS
    if (P.flags && e_w == 0 && lane == 0) {
      atomicAdd((unsigned long long*)(P.flags + HM_TC_FLAG_DEBUG + 10 + 8 * rank), (unsigned long long)t_pfull);
      atomicAdd((unsigned long long*)(P.flags + HM_TC_FLAG_DEBUG + 12 + 8 * rank), (unsigned long long)t_pbody);
      atomicAdd((unsigned long long*)(P.flags + HM_TC_FLAG_DEBUG + 14 + 8 * rank), (unsigned long long)t_fin);
      atomicAdd((unsigned long long*)(P.flags + HM_TC_FLAG_DEBUG + 16 + 8 * rank), (unsigned long long)(clock64() - t_epi_begin));
    }
#endif
  }
#ifdef HM_TESTING
  // per-CTA busy time of the launch (role loops only), parked behind the peer's timeline region: how evenly the statically
  // assigned tiles finish across the grid (scripts/probe_decoder.py trace prints min / mean / max)
  if (P.trace != nullptr && !kRedo && threadIdx.x == kCtrlWarps * 32 && blockIdx.x < 1024)
    P.trace[((size_t)2 * kTraceCap + kTraceCap / 2 + blockIdx.x) * 2 + 1] = (uint32_t)(clock64() - t_cta_begin);
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // the peer's TMEM / shared memory stay valid until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side: plans + weight blob
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

// the order in which the 64-wide chunks of a layer are filled with alive units: k-steps pair the chunks {0,2} {1,3} {4,6} {5,7},
// so filling 0, 2, 1, 3, ... keeps whole k-steps (= whole accumulation groups) busy and fills output half 0 before half 1
constexpr int kChunkFill[8] = {0, 2, 1, 3, 4, 6, 5, 7};

uint8_t groups_from_masks(const hm_tc_op& o) {
  uint8_t gm = 0;
  for (int g = 0; g < groups_of(o.n_kchunks, o.n_nblocks); ++g) {
    int step, nh;
    group_of(o.n_kchunks, o.n_nblocks, g, step, nh);
    const int c0 = (o.n_kchunks == 1) ? 0 : chunk_of(step, 0), c1 = (o.n_kchunks == 1) ? 0 : chunk_of(step, 1);
    if (((o.half_mask >> nh) & 1) && (((o.chunk_mask >> c0) | (o.chunk_mask >> c1)) & 1)) gm |= (uint8_t)(1u << g);
  }
  return gm;
}

// amask[l] = chunks of h_l that may hold non-zeros (l = 0..7).  Fills the masks of `plan` (geometry / scales / offsets are set
// by the caller).
void fill_plan_masks(hm_tc_plan& plan, const uint8_t (&amask)[8]) {
  const bool cut = amask[3] == 0;                  // lin3 dead: the gradient pass ends at lin4's skip columns
  for (int op = 0; op < HM_TC_NOPS_ALL; ++op) {
    hm_tc_op& o = plan.ops[op];
    o.is_last = 0;
    o.pad_[0] = o.pad_[1] = 0;
    if (op < 8) {
      o.chunk_mask = (op == 0) ? 0x01 : (op == 4) ? (uint8_t)(amask[3] | 0x80) : amask[op - 1];
      o.half_mask = 3;
      o.verify_alive = amask[op];
    } else {
      const int L = 15 - op;                        // backward through lin_L: A = d_L, output = d_{L-1} (L = 0: the input gradient)
      const uint8_t out_cols = (L == 0) ? 0xFF : (L == 4) ? (uint8_t)(amask[3] | 0x80) : amask[L - 1];     // (B0: 35 columns, all read)
      o.chunk_mask = (cut && L <= 3) ? 0 : amask[L];
      o.half_mask = (L == 0) ? 1 : (uint8_t)(((out_cols & 0x0F) ? 1 : 0) | ((out_cols & 0xF0) ? 2 : 0));
      o.verify_alive = 0xFF;
      o.need_out = out_cols;
      if ((cut && L == 4) || (!cut && L == 0)) o.is_last = 1;
    }
    o.group_mask = groups_from_masks(o);
  }
  for (int op = 0; op < 8; ++op) plan.ops[op].need_out = plan.ops[op + 1].chunk_mask;      // (op 7 feeds B7; ignored by forward-only passes)
  plan.last_op_fwd = 7;
  plan.last_op_jac = cut ? 11 : 15;
  // F0's operand goes to a chunk nobody reads while the LAST op of a tile runs (either kind of pass), if there is one: the next
  // tile's operand is then written during that op and F0's MMAs are issued right behind it (kernel: X0_READY)
  {
    const uint8_t busy = (uint8_t)(plan.ops[7].chunk_mask | plan.ops[plan.last_op_jac].chunk_mask);
    plan.x0_chunk = 0;
    plan.x0_early = 0;
    const int pref[8] = {5, 6, 4, 3, 7, 2, 1, 0};
    if (!getenv("HM_TC_NO_X0_EARLY"))
      for (int c : pref)
        if (!((busy >> c) & 1u)) { plan.x0_chunk = c; plan.x0_early = 1; break; }
  }
  plan.sparse = 0;
  for (int l = 0; l < 8; ++l) plan.sparse |= (amask[l] != 0xFF);
  // ReLU bits the gradient pass reads, and a shared-memory home for them (kernel: mask_smem)
  plan.mask_layers = 0;
  for (int l = 0; l < 7; ++l) {
    const hm_tc_op& b = plan.ops[14 - l];            // B_{l+1}: d_l = (d_{l+1} W_{l+1}) * relu'(h_l)
    if (b.group_mask != 0 && !b.is_last) plan.mask_layers |= 1 << l;
  }
  plan.mask_chunk = -1;
  if (plan.mask_layers != 0 && __builtin_popcount(plan.mask_layers) <= 4 && !getenv("HM_TC_NO_SMEM_MASKS")) {
    const int first_l = __builtin_ctz(plan.mask_layers);
    uint8_t touched = (uint8_t)(1u << plan.x0_chunk);
    for (int op = first_l; op <= 14 - first_l; ++op) touched |= (uint8_t)(plan.ops[op].chunk_mask | (plan.ops[op].group_mask ? plan.ops[op].need_out : 0));
    for (int c = 6; c >= 0; --c)
      if (!((touched >> c) & 1u)) { plan.mask_chunk = c; break; }
  }
  // ---- stage program (common.cuh): one record per issued stage, in the order the blob stores them
  int n = 0;
  plan.n_rec_fwd = 0;
  for (int op = 0; op < HM_TC_NOPS_ALL; ++op) {
    if (op == HM_TC_NOPS_FWD) plan.n_rec_fwd = n;
    const hm_tc_op& o = plan.ops[op];
    if (o.group_mask == 0) continue;
    const int ng = groups_of(o.n_kchunks, o.n_nblocks), nwhich = (o.n_kchunks == 1) ? 1 : 2;
    const int op_begin = n;
    for (int g = 0; g < ng; ++g) {
      if (!((o.group_mask >> g) & 1u)) continue;
      int step, nh;
      group_of(o.n_kchunks, o.n_nblocks, g, step, nh);
      const int group_begin = n;
      for (int part = 0; part < 2; ++part)
        for (int which = 0; which < nwhich; ++which) {
          int chunk = (o.n_kchunks == 1) ? 0 : chunk_of(step, which);
          if (!((o.chunk_mask >> chunk) & 1u)) continue;
          if (op == 0) chunk = plan.x0_chunk;                              // where the A operand of F0 lives (fill_plan_masks)
          const int sidx = (g * 2 + part) * nwhich + which;               // stage index inside the op (blob order)
          const int64_t off = o.blob_offset + (int64_t)sidx * o.stage_rows * 128;
          uint32_t r = (uint32_t)chunk | (part ? HM_TC_REC_PART : 0u) | (o.stage_rows == 64 ? HM_TC_REC_NARROW : 0u) | ((uint32_t)op << 12) |
                       ((uint32_t)(off / 8192) << 16) | ((uint32_t)g << 28);
          if (off % 8192 != 0 || off / 8192 > 0xFFF || n >= HM_TC_MAX_RECS) { fprintf(stderr, "hm_tc: stage program overflow\n"); abort(); }
          plan.rec[n++] = r;
        }
      // the group opener waits for the A operand's k-steps up to its own (F0 reads chunk 0 only: all four phases up front)
      const uint32_t need_ready = op == 0 ? (g == 0 ? 7u : 0u) : (uint32_t)step + 1u;      // F0: X0_READY (waited once, at the op's first stage)
      plan.rec[group_begin] |= HM_TC_REC_GROUP_FIRST | (need_ready << 6);
      plan.rec[n - 1] |= HM_TC_REC_GROUP_LAST;
    }
    plan.rec[op_begin] |= HM_TC_REC_OP_FIRST;
    plan.rec[n - 1] |= HM_TC_REC_OP_LAST;
  }
  plan.n_rec_all = n;
  for (int i = n; i < HM_TC_MAX_RECS; ++i) plan.rec[i] = 0;
}

// geometry of the 16 ops and where each op's stages start in the weight blob (all stages of all groups, lo tiles and hi tiles,
// whether a plan issues them or not); returns the blob size
int64_t plan_geometry(hm_tc_plan& plan) {
  int64_t bytes = 0;
  for (int op = 0; op < HM_TC_NOPS_ALL; ++op) {
    hm_tc_op& o = plan.ops[op];
    o.n_kchunks = (op == 0) ? 1 : 8;
    o.n_nblocks = (op == 15) ? 1 : 2;                // 256-column output halves (B0: one 64-column block)
    o.stage_rows = (op == 15) ? 64 : 256;
    o.blob_offset = bytes;
    bytes += (int64_t)groups_of(o.n_kchunks, o.n_nblocks) * 2 * (o.n_kchunks == 1 ? 1 : 2) * o.stage_rows * 128;
  }
  return bytes;
}

}  // namespace

#ifdef HM_TESTING
// TEST-ONLY, host only (no context, no GPU): the plan the engine would build for the given per-layer alive-chunk masks --
// masks, stage program, chunk assignments -- so that tests/test_host_logic.py can re-derive it independently.
// out_ops[op] = {n_kchunks, n_nblocks, stage_rows / 64, chunk_mask, half_mask, group_mask, need_out, verify_alive, is_last};
// out_info = {n_rec_fwd, n_rec_all, last_op_fwd, last_op_jac, sparse, x0_chunk, x0_early, mask_layers, mask_chunk}.
extern "C" int hm_debug_tc_plan(const uint8_t* amask8, int32_t* out_info, uint32_t* out_rec, uint8_t* out_ops, int64_t* out_blob_offset) {
  hm_tc_plan plan;
  memset(&plan, 0, sizeof(plan));
  plan_geometry(plan);
  uint8_t am[8];
  memcpy(am, amask8, 8);
  fill_plan_masks(plan, am);
  const int32_t info[9] = {plan.n_rec_fwd, plan.n_rec_all, plan.last_op_fwd, plan.last_op_jac, plan.sparse, plan.x0_chunk, plan.x0_early,
                           plan.mask_layers, plan.mask_chunk};
  memcpy(out_info, info, sizeof(info));
  memcpy(out_rec, plan.rec, sizeof(plan.rec));
  for (int op = 0; op < HM_TC_NOPS_ALL; ++op) {
    const hm_tc_op& o = plan.ops[op];
    const uint8_t v[9] = {(uint8_t)o.n_kchunks, (uint8_t)o.n_nblocks, (uint8_t)(o.stage_rows / 64), o.chunk_mask, o.half_mask, o.group_mask, o.need_out,
                          o.verify_alive, o.is_last};
    memcpy(out_ops + 9 * op, v, 9);
    out_blob_offset[op] = o.blob_offset;
  }
  return HM_OK;
}
#endif

int hm_tc_init(hm_context* ctx) {
  // ---- which hidden units were alive in the calibration pass -> unit permutation + chunk masks
  // perm[l][new] = old index of the unit that sits at position `new` of layer l in the tensor-core engine's order
  std::vector<int> perm[8];
  uint8_t amask[8], afull[8];
  const bool have_cal = ctx->unit_max.size() == (size_t)8 * HM_HIDDEN && !getenv("HM_TC_NO_SPARSE");
  for (int l = 0; l < 8; ++l) {
    perm[l].resize(HM_HIDDEN);
    for (int u = 0; u < HM_HIDDEN; ++u) perm[l][u] = u;
    amask[l] = afull[l] = 0xFF;
    if (!have_cal) continue;
    const float* um = ctx->unit_max.data() + (size_t)l * HM_HIDDEN;
    if (l == 3) {                                  // lin3's columns keep their places (477 outputs + the skip-concat columns)
      int n_alive = 0;
      for (int u = 0; u < HM_SKIP_COL; ++u) n_alive += um[u] > 0.f;
      amask[3] = n_alive ? 0xFF : 0x00;
      continue;
    }
    std::vector<int> alive, dead;
    for (int u = 0; u < HM_HIDDEN; ++u) (um[u] > 0.f ? alive : dead).push_back(u);
    const int c = std::max(1, (int)(alive.size() + 63) / 64);
    uint8_t m = 0;
    for (int i = 0; i < c && i < 8; ++i) m |= (uint8_t)(1u << kChunkFill[i]);
    amask[l] = m;
    std::vector<int> order = alive;
    order.insert(order.end(), dead.begin(), dead.end());
    for (int k = 0; k < HM_HIDDEN; ++k) perm[l][kChunkFill[k / 64] * 64 + k % 64] = order[k];
  }
  // permuted weights: Wp[l][r][c] = W[l][perm_l[r]][perm_{l-1}[c]] (lin0's inputs, lin4's inputs = [h3 | x0] and lin3's outputs
  // keep their order); lin8: columns by perm_7
  std::vector<float> Wp[HM_LAYERS], bp[HM_LAYERS];
  for (int l = 0; l < HM_LAYERS; ++l) {
    const int od = ctx->out_dim[l], id = ctx->in_dim[l];
    Wp[l].assign((size_t)od * id, 0.f);
    bp[l].assign(od, 0.f);
    const bool perm_rows = l < 8 && l != 3, perm_cols = l >= 1 && l != 4;
    for (int r = 0; r < od; ++r) {
      const int ro = perm_rows ? perm[l][r] : r;
      bp[l][r] = ctx->h_b[l][ro];
      for (int c = 0; c < id; ++c) Wp[l][(size_t)r * id + c] = ctx->h_W[l][(size_t)ro * id + (perm_cols ? perm[l - 1][c] : c)];
    }
  }
  // op list: F0..F7 (lin0..lin7), B7..B1, B0
  hm_tc_plan& plan = ctx->tc_plan;
  const int64_t blob_total = plan_geometry(plan);
  std::vector<uint8_t> blob;
  blob.reserve((size_t)blob_total);
  auto wmax = [&](int l) {
    float m = 0.f;
    for (float v : Wp[l]) m = std::max(m, std::fabs(v));
    return std::max(m, 1e-20f);
  };
  for (int op = 0; op < HM_TC_NOPS_ALL; ++op) {
    const bool fwd = op < 8;
    const int l = fwd ? op : 15 - op;
    hm_tc_op& o = plan.ops[op];                      // (geometry and blob offsets: plan_geometry)
    const float amax = std::max(ctx->act_absmax[op], 1e-20f);
    o.in_scale = pow2_floor(1024.f / amax);          // 64x headroom below the fp16 maximum
    const float w_scale = pow2_floor(8192.f / wmax(l));
    o.out_unscale = 1.f / (o.in_scale * w_scale);
    HM_CHECK(o.blob_offset == (int64_t)blob.size(), "tensor-core blob layout out of step with plan_geometry");
    const std::vector<float>& W = Wp[l];
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
  ctx->tc_plan_full = plan;
  fill_plan_masks(ctx->tc_plan, amask);
  fill_plan_masks(ctx->tc_plan_full, afull);
  if (ctx->d_tc_blob) HM_CUDA(cudaDeviceSynchronize());      // re-calibration: no kernel on any stream may still read the old blob / biases
  if (ctx->d_tc_blob && ctx->tc_blob_bytes != blob.size()) { cudaFree(ctx->d_tc_blob); ctx->d_tc_blob = nullptr; }
  if (!ctx->d_tc_blob) HM_CUDA(cudaMalloc(&ctx->d_tc_blob, blob.size()));
  ctx->tc_blob_bytes = blob.size();
  HM_CUDA(cudaMemcpy(ctx->d_tc_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  if (!ctx->d_tc_masks) {
    HM_CUDA(cudaMalloc(&ctx->d_tc_masks, sizeof(uint32_t) * (size_t)ctx->sm_count * 8 * kMaskWordsPerOp));
    HM_CUDA(cudaMalloc(&ctx->d_tc_flags, sizeof(int32_t) * HM_TC_FLAG_COUNT));
    HM_CUDA(cudaMemset(ctx->d_tc_flags, 0, sizeof(int32_t) * HM_TC_FLAG_COUNT));
    HM_CUDA(cudaFuncSetAttribute(tc_decoder_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    HM_CUDA(cudaFuncSetAttribute(tc_decoder_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    HM_CUDA(cudaFuncSetAttribute(tc_decoder_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    HM_CUDA(cudaFuncSetAttribute(tc_decoder_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    HM_CUDA(cudaFuncSetAttribute(tc_decoder_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    HM_CUDA(cudaFuncSetAttribute(tc_decoder_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
  }
  std::vector<float> bias(8 * HM_HIDDEN, 0.f);
  for (int l = 0; l < 8; ++l) {
    memcpy(bias.data() + l * HM_HIDDEN, bp[l].data(), sizeof(float) * bp[l].size());
    if (l < 7)                                       // the F_l epilogue emits h_l * in_scale(F_{l+1}): fold the scale into the bias
      for (int c = 0; c < HM_HIDDEN; ++c) bias[l * HM_HIDDEN + c] *= plan.ops[l + 1].in_scale;
  }
  ctx->h_tc_bias = bias;
  memcpy(ctx->h_w8p, Wp[8].data(), sizeof(float) * HM_HIDDEN);
  return HM_OK;
}

// Tensor-core FLOP issued per row by a plan: every issued stage pair is 4 (A_hi x W_lo) + 8 (A_lo x W_hi, A_hi x W_hi) MMAs of
// M = 128 rows (one tile pair) x N = stage_rows x K = 16.
static double plan_flop_per_row(const hm_tc_plan& plan, bool jac) {
  double f = 0;
  const int n_ops = jac ? HM_TC_NOPS_ALL : HM_TC_NOPS_FWD;
  for (int op = 0; op < n_ops; ++op) {
    const hm_tc_op& o = plan.ops[op];
    for (int g = 0; g < groups_of(o.n_kchunks, o.n_nblocks); ++g) {
      if (!((o.group_mask >> g) & 1u)) continue;
      int step, nh;
      group_of(o.n_kchunks, o.n_nblocks, g, step, nh);
      for (int which = 0; which < (o.n_kchunks == 1 ? 1 : 2); ++which) {
        const int chunk = (o.n_kchunks == 1) ? 0 : chunk_of(step, which);
        if ((o.chunk_mask >> chunk) & 1u) f += 12.0 * 2.0 * o.stage_rows * 16;
      }
    }
  }
  return f;
}

void hm_tc_plan_info(const hm_context* ctx, double* out) {
  out[0] = plan_flop_per_row(ctx->tc_plan, false);
  out[1] = plan_flop_per_row(ctx->tc_plan, true);
  out[2] = plan_flop_per_row(ctx->tc_plan_full, false);
  out[3] = plan_flop_per_row(ctx->tc_plan_full, true);
  for (int l = 0; l < 8; ++l) out[4 + l] = __builtin_popcount(ctx->tc_plan.ops[l].verify_alive);
}

void hm_tc_free(hm_context* ctx) {
  if (ctx->d_tc_blob) cudaFree(ctx->d_tc_blob);
  if (ctx->d_tc_masks) cudaFree(ctx->d_tc_masks);
  if (ctx->d_tc_flags) cudaFree(ctx->d_tc_flags);
  if (ctx->d_tc_trace) cudaFree(ctx->d_tc_trace);
  if (ctx->d_tc_redo) cudaFree(ctx->d_tc_redo);
  ctx->d_tc_trace = nullptr;
  ctx->d_tc_blob = nullptr;
  ctx->d_tc_masks = nullptr;
  ctx->d_tc_flags = nullptr;
  ctx->d_tc_redo = nullptr;
  ctx->tc_redo_cap = 0;
}

int hm_tc_decode(hm_context* ctx, const hm_rows& rows, float* d_sdf, float* d_jac, cudaStream_t st) {
  HM_CHECK(ctx->d_tc_blob, "tensor-core engine not initialised");
  TcParams P;
  P.plan = ctx->sparse_plan ? ctx->tc_plan : ctx->tc_plan_full;
  P.blob = ctx->d_tc_blob;
  memcpy(P.bias, ctx->h_tc_bias.data(), sizeof(P.bias));
  memcpy(P.w8, ctx->h_w8p, sizeof(P.w8));
  P.b8 = ctx->h_b[8][0];
  P.rows = rows.d_rows;
  P.xyz = rows.d_xyz;
  P.latents = rows.d_latents;
  P.row_latent = rows.d_row_latent;
  P.n_dynamic = rows.d_n_dynamic;
  P.n = rows.n;
  P.sdf = d_sdf;
  P.out_index = rows.d_out_index;
  P.jac = d_jac;
  P.masks = reinterpret_cast<uint32_t*>(ctx->d_tc_masks);
  P.flags = ctx->d_tc_flags;
  P.latent_sat = rows.d_latent_sat;
  P.trace = ctx->d_tc_trace;
  P.grid_n = rows.grid_n;
  P.grid_voxel = rows.grid_voxel;
  P.grid_radius = rows.grid_radius;
  P.redo = nullptr;
  P.mask_out = d_jac ? nullptr : rows.d_mask_out;
  P.mask_in = rows.d_mask_in;
  P.src_row = rows.d_src_row;
  const int mode = rows.d_mask_in ? 2 : d_jac ? 1 : 0;
  HM_CHECK(mode != 2 || (d_jac && d_sdf && rows.d_src_row), "gradient-only decode needs the Jacobian output, the SDF input and the source-row table");
  const int64_t n_tiles = (rows.n + HM_TC_TILE_M - 1) / HM_TC_TILE_M;
  const int csize = 2;                                              // the kernel is a CTA-pair kernel
  const int64_t n_units = (n_tiles + csize - 1) / csize;             // a unit = one 64-row tile per CTA of the cluster
  const int grid = (int)std::min<int64_t>(n_units, ctx->sm_count / csize) * csize;
  const bool sparse = P.plan.sparse != 0;
  if (sparse) {                                                     // queue of the tiles that contradict the sparse plan
    const size_t need = (size_t)n_tiles + 8;
    if (need > ctx->tc_redo_cap) {
      if (ctx->d_tc_redo) { HM_CUDA(cudaDeviceSynchronize()); cudaFree(ctx->d_tc_redo); ctx->d_tc_redo = nullptr; }
      const size_t cap = std::max(need, (size_t)1 << 16);
      HM_CUDA(cudaMalloc(&ctx->d_tc_redo, sizeof(int32_t) * cap));
      ctx->tc_redo_cap = cap;
    }
    HM_CUDA(cudaMemsetAsync(ctx->d_tc_redo, 0, 16, st));
    P.redo = ctx->d_tc_redo;
  }
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
  if (mode == 2) HM_CUDA(cudaLaunchKernelEx(&cfg, tc_decoder_kernel<2, false>, P));
  else if (mode == 1) HM_CUDA(cudaLaunchKernelEx(&cfg, tc_decoder_kernel<1, false>, P));
  else HM_CUDA(cudaLaunchKernelEx(&cfg, tc_decoder_kernel<0, false>, P));
  ctx->counters.kernel_launches += 1;
  if (sparse) {
    // second pass: the queued tiles with the full plan (the kernel returns at once when the queue is empty)
    P.plan = ctx->tc_plan_full;
    if (mode == 2) HM_CUDA(cudaLaunchKernelEx(&cfg, tc_decoder_kernel<2, true>, P));
    else if (mode == 1) HM_CUDA(cudaLaunchKernelEx(&cfg, tc_decoder_kernel<1, true>, P));
    else HM_CUDA(cudaLaunchKernelEx(&cfg, tc_decoder_kernel<0, true>, P));
    ctx->counters.kernel_launches += 1;
  }
  HM_CUDA(cudaGetLastError());
  return HM_OK;
}
