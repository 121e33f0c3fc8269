// Shared declarations of the hortimapping_b200 CUDA library (internal; the public ABI is
// include/hortimapping_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/hortimapping_b200.h"

#define HM_SKIP_COL 477   // lin3 output width = 512 - 35 (deep_sdf_decoder.py:41-42): columns 477..511 of
                          // the layer-4 input carry the raw [latent, xyz] vector (deep_sdf_decoder.py:87-88)

void hm_set_error(const char* fmt, ...);

#define HM_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      hm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return HM_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define HM_CHECK(cond, ...)          \
  do {                               \
    if (!(cond)) {                   \
      hm_set_error(__VA_ARGS__);     \
      return HM_ERR_INVALID;         \
    }                                \
  } while (0)

// ---------------------------------------------------------------------------------------------
// tensor-core engine: operator list of one tile pass.  An "op" is one fused layer GEMM
//   D[64 x N] = A[64 x K] * W_op^T  (fp16 hi/lo split operands stacked as 128 MMA rows, fp32 accumulate in TMEM).
// Forward ops F0..F7 are lin0..lin7; backward ops B7..B0 multiply by the transposed weights.
// ---------------------------------------------------------------------------------------------
#define HM_TC_NOPS_FWD 8
#define HM_TC_NOPS_ALL 16
#define HM_TC_TILE_M 64          // rows (points) per tile
#define HM_TC_STAGE_N 256        // weight rows (output features) per pipeline stage (64 for B0); a CTA pair splits them
#define HM_TC_CHUNK_K 64         // K extent of one 128-byte swizzle atom row (fp16)

// int32 slots of hm_context::d_tc_flags (device counters of the tensor-core engine; 64-bit totals take two slots)
#define HM_TC_FLAG_SAT 0            // saturation events: (tile, thread) pairs in which an fp16 operand conversion saturated
#define HM_TC_FLAG_ROWS_FWD 2       // rows evaluated by forward-only launches
#define HM_TC_FLAG_ROWS_JAC 4       // rows evaluated by forward + input-gradient launches
#define HM_TC_FLAG_TILES_FWD 6      // 64-row tiles processed (forward-only / forward + gradient)
#define HM_TC_FLAG_TILES_JAC 8
#define HM_TC_FLAG_DEAD_FWD 10      // ... of which failed the sparse plan's checks and were re-evaluated with the full plan
#define HM_TC_FLAG_DEAD_JAC 12
#define HM_TC_FLAG_ROWS_BWD 14      // rows / tiles / re-evaluated tiles of gradient-only launches (stored ReLU masks)
#define HM_TC_FLAG_TILES_BWD 16
#define HM_TC_FLAG_DEAD_BWD 18
#define HM_TC_FLAG_DEBUG 32         // wait-cycle counters of the instrumented testing build (HM_TC_COUNTERS)
#define HM_TC_FLAG_COUNT 128

struct hm_tc_op {
  int32_t n_kchunks;             // K / 64 of the A operand (1 for F0, 8 otherwise)
  int32_t n_nblocks;             // output halves of stage_rows columns (2; 1 for B0)
  int32_t stage_rows;            // weight rows per stage = MMA N (256; 64 for B0)
  // Sparse plan: hidden units are permuted so that the units calibration saw alive come first, which makes
  // whole 64-wide k-chunks of the activations exactly zero.  The masks below say what is still multiplied / produced; a tile for
  // which a forward op finds a non-zero outside `verify_alive` is re-evaluated with the full plan (all masks 0xFF).
  uint8_t chunk_mask;            // k-chunks of the A operand that are multiplied (bit c = chunk c); 0 = the op is dropped
  uint8_t half_mask;             // output halves that are computed (bit nh)
  uint8_t group_mask;            // = the accumulation groups issued, in issue order (derived from the two masks above)
  uint8_t need_out;              // output chunks the epilogue has to produce (what the next executed op multiplies / the final result)
  uint8_t verify_alive;          // forward ops: output chunks that may hold non-zeros after the ReLU (0xFF = no assumption)
  uint8_t is_last;               // backward ops: the last executed op of the gradient pass writes the final Jacobian rows
  uint8_t pad_[2];
  float in_scale;                // power of two applied to the A operand before the fp16 split
  float out_unscale;             // 1 / (in_scale * w_scale): turns the accumulator back into fp32 units
  int64_t blob_offset;           // byte offset of this op's first stage in the weight blob
};

// Stage program: the weight stages a tile consumes, flattened on the host into one 32-bit record per ISSUED stage in consumption
// order (ops F0..F7 first, then B7..B0), so that the MMA issuer, the weight producer and the peer's arrival forwarder walk a flat
// list instead of re-deriving the plan's nested op / group / part / chunk loops and masks for every stage: the issuing warp's own
// instruction stream sits on the tile's critical path (measured with the timeline trace, profiles/r02e_trace_*.txt).
#define HM_TC_REC_CHUNK(r) ((r) & 7u)               // A operand k-chunk of the stage
#define HM_TC_REC_PART 0x8u                         // 0: lo weight tile (x A_hi), 1: hi weight tile (x A_lo, x A_hi)
#define HM_TC_REC_GROUP_FIRST 0x10u                 // opens an accumulation group: fresh TMEM buffer, first MMA overwrites
#define HM_TC_REC_GROUP_LAST 0x20u                  // closes it: the partial accumulator is committed to the epilogue warps
#define HM_TC_REC_NEED_READY(r) (((r) >> 6) & 7u)   // group opener: A_READY k-step phases that must have been consumed (0..4); 7 = F0: X0_READY
#define HM_TC_REC_OP_FIRST 0x200u                   // first stage of an op (one A_READY phase per executed op)
#define HM_TC_REC_OP_LAST 0x400u                    // last stage of an op: the op's remaining A_READY phases are consumed
#define HM_TC_REC_NARROW 0x800u                     // B0: 64 weight rows per stage (MMA N = 64) instead of 256
#define HM_TC_REC_OP(r) (((r) >> 12) & 15u)         // op index (timeline trace)
#define HM_TC_REC_SRC(r) (((r) >> 16) & 0xFFFu)     // blob offset of the stage (both CTAs' halves) in units of 8 KB
#define HM_TC_REC_GROUP(r) ((r) >> 28)              // group index within the op (timeline trace)
#define HM_TC_MAX_RECS 480                          // full plan: 4 + 14 x 32 + 16 = 468

struct hm_tc_plan {
  hm_tc_op ops[HM_TC_NOPS_ALL];
  int32_t last_op_fwd, last_op_jac;     // last executed op of a forward-only / forward + gradient pass
  int32_t sparse;                       // 1: some mask is not full, i.e. tiles can fail the checks and need the full plan
  int32_t x0_chunk;                     // A chunk that holds F0's operand [x0 * s | 0 ...] (K padded 35 -> 64)
  int32_t x0_early;                     // 1: that chunk is free while the tile's LAST op runs (forward-only and forward + gradient passes), so the
                                        // next tile's operand is written there and F0's MMAs follow the last op's without a bubble
  int32_t mask_layers;                  // bit l: the gradient pass reads the ReLU bits of h_l (l = 0..6), i.e. B_{l+1} is executed and is not the last op
  int32_t mask_chunk;                   // forward + gradient pass: A chunk that no op touches between the first stored layer and the last reader and
                                        // that holds those bits (<= 4 layers x 4 KB) instead of the global scratch; -1 = none (full plan)
  int32_t n_rec_fwd, n_rec_all;         // stage program: records [0, n_rec_fwd) = forward ops, [n_rec_fwd, n_rec_all) = gradient ops
  uint32_t rec[HM_TC_MAX_RECS];
};

struct hm_context {
  int device = 0;
  int engine = HM_ENGINE_TC;
  int sm_count = 0;
  // fp32 weights (device): W[l] is [out_pad][in] with lin3 zero-padded to 512 rows
  float* d_W[HM_LAYERS] = {};
  float* d_b[HM_LAYERS] = {};
  int in_dim[HM_LAYERS] = {};
  int out_dim[HM_LAYERS] = {};     // padded (lin3 -> 512)
  std::vector<float> h_W[HM_LAYERS];
  std::vector<float> h_b[HM_LAYERS];
  // tensor-core engine
  uint8_t* d_tc_blob = nullptr;    // pre-swizzled fp16 hi/lo weight stages for all 16 ops
  size_t tc_blob_bytes = 0;
  hm_tc_plan tc_plan;              // the sparse plan (== the full plan for a model without dead units)
  hm_tc_plan tc_plan_full;         // every mask full: the reference evaluation, used for the tiles that fail the sparse plan's checks
  int32_t* d_tc_redo = nullptr;    // [0] = number of queued tiles, [4 ..] = their indices (grow-only)
  size_t tc_redo_cap = 0;
  float h_w8p[HM_HIDDEN] = {};     // lin8 weight in the permuted unit order of the tensor-core engine (travels as a kernel parameter)
  float act_absmax[HM_TC_NOPS_ALL] = {};   // calibration result: max |A operand| per op
  std::vector<float> unit_max;             // calibration result: [8][512] largest activation of every hidden unit (0 = never alive)
  std::vector<float> h_tc_bias;    // [8][512] biases of lin0..7 in the engine's unit order, pre-scaled (lin3 padded with 0); kernel parameter
  float* d_w8 = nullptr;           // [512] lin8 weight, d_b8 scalar in d_b[8]
  uint8_t* d_tc_masks = nullptr;   // per-CTA ReLU mask scratch
  int32_t* d_tc_flags = nullptr;   // saturation counter etc.
  uint32_t* d_tc_trace = nullptr;  // timeline buffer of the instrumented testing build (NULL in the product)
  int mask_reuse = 1;              // joint loop: gradient of the in-band samples from the forward pass's ReLU bits (hm_set_mask_reuse)
  int sparse_plan = 1;             // tensor-core engine: use the calibrated sparse plan (hm_set_sparse_plan); 0 = full plan always
  // grow-only workspace
  void* ws = nullptr;
  size_t ws_bytes = 0;
  void* ws2 = nullptr;             // optimiser workspace (separate so decoder calls never alias it)
  size_t ws2_bytes = 0;
  void* mesh_ws = nullptr;         // iso-surface scratch (edge flags / scans) and outputs of the last hm_isosurface call
  size_t mesh_ws_bytes = 0;
  void* mesh_out = nullptr;
  size_t mesh_out_bytes = 0;
  int64_t mesh_n_verts = 0, mesh_n_faces = 0;
  void* io_arena = nullptr;        // device-side copies of the host buffers of the *_host entry points
  size_t io_arena_bytes = 0;
  void* h_stage = nullptr;         // pinned staging of the optimiser's per-call host tables (optimizer.cu stage_reserve)
  size_t stage_bytes = 0;
  cudaEvent_t stage_event = nullptr;
  hm_counters counters = {};
  int profiling = 0;
  std::vector<cudaEvent_t> prof_events;      // (begin, end) pairs awaiting read-back
  std::vector<int> prof_kinds;               // per pair: 0 = forward-only launch, 1 = forward + input gradient
  std::vector<cudaEvent_t> prof_pool;
  // last LM system (test hook)
  float* d_last_H = nullptr;
  float* d_last_b = nullptr;
  float* d_last_dx = nullptr;
  int last_est = 0;
  int last_n_fruits = 0;           // capacity of the three buffers above (fruits)
  int last_call_fruits = 0;        // fruits of the most recent optimise call
  // cross-stream ordering of the calls of this context (hm_stream_scope)
  cudaEvent_t busy_event = nullptr;
  cudaStream_t last_stream = nullptr;
};

// A context owns scratch that every call reuses (decoder workspaces, per-CTA ReLU masks, device counters, the optimiser workspace,
// the last-system buffers), so two calls of the SAME context must not overlap on the GPU even when they are issued on different
// streams.  Every public entry point that touches that scratch opens one of these: if the previous call ran on another stream, the
// new stream first waits (on the GPU, no host synchronisation) for the event recorded when that call was enqueued; on exit the
// event is re-recorded on the current stream.  Calls from different host THREADS still need external locking (header).
struct hm_stream_scope {
  hm_context* ctx;
  cudaStream_t st;
  hm_stream_scope(hm_context* c, cudaStream_t s) : ctx(c), st(s) {
    if (ctx && ctx->busy_event && ctx->last_stream != st) cudaStreamWaitEvent(st, ctx->busy_event, 0);
  }
  ~hm_stream_scope() {
    if (!ctx) return;
    if (!ctx->busy_event && cudaEventCreateWithFlags(&ctx->busy_event, cudaEventDisableTiming) != cudaSuccess) return;
    cudaEventRecord(ctx->busy_event, st);
    ctx->last_stream = st;
  }
};

int hm_ws_reserve(hm_context* ctx, size_t bytes);     // ctx->ws
int hm_ws2_reserve(hm_context* ctx, size_t bytes);    // ctx->ws2

// ---- decoder engines (decoder_simt.cu / decoder_tc.cu) ----
// Rows are described by xyz [n][3] + a latent table [L][32] + optional per-row latent index (NULL = all
// rows use latent 0), or by full rows [n][35] (d_rows != NULL takes precedence).
struct hm_rows {
  const float* d_rows;       // [n][35] or NULL
  const float* d_xyz;        // [n][3]
  const float* d_latents;    // [L][32]
  const int32_t* d_row_latent;  // [n] or NULL
  int64_t n;
  const int32_t* d_n_dynamic;   // optional device-side row count (<= n); NULL = use n
  const int32_t* d_out_index = nullptr;   // optional: the SDF of row i goes to d_sdf[d_out_index[i]] (rows compacted from a larger set)
  int32_t* d_latent_sat = nullptr;        // optional [L]: the tensor-core engine sets entry l when a row of latent-table row l saturated fp16
  // tensor-core engine only.  Forward-only call: store the ReLU bits of all 8 layers per 64-row tile ([tile][8][512][2] words).
  uint32_t* d_mask_out = nullptr;
  // Gradient-only call (d_jac given, d_sdf is an INPUT): the stored bits and, per row, the row index it had in the call that stored them.
  const uint32_t* d_mask_in = nullptr;
  const int32_t* d_src_row = nullptr;
  // fused mesher grid (wild_completion/utils.py:542-562): when grid_n > 0 the xyz of row i is create_voxel_grid(grid_n)[i] *
  // grid_radius, generated inside the decoder kernel (d_xyz is ignored; all rows use latent 0)
  int32_t grid_n = 0;
  float grid_voxel = 0.f;
  float grid_radius = 0.f;
};

// one coordinate (c = 0,1,2) of point i of the reference's voxel grid, in the reference's fp32 operation order (the grid is
// "sheared": LongTensor / int is true division): idx -> float32, / n, fmod n, * voxel_size, + (-1), * cube_radius
__device__ __forceinline__ float hm_grid_coord(int64_t i, int c, int n, float voxel_size, float cube_radius) {
  const float fn = (float)n;
  float v;
  if (c == 2) v = (float)(i % n);
  else {
    const float q1 = __fdiv_rn((float)i, fn);
    v = (c == 1) ? fmodf(q1, fn) : fmodf(__fdiv_rn(q1, fn), fn);
  }
  return __fmul_rn(__fadd_rn(__fmul_rn(v, voxel_size), -1.f), cube_radius);
}

int hm_simt_decode(hm_context* ctx, const hm_rows& rows, float* d_sdf, float* d_jac, cudaStream_t st,
                   float* h_absmax_out /* [16] or NULL: calibration */, float* h_unit_max_out = nullptr /* [8][512] or NULL */);
int hm_tc_decode(hm_context* ctx, const hm_rows& rows, float* d_sdf, float* d_jac, cudaStream_t st);
int hm_tc_init(hm_context* ctx);          // build weight blob + plan from ctx->h_W and act_absmax
void hm_tc_free(hm_context* ctx);
void hm_tc_plan_info(const hm_context* ctx, double* out /* [12], hm_plan_info */);
int hm_decode(hm_context* ctx, const hm_rows& rows, float* d_sdf, float* d_jac, cudaStream_t st);

// ---- optimiser (optimizer.cu) ----
int hm_optimize_impl(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* b, bool joint, cudaStream_t st);
