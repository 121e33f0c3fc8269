// C-ABI entry points (include/hortimapping_b200.h): context lifecycle, decoder calls, grid.
#include <stdarg.h>
#include <string.h>

#include <cmath>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void hm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* hm_last_error(void) { return g_err; }
extern "C" int hm_version(void) { return 200; }

int hm_ws_reserve(hm_context* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return HM_OK;
  if (ctx->ws) {
    HM_CUDA(cudaDeviceSynchronize());
    HM_CUDA(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
  }
  bytes = (bytes + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);
  HM_CUDA(cudaMalloc(&ctx->ws, bytes));
  ctx->ws_bytes = bytes;
  return HM_OK;
}

int hm_ws2_reserve(hm_context* ctx, size_t bytes) {
  if (bytes <= ctx->ws2_bytes) return HM_OK;
  if (ctx->ws2) {
    HM_CUDA(cudaDeviceSynchronize());
    HM_CUDA(cudaFree(ctx->ws2));
    ctx->ws2 = nullptr;
    ctx->ws2_bytes = 0;
  }
  bytes = (bytes + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);
  HM_CUDA(cudaMalloc(&ctx->ws2, bytes));
  ctx->ws2_bytes = bytes;
  return HM_OK;
}

static const int kInDim[HM_LAYERS] = {35, 512, 512, 512, 512, 512, 512, 512, 512};
static const int kOutDim[HM_LAYERS] = {512, 512, 512, 477, 512, 512, 512, 512, 1};

extern "C" int hm_create(hm_context** out, int device, const hm_decoder_desc* dec) {
  HM_CHECK(out && dec, "hm_create: null argument");
  *out = nullptr;
  HM_CHECK(dec->n_layers == HM_LAYERS && dec->latent_size == HM_LATENT && dec->latent_in_layer == 4,
           "hm_create: unsupported decoder (need 9 linear layers, latent 32, latent_in=[4]; got %d layers, latent %d, latent_in %d)",
           dec->n_layers, dec->latent_size, dec->latent_in_layer);
  for (int l = 0; l < HM_LAYERS; ++l) {
    HM_CHECK(dec->in_dim[l] == kInDim[l] && dec->out_dim[l] == kOutDim[l],
             "hm_create: layer %d is %dx%d, expected %dx%d", l, dec->out_dim[l], dec->in_dim[l], kOutDim[l], kInDim[l]);
    HM_CHECK(dec->weight[l] && dec->bias[l], "hm_create: layer %d has a null weight/bias pointer", l);
  }
  int ndev = 0;
  HM_CUDA(cudaGetDeviceCount(&ndev));
  HM_CHECK(device >= 0 && device < ndev, "hm_create: device %d out of range (%d devices)", device, ndev);
  HM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  HM_CUDA(cudaGetDeviceProperties(&prop, device));
  HM_CHECK(prop.major == 10, "hm_create: this library is built for sm_100a (Blackwell B200); device %d is sm_%d%d",
           device, prop.major, prop.minor);
  hm_context* ctx = new hm_context();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  for (int l = 0; l < HM_LAYERS; ++l) {
    int od = (l == 3) ? HM_HIDDEN : kOutDim[l];    // lin3 zero-padded to 512 rows
    ctx->in_dim[l] = kInDim[l];
    ctx->out_dim[l] = od;
    ctx->h_W[l].assign((size_t)od * kInDim[l], 0.f);
    ctx->h_b[l].assign(od, 0.f);
    memcpy(ctx->h_W[l].data(), dec->weight[l], sizeof(float) * kOutDim[l] * kInDim[l]);
    memcpy(ctx->h_b[l].data(), dec->bias[l], sizeof(float) * kOutDim[l]);
    for (float v : ctx->h_W[l]) {
      if (!std::isfinite(v)) { hm_set_error("hm_create: non-finite weight in layer %d", l); hm_destroy(ctx); return HM_ERR_INVALID; }
    }
    if (cudaMalloc(&ctx->d_W[l], sizeof(float) * ctx->h_W[l].size()) != cudaSuccess ||
        cudaMalloc(&ctx->d_b[l], sizeof(float) * od) != cudaSuccess) {
      hm_set_error("hm_create: cudaMalloc of layer %d failed", l);
      hm_destroy(ctx);
      return HM_ERR_NOMEM;
    }
    cudaMemcpy(ctx->d_W[l], ctx->h_W[l].data(), sizeof(float) * ctx->h_W[l].size(), cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->d_b[l], ctx->h_b[l].data(), sizeof(float) * od, cudaMemcpyHostToDevice);
  }
  // default calibration of the tensor-core operand scales: random latents ~ N(0, 0.1), xyz ~ U(-0.12, 0.12)
  {
    const int n = 8192;
    std::vector<float> rows((size_t)n * HM_IN);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto u01 = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)((s >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 53); };
    for (int i = 0; i < n; ++i) {
      for (int c = 0; c < HM_LATENT; ++c) {
        double g = std::sqrt(-2.0 * std::log(u01() + 1e-12)) * std::cos(6.283185307179586 * u01());
        rows[(size_t)i * HM_IN + c] = (float)(0.1 * g);
      }
      for (int c = 0; c < 3; ++c) rows[(size_t)i * HM_IN + HM_LATENT + c] = (float)((u01() * 2 - 1) * 0.12);
    }
    float* d_rows = nullptr;
    if (cudaMalloc(&d_rows, sizeof(float) * rows.size()) != cudaSuccess) { hm_set_error("hm_create: cudaMalloc failed"); hm_destroy(ctx); return HM_ERR_NOMEM; }
    cudaMemcpy(d_rows, rows.data(), sizeof(float) * rows.size(), cudaMemcpyHostToDevice);
    int rc = hm_calibrate(ctx, d_rows, n, nullptr);
    cudaFree(d_rows);
    if (rc) { hm_destroy(ctx); return rc; }
  }
  *out = ctx;
  return HM_OK;
}

extern "C" void hm_destroy(hm_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int l = 0; l < HM_LAYERS; ++l) {
    if (ctx->d_W[l]) cudaFree(ctx->d_W[l]);
    if (ctx->d_b[l]) cudaFree(ctx->d_b[l]);
  }
  hm_tc_free(ctx);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->ws2) cudaFree(ctx->ws2);
  if (ctx->io_arena) cudaFree(ctx->io_arena);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->stage_event) cudaEventDestroy(ctx->stage_event);
  if (ctx->busy_event) cudaEventDestroy(ctx->busy_event);
  if (ctx->mesh_ws) cudaFree(ctx->mesh_ws);
  if (ctx->mesh_out) cudaFree(ctx->mesh_out);
  for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
  if (ctx->d_last_H) cudaFree(ctx->d_last_H);
  if (ctx->d_last_b) cudaFree(ctx->d_last_b);
  if (ctx->d_last_dx) cudaFree(ctx->d_last_dx);
  delete ctx;
}

extern "C" int hm_set_engine(hm_context* ctx, int engine) {
  HM_CHECK(ctx, "hm_set_engine: null context");
  HM_CHECK(engine == HM_ENGINE_TC || engine == HM_ENGINE_SIMT, "hm_set_engine: unknown engine %d", engine);
  ctx->engine = engine;
  return HM_OK;
}

extern "C" int hm_get_engine(const hm_context* ctx) { return ctx ? ctx->engine : HM_ERR_INVALID; }

extern "C" int hm_set_sparse_plan(hm_context* ctx, int on) {
  HM_CHECK(ctx, "hm_set_sparse_plan: null context");
  ctx->sparse_plan = on ? 1 : 0;
  return HM_OK;
}

extern "C" int hm_set_mask_reuse(hm_context* ctx, int on) {
  HM_CHECK(ctx, "hm_set_mask_reuse: null context");
  ctx->mask_reuse = on ? 1 : 0;
  return HM_OK;
}

extern "C" int hm_plan_info(const hm_context* ctx, double* h_out) {
  HM_CHECK(ctx && h_out, "hm_plan_info: null argument");
  hm_tc_plan_info(ctx, h_out);
  return HM_OK;
}

extern "C" int hm_get_counters(hm_context* ctx, hm_counters* out) {
  HM_CHECK(ctx && out, "hm_get_counters: null argument");
  HM_CUDA(cudaSetDevice(ctx->device));
  HM_CUDA(cudaDeviceSynchronize());
  for (size_t i = 0; i + 1 < ctx->prof_events.size(); i += 2) {
    float ms = 0.f;
    HM_CUDA(cudaEventSynchronize(ctx->prof_events[i + 1]));
    HM_CUDA(cudaEventElapsedTime(&ms, ctx->prof_events[i], ctx->prof_events[i + 1]));
    ctx->counters.decoder_ms += ms;
    ctx->counters.decoder_launches += 1;
    if (ctx->prof_kinds[i / 2] == 2) { ctx->counters.backward_ms += ms; ctx->counters.backward_launches += 1; }
    else if (ctx->prof_kinds[i / 2]) { ctx->counters.jacobian_ms += ms; ctx->counters.jacobian_launches += 1; }
    else { ctx->counters.forward_ms += ms; ctx->counters.forward_launches += 1; }
    ctx->prof_pool.push_back(ctx->prof_events[i]);
    ctx->prof_pool.push_back(ctx->prof_events[i + 1]);
  }
  ctx->prof_events.clear();
  ctx->prof_kinds.clear();
  *out = ctx->counters;
  if (ctx->d_tc_flags) {      // exact device-side totals of the tensor-core engine (HM_TC_FLAG_* slots, common.cuh)
    unsigned long long dev[9] = {};
    HM_CUDA(cudaMemcpy(dev, ctx->d_tc_flags + HM_TC_FLAG_ROWS_FWD, sizeof(dev), cudaMemcpyDeviceToHost));
    out->rows_forward += (int64_t)dev[0];
    out->rows_jacobian += (int64_t)dev[1];
    out->tiles_forward = (int64_t)dev[2];
    out->tiles_jacobian = (int64_t)dev[3];
    out->tiles_redone_forward = (int64_t)dev[4];
    out->tiles_redone_jacobian = (int64_t)dev[5];
    out->rows_backward = (int64_t)dev[6];
    out->tiles_backward = (int64_t)dev[7];
    out->tiles_redone_backward = (int64_t)dev[8];
  }
  return HM_OK;
}

extern "C" int hm_saturation_count(hm_context* ctx, int64_t* h_count) {
  HM_CHECK(ctx && h_count, "hm_saturation_count: null argument");
  HM_CUDA(cudaSetDevice(ctx->device));
  *h_count = 0;
  if (!ctx->d_tc_flags) return HM_OK;
  int32_t v = 0;
  HM_CUDA(cudaDeviceSynchronize());
  HM_CUDA(cudaMemcpy(&v, ctx->d_tc_flags, sizeof(int32_t), cudaMemcpyDeviceToHost));
  HM_CUDA(cudaMemset(ctx->d_tc_flags, 0, sizeof(int32_t)));
  *h_count = v;
  return HM_OK;
}

extern "C" int hm_profile_enable(hm_context* ctx, int on) {
  HM_CHECK(ctx, "hm_profile_enable: null context");
  ctx->profiling = on ? 1 : 0;
  return HM_OK;
}

extern "C" int hm_calibrate(hm_context* ctx, const float* d_rows, int64_t n, void* stream) {
  HM_CHECK(ctx && d_rows && n > 0, "hm_calibrate: bad argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CUDA(cudaSetDevice(ctx->device));
  hm_rows rows = {d_rows, nullptr, nullptr, nullptr, n, nullptr};
  float absmax[16];
  // sdf output is discarded: put it at the end of a scratch allocation
  float* d_sdf = nullptr;
  HM_CUDA(cudaMalloc(&d_sdf, sizeof(float) * n));
  std::vector<float> unit_max(8 * HM_HIDDEN, 0.f);
  int rc = hm_simt_decode(ctx, rows, d_sdf, nullptr, (cudaStream_t)stream, absmax, unit_max.data());
  cudaFree(d_sdf);
  if (rc) return rc;
  for (int i = 0; i < 16; ++i) ctx->act_absmax[i] = absmax[i];
  ctx->unit_max = unit_max;
  return hm_tc_init(ctx);
}

int hm_decode(hm_context* ctx, const hm_rows& rows, float* d_sdf, float* d_jac, cudaStream_t st) {
  if (rows.n <= 0) return HM_OK;
  if (ctx->engine == HM_ENGINE_SIMT) return hm_simt_decode(ctx, rows, d_sdf, d_jac, st, nullptr);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->profiling) {
    auto get = [&](cudaEvent_t& e) -> cudaError_t {
      if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); return cudaSuccess; }
      return cudaEventCreate(&e);
    };
    HM_CUDA(get(e0));
    HM_CUDA(get(e1));
    HM_CUDA(cudaEventRecord(e0, st));
  }
  int rc = hm_tc_decode(ctx, rows, d_sdf, d_jac, st);
  if (ctx->profiling) {
    HM_CUDA(cudaEventRecord(e1, st));
    ctx->prof_events.push_back(e0);
    ctx->prof_events.push_back(e1);
    ctx->prof_kinds.push_back(rows.d_mask_in ? 2 : d_jac ? 1 : 0);
  }
  return rc;
}

extern "C" int hm_sdf_forward(hm_context* ctx, const float* d_latent, const float* d_xyz, int64_t n, float* d_sdf, void* stream) {
  HM_CHECK(ctx && d_latent && (n == 0 || (d_xyz && d_sdf)) && n >= 0, "hm_sdf_forward: bad argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CUDA(cudaSetDevice(ctx->device));
  hm_rows rows = {nullptr, d_xyz, d_latent, nullptr, n, nullptr};
  return hm_decode(ctx, rows, d_sdf, nullptr, (cudaStream_t)stream);
}

extern "C" int hm_sdf_forward_rows(hm_context* ctx, const float* d_rows, int64_t n, float* d_sdf, void* stream) {
  HM_CHECK(ctx && (n == 0 || (d_rows && d_sdf)) && n >= 0, "hm_sdf_forward_rows: bad argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CUDA(cudaSetDevice(ctx->device));
  hm_rows rows = {d_rows, nullptr, nullptr, nullptr, n, nullptr};
  return hm_decode(ctx, rows, d_sdf, nullptr, (cudaStream_t)stream);
}

extern "C" int hm_sdf_jacobian(hm_context* ctx, const float* d_latent, const float* d_xyz, int64_t n, float* d_sdf,
                               float* d_jac, void* stream) {
  HM_CHECK(ctx && d_latent && (n == 0 || (d_xyz && d_sdf && d_jac)) && n >= 0, "hm_sdf_jacobian: bad argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CUDA(cudaSetDevice(ctx->device));
  hm_rows rows = {nullptr, d_xyz, d_latent, nullptr, n, nullptr};
  return hm_decode(ctx, rows, d_sdf, d_jac, (cudaStream_t)stream);
}

extern "C" int hm_sdf_jacobian_rows(hm_context* ctx, const float* d_rows, int64_t n, float* d_sdf, float* d_jac, void* stream) {
  HM_CHECK(ctx && (n == 0 || (d_rows && d_sdf && d_jac)) && n >= 0, "hm_sdf_jacobian_rows: bad argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CUDA(cudaSetDevice(ctx->device));
  hm_rows rows = {d_rows, nullptr, nullptr, nullptr, n, nullptr};
  return hm_decode(ctx, rows, d_sdf, d_jac, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Mesher grid: wild_completion/utils.py:542-562 create_voxel_grid, scaled by cube_radius
// (mesher.py:12).  `LongTensor / int` is true division, so columns 0 and 1 are fractional (the
// "sheared" grid); the fp32 op order of the reference is kept so the points are
// bit-identical: idx -> float32, / n, fmod n, * voxel_size, + (-1), * cube_radius.
// ---------------------------------------------------------------------------------------------
__global__ void voxel_grid_kernel(int n, float voxel_size, float cube_radius, int64_t total, float* __restrict__ xyz) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  for (int c = 0; c < 3; ++c) xyz[i * 3 + c] = hm_grid_coord(i, c, n, voxel_size, cube_radius);
}

extern "C" int hm_voxel_grid(hm_context* ctx, int32_t vol_dim, float cube_radius, float* d_xyz, void* stream) {
  HM_CHECK(ctx && d_xyz && vol_dim >= 2 && vol_dim <= 1024, "hm_voxel_grid: bad argument");
  HM_CUDA(cudaSetDevice(ctx->device));
  int64_t total = (int64_t)vol_dim * vol_dim * vol_dim;
  float voxel_size = (float)(2.0 / (vol_dim - 1));
  voxel_grid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(vol_dim, voxel_size, cube_radius, total, d_xyz);
  ctx->counters.kernel_launches += 1;
  HM_CUDA(cudaGetLastError());
  return HM_OK;
}

extern "C" int hm_sdf_grid(hm_context* ctx, const float* d_latent, int32_t vol_dim, float cube_radius, float* d_sdf, void* stream) {
  HM_CHECK(ctx && d_latent && d_sdf && vol_dim >= 2 && vol_dim <= 1024, "hm_sdf_grid: bad argument");
  hm_stream_scope scope_(ctx, (cudaStream_t)stream);
  HM_CUDA(cudaSetDevice(ctx->device));
  int64_t total = (int64_t)vol_dim * vol_dim * vol_dim;
  hm_rows rows = {nullptr, nullptr, d_latent, nullptr, total, nullptr};
  rows.grid_n = vol_dim;
  rows.grid_voxel = (float)(2.0 / (vol_dim - 1));
  rows.grid_radius = cube_radius;
  if (ctx->engine == HM_ENGINE_SIMT) {          // the validation engine reads explicit points
    int rc = hm_ws2_reserve(ctx, sizeof(float) * 3 * total);
    if (rc) return rc;
    rc = hm_voxel_grid(ctx, vol_dim, cube_radius, (float*)ctx->ws2, stream);
    if (rc) return rc;
    rows.d_xyz = (const float*)ctx->ws2;
    rows.grid_n = 0;
  }
  // tensor-core engine: the grid points are generated inside the decoder kernel (fused grid sample + decode)
  return hm_decode(ctx, rows, d_sdf, nullptr, (cudaStream_t)stream);
}
