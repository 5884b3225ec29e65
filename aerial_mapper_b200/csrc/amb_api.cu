// amb_api.cu — the extern "C" boundary (include/aerial_mapper_b200.h): context, layers, host entry points.
#include <cmath>
#include <cstdlib>
#include <limits>
#include <new>

#include "amb_context.h"

namespace amb {

__global__ void fill_kernel(float* p, size_t n, float v) {
  for (size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<size_t>(gridDim.x) * blockDim.x)
    p[k] = v;
}

int wait_layer_copy(amb_ctx* ctx, int layer) {
  if (layer >= 0 && layer < AMB_NUM_LAYERS && ctx->layer_copy_pending[layer]) {
    AMB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->layer_copy_event[layer], 0));
    ctx->layer_copy_pending[layer] = false;
  }
  return AMB_OK;
}

int enqueue_layer_download(amb_ctx* ctx, int layer, float* host_slab) {
  // ordered after everything enqueued so far on the compute stream, executed on the copy stream so that it
  // overlaps later kernels and host->device copies (PCIe is full duplex)
  AMB_CUDA(ctx, cudaEventRecord(ctx->copy_done[0], ctx->stream));
  AMB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done[0], 0));
  AMB_CUDA(ctx, cudaMemcpyAsync(host_slab, ctx->layers[layer], ctx->slab_cells() * sizeof(float),
                                cudaMemcpyDeviceToHost, ctx->copy_stream));
  // later WRITERS of this layer on the compute stream wait for the copy (wait_layer_copy); readers do not
  if (!ctx->layer_copy_event[layer])
    AMB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->layer_copy_event[layer], cudaEventDisableTiming));
  AMB_CUDA(ctx, cudaEventRecord(ctx->layer_copy_event[layer], ctx->copy_stream));
  ctx->layer_copy_pending[layer] = true;
  return AMB_OK;
}

// Columns [col0, col1) of the slab (slab-local) to the layer's host mirror: the part of a layer that is already final
// while the rest is still being computed (chunked DSM, amb_dsm_set_stream_chunks).  Contiguous: layers are column-major.
int mirror_layer_columns(amb_ctx* ctx, int layer, int col0, int col1) {
  if (layer < 0 || layer >= AMB_NUM_LAYERS || !ctx->host_mirror[layer] || !ctx->layers[layer] || col1 <= col0)
    return AMB_OK;
  // column chunks always travel as float32; the expander pool may still be widening an earlier compact round into the mirror
  if (ctx->compact[layer].enabled) wait_compact_layer(ctx, layer);
  const size_t off = static_cast<size_t>(ctx->geom.rows) * static_cast<size_t>(col0);
  const size_t cnt = static_cast<size_t>(ctx->geom.rows) * static_cast<size_t>(col1 - col0);
  AMB_CUDA(ctx, cudaEventRecord(ctx->copy_done[0], ctx->stream));
  AMB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done[0], 0));
  AMB_CUDA(ctx, cudaMemcpyAsync(ctx->host_mirror[layer] + off, ctx->layers[layer] + off, cnt * sizeof(float),
                                cudaMemcpyDeviceToHost, ctx->copy_stream));
  if (!ctx->layer_copy_event[layer])
    AMB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->layer_copy_event[layer], cudaEventDisableTiming));
  AMB_CUDA(ctx, cudaEventRecord(ctx->layer_copy_event[layer], ctx->copy_stream));
  ctx->layer_copy_pending[layer] = true;
  return AMB_OK;
}

int mirror_layer(amb_ctx* ctx, int layer) {
  if (layer < 0 || layer >= AMB_NUM_LAYERS || !ctx->host_mirror[layer] || !ctx->layers[layer]) return AMB_OK;
  if (ctx->compact[layer].enabled && (layer == AMB_LAYER_ORTHO || layer == AMB_LAYER_OBSERVATION_INDEX))
    return mirror_layer_compact(ctx, layer);  // one byte per cell when every value has a code (amb_set_host_mirror_compact)
  return enqueue_layer_download(ctx, layer, ctx->host_mirror[layer]);
}

int ensure_counters(amb_ctx* ctx) {
  if (ctx->counters.ptr) return AMB_OK;
  AMB_CUDA(ctx, ctx->counters.reserve(CTR_COUNT * sizeof(unsigned int)));
  AMB_CUDA(ctx, cudaMemsetAsync(ctx->counters.ptr, 0, CTR_COUNT * sizeof(unsigned int), ctx->stream));
  return AMB_OK;
}

int ensure_layer(amb_ctx* ctx, int layer) {
  if (layer < 0 || layer >= AMB_NUM_LAYERS) return AMB_ERR_INVALID_ARGUMENT;
  if (ctx->layers[layer]) return AMB_OK;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  float* p = nullptr;
  AMB_CUDA(ctx, cudaMalloc(&p, ctx->slab_cells() * sizeof(float)));
  ctx->layers[layer] = p;
  // A layer that was never uploaded starts with AerialGridMap's initial value (aerial-mapper-grid-map.cc:40-48).
  float v = std::numeric_limits<float>::quiet_NaN();
  if (layer == AMB_LAYER_ORTHO) v = 255.0f;
  if (layer == AMB_LAYER_ELEVATION_ANGLE || layer == AMB_LAYER_NUM_OBSERVATIONS) v = 0.0f;
  fill_kernel<<<kNumSMsB200 * 4, 256, 0, ctx->stream>>>(p, ctx->slab_cells(), v);
  AMB_CUDA(ctx, cudaGetLastError());
  return AMB_OK;
}

// The flag travels through host-MAPPED memory (a store from a one-thread kernel), not through the device->host
// copy engine, where it would queue behind result layers still streaming to their host mirrors.
__global__ void flag_to_host_kernel(const unsigned int* __restrict__ src, unsigned int* __restrict__ dst) {
  *dst = *src;
  __threadfence_system();
}

static int finish_flags(amb_ctx* ctx, unsigned int offset_words, int err_code) {
  if (!ctx->host_flags)
    AMB_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->host_flags), 64, cudaHostAllocMapped));
  unsigned int* mapped = nullptr;
  AMB_CUDA(ctx, cudaHostGetDevicePointer(reinterpret_cast<void**>(&mapped), ctx->host_flags, 0));
  ctx->host_flags[0] = 0;
  flag_to_host_kernel<<<1, 1, 0, ctx->stream>>>(ctx->counters.as<unsigned int>() + offset_words, mapped);
  AMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (!ctx->host_flags[0]) return AMB_OK;
  // reported: the sticky flag is cleared (ordered before any later kernel on the stream)
  AMB_CUDA(ctx, cudaMemsetAsync(ctx->counters.as<unsigned int>() + offset_words, 0, sizeof(unsigned int), ctx->stream));
  return err_code;
}

}  // namespace amb

using namespace amb;

extern "C" {

int amb_abi_version(void) { return AMB_ABI_VERSION; }

const char* amb_status_string(int status) {
  switch (status) {
    case AMB_OK: return "ok";
    case AMB_ERR_EMPTY: return "empty input";
    case AMB_ERR_SIZE_MISMATCH: return "size mismatch";
    case AMB_ERR_COINCIDENT_POINT: return "a point coincides with a cell centre (reference CHECK(distances[i] > 0.0))";
    case AMB_ERR_CUDA: return "CUDA error";
    case AMB_ERR_INVALID_ARGUMENT: return "invalid argument";
    case AMB_ERR_CHECK_FAILED: return "reference CHECK(alpha > 0.0) failed";
    case AMB_ERR_NO_DEVICE: return "no CUDA device";
    case AMB_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

const char* amb_last_error(const amb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int amb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return AMB_ERR_NO_DEVICE;
  }
  return n;
}

int amb_geometry_init(double delta_easting, double delta_northing, double resolution, double center_easting,
                      double center_northing, amb_geometry* out) {
  if (!out || !(resolution > 0.0) || !(delta_easting > 0.0) || !(delta_northing > 0.0))
    return AMB_ERR_INVALID_ARGUMENT;
  // grid_map::GridMap::setGeometry: size = round(length / resolution); length = size * resolution.
  out->rows = static_cast<int32_t>(std::round(delta_easting / resolution));
  out->cols = static_cast<int32_t>(std::round(delta_northing / resolution));
  if (out->rows <= 0 || out->cols <= 0) return AMB_ERR_INVALID_ARGUMENT;
  out->resolution = resolution;
  out->length_x = out->rows * resolution;
  out->length_y = out->cols * resolution;
  out->pos_x = center_easting;
  out->pos_y = center_northing;
  return AMB_OK;
}

int amb_geometry_position(const amb_geometry* g, int32_t i, int32_t j, double* x, double* y) {
  if (!g || !x || !y) return AMB_ERR_INVALID_ARGUMENT;
  if (i < 0 || j < 0 || i >= g->rows || j >= g->cols) return AMB_ERR_SIZE_MISMATCH;
  // grid_map::getPositionFromIndex: (mapPosition + (0.5*length - 0.5*res)) + res * (-(double)index)
  *x = (g->pos_x + (0.5 * g->length_x - 0.5 * g->resolution)) + g->resolution * (-static_cast<double>(i));
  *y = (g->pos_y + (0.5 * g->length_y - 0.5 * g->resolution)) + g->resolution * (-static_cast<double>(j));
  return AMB_OK;
}

int amb_create(const amb_geometry* geom, int device, int32_t col_begin, int32_t col_end, amb_ctx** out) {
  if (!geom || !out) return AMB_ERR_INVALID_ARGUMENT;
  if (geom->rows <= 0 || geom->cols <= 0 || !(geom->resolution > 0.0)) return AMB_ERR_INVALID_ARGUMENT;
  if (col_begin < 0 || col_end > geom->cols || col_begin >= col_end) return AMB_ERR_SIZE_MISMATCH;
  if (static_cast<size_t>(geom->rows) * static_cast<size_t>(col_end - col_begin) >= size_t(0xffffffffu))
    return AMB_ERR_UNSUPPORTED;
  int n = amb_device_count();
  if (n <= 0) return AMB_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return AMB_ERR_INVALID_ARGUMENT;
  amb_ctx* ctx = new (std::nothrow) amb_ctx();
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  ctx->geom = *geom;
  ctx->device = device;
  ctx->col_begin = col_begin;
  ctx->col_end = col_end;
  // process-wide override of the gather's default arithmetic (amb_dsm_set_precision still wins per context): lets the
  // same test suite / demo binary run in either mode without code changes
  if (const char* p = std::getenv("AMB_DSM_PRECISION")) {
    if (p[0] == 'f' && p[1] == '6') ctx->dsm_precision = AMB_DSM_F64;
    if (p[0] == 'f' && p[1] == '3') ctx->dsm_precision = AMB_DSM_F32;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  for (int k = 0; k < EV_COUNT && e == cudaSuccess; ++k) e = cudaEventCreate(&ctx->events[k]);
  for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ctx->copy_done[k], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    cudaGetLastError();
    amb_destroy(ctx);
    return AMB_ERR_CUDA;
  }
  *out = ctx;
  return AMB_OK;
}

void amb_destroy(amb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  amb_comm_destroy(ctx);
  ctx->halo_send.release();
  ctx->halo_recv.release();
  release_compact_mirrors(ctx);
  for (int l = 0; l < AMB_NUM_LAYERS; ++l)
    if (ctx->layers[l]) cudaFree(ctx->layers[l]);
  DeviceBuffer* bufs[] = {&ctx->points,  &ctx->point_ids, &ctx->intensities, &ctx->records, &ctx->records_tmp, &ctx->tile_offsets, &ctx->point_order, &ctx->bucket_flags,
                            &ctx->bin_starts, &ctx->block_sums, &ctx->empty_cells,
                          &ctx->counters, &ctx->dbg_count, &ctx->dbg_level,  &ctx->frames,     &ctx->frame_table,
                          &ctx->frame_cull, &ctx->frame_rects, &ctx->ortho_pix, &ctx->ortho_bbox};
  for (DeviceBuffer* b : bufs) b->release();
  for (int k = 0; k < EV_COUNT; ++k)
    if (ctx->events[k]) cudaEventDestroy(ctx->events[k]);
  for (int k = 0; k < 2; ++k)
    if (ctx->copy_done[k]) cudaEventDestroy(ctx->copy_done[k]);
  for (int l = 0; l < AMB_NUM_LAYERS; ++l)
    if (ctx->layer_copy_event[l]) cudaEventDestroy(ctx->layer_copy_event[l]);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  for (int k = 0; k < 2; ++k) {
    ctx->stages[k].release();
    if (ctx->stage_events[k]) cudaEventDestroy(ctx->stage_events[k]);
  }
  if (ctx->host_flags) cudaFreeHost(ctx->host_flags);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

int amb_sync(amb_ctx* ctx) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  AMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  AMB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  const int compact_status = join_compact_mirrors(ctx);  // host threads widening one-byte codes (opt-in; none otherwise)
  for (int l = 0; l < AMB_NUM_LAYERS; ++l) ctx->layer_copy_pending[l] = false;
  if (compact_status != AMB_OK) return compact_status;
  // Deferred reference CHECKs of the asynchronous `_device` entry points.
  if (ctx->counters.ptr) {
    unsigned int c[CTR_COUNT];
    AMB_CUDA(ctx, cudaMemcpy(c, ctx->counters.ptr, sizeof(c), cudaMemcpyDeviceToHost));
    if (c[CTR_DSM_COINCIDENT] || c[CTR_ORTHO_CHECK] || c[CTR_HALO_OVERFLOW]) {  // reported once, then cleared
      unsigned int* d = ctx->counters.as<unsigned int>();
      AMB_CUDA(ctx, cudaMemset(d + CTR_DSM_COINCIDENT, 0, sizeof(unsigned int)));
      AMB_CUDA(ctx, cudaMemset(d + CTR_ORTHO_CHECK, 0, sizeof(unsigned int)));
      AMB_CUDA(ctx, cudaMemset(d + CTR_HALO_OVERFLOW, 0, sizeof(unsigned int)));
    }
    if (c[CTR_HALO_OVERFLOW]) {
      ctx->last_error = "a rank's border halo exceeded halo_capacity (amb_dsm_process_sharded_device): result incomplete";
      return AMB_ERR_SIZE_MISMATCH;
    }
    if (c[CTR_DSM_COINCIDENT]) return AMB_ERR_COINCIDENT_POINT;
    if (c[CTR_ORTHO_CHECK]) return AMB_ERR_CHECK_FAILED;
  }
  return AMB_OK;
}

void* amb_stream(amb_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }

int amb_init_layers(amb_ctx* ctx) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const float nan = std::numeric_limits<float>::quiet_NaN();
  for (int l = 0; l < AMB_NUM_LAYERS; ++l) {
    // Slabs are allocated lazily and a fresh slab is born with its initial value (ensure_layer), so only the
    // ones that already exist need refilling.
    if (!ctx->layers[l]) continue;
    wait_layer_copy(ctx, l);
    float v = nan;
    if (l == AMB_LAYER_ORTHO) v = 255.0f;
    if (l == AMB_LAYER_ELEVATION_ANGLE || l == AMB_LAYER_NUM_OBSERVATIONS) v = 0.0f;
    fill_kernel<<<kNumSMsB200 * 4, 256, 0, ctx->stream>>>(ctx->layers[l], ctx->slab_cells(), v);
  }
  AMB_CUDA(ctx, cudaGetLastError());
  return AMB_OK;
}

int amb_upload_layer(amb_ctx* ctx, int layer, const float* host_slab) {
  if (!ctx || !host_slab) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  int st = ensure_layer(ctx, layer);
  if (st != AMB_OK) return st;
  wait_layer_copy(ctx, layer);
  st = staged_h2d(ctx, ctx->layers[layer], host_slab, ctx->slab_cells() * sizeof(float), ctx->stream);
  if (st != AMB_OK) return st;
  AMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMB_OK;
}

int amb_upload_layer_device(amb_ctx* ctx, int layer, const float* device_slab) {
  if (!ctx || !device_slab) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  int st = ensure_layer(ctx, layer);
  if (st != AMB_OK) return st;
  wait_layer_copy(ctx, layer);
  AMB_CUDA(ctx, cudaMemcpyAsync(ctx->layers[layer], device_slab, ctx->slab_cells() * sizeof(float),
                                cudaMemcpyDeviceToDevice, ctx->stream));
  return AMB_OK;
}

int amb_download_layer(amb_ctx* ctx, int layer, float* host_slab) {
  if (!ctx || !host_slab) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  int st = ensure_layer(ctx, layer);
  if (st != AMB_OK) return st;
  return staged_d2h(ctx, host_slab, ctx->layers[layer], ctx->slab_cells() * sizeof(float), ctx->stream);
}

int amb_download_layer_async(amb_ctx* ctx, int layer, float* host_slab) {
  if (!ctx || !host_slab) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  int st = ensure_layer(ctx, layer);
  if (st != AMB_OK) return st;
  return enqueue_layer_download(ctx, layer, host_slab);
}

int amb_set_host_mirror(amb_ctx* ctx, int layer, float* host_slab) {
  if (!ctx || layer < 0 || layer >= AMB_NUM_LAYERS) return AMB_ERR_INVALID_ARGUMENT;
  ctx->host_mirror[layer] = host_slab;
  return AMB_OK;
}

int amb_layer_device_ptr(amb_ctx* ctx, int layer, float** device_slab) {
  if (!ctx || !device_slab) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  int st = ensure_layer(ctx, layer);
  if (st != AMB_OK) return st;
  *device_slab = ctx->layers[layer];
  return AMB_OK;
}

// ---- DSM ----
int amb_dsm_process_device(amb_ctx* ctx, const double* d_xyz, size_t n, int32_t interpolation_radius,
                           double center_easting, double center_northing) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY;  // dsm.cc:189-192
  if (!d_xyz) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_BEGIN], ctx->stream));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_H2D_END], ctx->stream));
  ctx->dsm_had_h2d = false;
  int st = dsm_run(ctx, d_xyz, nullptr, n, interpolation_radius, center_easting, center_northing);
  ctx->dsm_timed = (st == AMB_OK);
  return st;
}

int amb_dsm_process_device_ids(amb_ctx* ctx, const double* d_xyz, const uint64_t* d_ids, size_t n,
                               int32_t interpolation_radius, double center_easting, double center_northing) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY;
  if (!d_xyz) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_BEGIN], ctx->stream));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_H2D_END], ctx->stream));
  ctx->dsm_had_h2d = false;
  int st = dsm_run(ctx, d_xyz, reinterpret_cast<const unsigned long long*>(d_ids), n, interpolation_radius,
                   center_easting, center_northing);
  ctx->dsm_timed = (st == AMB_OK);
  return st;
}

// ---- OrthoFromPcl ("next" row N1) ----
int amb_ortho_from_pcl_process_device(amb_ctx* ctx, const double* d_xyz, const int32_t* d_intensities, size_t n,
                                      int32_t interpolation_radius, int32_t use_adaptive_interpolation) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY;  // CHECK(!pointcloud.empty()), ortho-from-pcl.cc:23
  if (!d_xyz || !d_intensities) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_BEGIN], ctx->stream));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_H2D_END], ctx->stream));
  ctx->dsm_had_h2d = false;
  float* saved = nullptr;  // adaptive interpolation (ortho-from-pcl.cc:63-72): pcl_adaptive_kernels.cu
  if (use_adaptive_interpolation) {
    const int pst = pcl_adaptive_prepare(ctx, &saved);
    if (pst != AMB_OK) return pcl_adaptive_finish(ctx, pst, n, interpolation_radius, saved);
  }
  int st = dsm_run(ctx, d_xyz, nullptr, n, interpolation_radius, 0.0, 0.0, 1, d_intensities);
  if (use_adaptive_interpolation) st = pcl_adaptive_finish(ctx, st, n, interpolation_radius, saved);
  ctx->dsm_timed = (st == AMB_OK);
  return st;
}

int amb_ortho_from_pcl_process(amb_ctx* ctx, const double* xyz, const int32_t* intensities, size_t n,
                               int32_t interpolation_radius, int32_t use_adaptive_interpolation) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY;
  if (!xyz || !intensities) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  AMB_CUDA(ctx, ctx->points.reserve(n * 3 * sizeof(double)));
  AMB_CUDA(ctx, ctx->intensities.reserve(n * sizeof(int32_t)));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_BEGIN], ctx->stream));
  {
    const int sst = staged_h2d(ctx, ctx->points.ptr, xyz, n * 3 * sizeof(double), ctx->stream);
    if (sst != AMB_OK) return sst;
  }
  AMB_CUDA(ctx, cudaMemcpyAsync(ctx->intensities.ptr, intensities, n * sizeof(int32_t), cudaMemcpyHostToDevice,
                                ctx->stream));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_H2D_END], ctx->stream));
  ctx->dsm_had_h2d = true;
  float* saved = nullptr;
  if (use_adaptive_interpolation) {
    const int pst = pcl_adaptive_prepare(ctx, &saved);
    if (pst != AMB_OK) return pcl_adaptive_finish(ctx, pst, n, interpolation_radius, saved);
  }
  int st = dsm_run(ctx, ctx->points.as<double>(), nullptr, n, interpolation_radius, 0.0, 0.0, 1,
                   ctx->intensities.as<int>());
  if (use_adaptive_interpolation) st = pcl_adaptive_finish(ctx, st, n, interpolation_radius, saved);
  ctx->dsm_timed = (st == AMB_OK);
  if (st != AMB_OK) return st;
  AMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMB_OK;
}

int amb_dsm_set_stream_chunks(amb_ctx* ctx, int chunks) {
  if (!ctx || chunks < 1) return AMB_ERR_INVALID_ARGUMENT;
  ctx->dsm_stream_chunks = chunks;
  return AMB_OK;
}

int amb_dsm_set_precision(amb_ctx* ctx, int precision) {
  if (!ctx || (precision != AMB_DSM_F64 && precision != AMB_DSM_F32)) return AMB_ERR_INVALID_ARGUMENT;
  ctx->dsm_precision = precision;
  return AMB_OK;
}

int amb_dsm_set_density_hint(amb_ctx* ctx, double points_per_cell) {
  if (!ctx || !(points_per_cell >= 0.0)) return AMB_ERR_INVALID_ARGUMENT;
  ctx->dsm_density_hint = points_per_cell;
  return AMB_OK;
}

int amb_stripe_y_interval(const amb_geometry* g, int32_t col_begin, int32_t col_end, double* y_lo, double* y_hi) {
  if (!g || !y_lo || !y_hi || col_begin < 0 || col_end > g->cols || col_begin >= col_end) return AMB_ERR_INVALID_ARGUMENT;
  // cell column j is centred at base_y - res*j; the stripe [col_begin, col_end) covers (y_lo, y_hi]
  const double base_y = g->pos_y + (0.5 * g->length_y - 0.5 * g->resolution);
  *y_hi = base_y - g->resolution * col_begin + 0.5 * g->resolution;
  *y_lo = base_y - g->resolution * (col_end - 1) - 0.5 * g->resolution;
  return AMB_OK;
}

double amb_dsm_halo_reach(const amb_geometry* g, int32_t interpolation_radius) {
  if (!g || interpolation_radius < 1) return -1.0;
  const std::vector<double> thr = dsm_thresholds(interpolation_radius);
  double m = 0.0;
  for (double t : thr) m = std::max(m, t);
  // (i) the largest retry threshold's reach + one cell of slack: what a border CELL can see; (ii) everything a border
  // TILE stages — tiles are aligned to global columns, so a tile cut by the stripe border extends up to tile-width - 1
  // columns beyond it, plus the window apron: with those points present every tile of the stripe stages exactly the
  // point set it stages in the undivided map, so tile-level quantities (shared-memory / warp-per-cell decision, the FP32
  // gather's height offset) are independent of the sharding.
  return std::max(std::sqrt(m) + g->resolution, dsm_tile_reach_cells(g->resolution, interpolation_radius) * g->resolution);
}

int amb_dsm_extract_halo(amb_ctx* ctx, const double* d_xyz, const uint64_t* d_ids, size_t n, double y_lo, double y_hi,
                         double reach, double center_easting, double* d_out_xyz, uint64_t* d_out_ids,
                         uint32_t capacity, uint32_t* d_count) {
  if (!ctx || !d_xyz || !d_out_xyz || !d_out_ids || !d_count) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  return dsm_extract_halo(ctx, d_xyz, reinterpret_cast<const unsigned long long*>(d_ids), n, y_lo, y_hi, reach,
                          center_easting, d_out_xyz, reinterpret_cast<unsigned long long*>(d_out_ids), capacity,
                          d_count);
}

int amb_dsm_process(amb_ctx* ctx, const double* xyz, size_t n, int32_t interpolation_radius,
                    double center_easting, double center_northing) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY;
  if (!xyz) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  AMB_CUDA(ctx, ctx->points.reserve(n * 3 * sizeof(double)));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_BEGIN], ctx->stream));
  int st = staged_h2d(ctx, ctx->points.ptr, xyz, n * 3 * sizeof(double), ctx->stream);  // (pageable clouds: host_staging.cu)
  if (st != AMB_OK) return st;
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_H2D_END], ctx->stream));
  ctx->dsm_had_h2d = true;
  st = dsm_run(ctx, ctx->points.as<double>(), nullptr, n, interpolation_radius, center_easting, center_northing);
  ctx->dsm_timed = (st == AMB_OK);
  if (st != AMB_OK) return st;
  return finish_flags(ctx, CTR_DSM_COINCIDENT, AMB_ERR_COINCIDENT_POINT);
}

int amb_dsm_enable_debug(amb_ctx* ctx, int enable) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  ctx->dsm_debug = enable != 0;
  if (!enable) ctx->dsm_debug_valid = false;
  return AMB_OK;
}

int amb_dsm_download_debug(amb_ctx* ctx, int32_t* neighbour_count, int8_t* threshold_index) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  if (!ctx->dsm_debug_valid) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t cells = ctx->slab_cells();
  if (neighbour_count)
    AMB_CUDA(ctx, cudaMemcpyAsync(neighbour_count, ctx->dbg_count.ptr, cells * sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, ctx->stream));
  if (threshold_index)
    AMB_CUDA(ctx, cudaMemcpyAsync(threshold_index, ctx->dbg_level.ptr, cells, cudaMemcpyDeviceToHost, ctx->stream));
  AMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AMB_OK;
}

int amb_dsm_thresholds(int32_t interpolation_radius, double* thresholds, int32_t capacity) {
  if (interpolation_radius < 1 || !thresholds || capacity <= 0) return AMB_ERR_INVALID_ARGUMENT;
  const std::vector<double> thr = dsm_thresholds(interpolation_radius);
  const int n = static_cast<int>(std::min<size_t>(thr.size(), static_cast<size_t>(capacity)));
  for (int k = 0; k < n; ++k) thresholds[k] = thr[k];
  return n;
}

// ---- Ortho ----
int amb_ortho_process_device(amb_ctx* ctx, const amb_camera* camera, const double* T_G_B,
                             const uint8_t* const* d_images, size_t n, int32_t channels, size_t row_step,
                             int32_t colored_ortho) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_ORTHO_BEGIN], ctx->stream));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_ORTHO_H2D_END], ctx->stream));
  ctx->ortho_had_h2d = false;
  ctx->ortho_two_phase = false;
  int st = ortho_run(ctx, camera, T_G_B, d_images, nullptr, n, channels, row_step, colored_ortho);
  if (st == AMB_OK) AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_ORTHO_END], ctx->stream));
  ctx->ortho_timed = (st == AMB_OK);
  return st;
}

int amb_ortho_process(amb_ctx* ctx, const amb_camera* camera, const double* T_G_B, const uint8_t* const* images,
                      size_t n, int32_t channels, size_t row_step, int32_t colored_ortho) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  if (n == 0) return AMB_ERR_EMPTY;
  if (!camera || !T_G_B || !images) return AMB_ERR_INVALID_ARGUMENT;
  if (camera->width <= 0 || camera->height <= 0) return AMB_ERR_SIZE_MISMATCH;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  for (size_t f = 0; f < n; ++f)
    if (!images[f]) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_ORTHO_BEGIN], ctx->stream));
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_ORTHO_H2D_END], ctx->stream));
  ctx->ortho_had_h2d = true;
  ctx->ortho_two_phase = true;
  int st = ortho_run(ctx, camera, T_G_B, nullptr, images, n, channels, row_step, colored_ortho);
  if (st != AMB_OK) return st;
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_ORTHO_END], ctx->stream));
  ctx->ortho_timed = true;
  return finish_flags(ctx, CTR_ORTHO_CHECK, AMB_ERR_CHECK_FAILED);
}

int amb_ortho_set_brute_force(amb_ctx* ctx, int brute_force) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  ctx->ortho_brute_force = brute_force != 0;
  return AMB_OK;
}

int amb_ortho_set_dominance_cull(amb_ctx* ctx, int enable) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  ctx->ortho_dominance = enable != 0;
  return AMB_OK;
}

// ---- measurement ----
int amb_get_timings(amb_ctx* ctx, amb_timings* out) {
  if (!ctx || !out) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  AMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::memset(out, 0, sizeof(*out));
  auto ms = [&](int a, int b) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, ctx->events[a], ctx->events[b]) != cudaSuccess) {
      cudaGetLastError();
      t = 0.f;
    }
    return t;
  };
  if (ctx->dsm_timed) {
    out->dsm_h2d_ms = ctx->dsm_had_h2d ? ms(EV_DSM_BEGIN, EV_DSM_H2D_END) : 0.f;
    out->dsm_bin_ms = ms(EV_DSM_H2D_END, EV_DSM_BIN_END);
    out->dsm_gather_ms = ms(EV_DSM_BIN_END, EV_DSM_GATHER_END);
    out->dsm_fill_ms = ms(EV_DSM_GATHER_END, EV_DSM_FILL_END);
    out->dsm_total_ms = ms(EV_DSM_BEGIN, EV_DSM_FILL_END);
    out->dsm_kernel_launches = ctx->dsm_launches;
    unsigned int c[CTR_COUNT] = {};
    if (ctx->counters.ptr) {
      AMB_CUDA(ctx, cudaMemcpy(c, ctx->counters.ptr, sizeof(c), cudaMemcpyDeviceToHost));
    }
    out->dsm_cells_empty = static_cast<int64_t>(c[CTR_DSM_LIST]) + c[CTR_DSM_LIST_DONE];
    out->dsm_points_binned = c[CTR_DSM_BINNED];
  }
  if (ctx->ortho_timed) {
    if (ctx->ortho_two_phase) {  // select kernel | sub-rectangle copies | texel kernel
      out->ortho_h2d_ms = ms(EV_ORTHO_SELECT_END, EV_ORTHO_COPY_END);
      out->ortho_kernel_ms = ms(EV_ORTHO_H2D_END, EV_ORTHO_SELECT_END) + ms(EV_ORTHO_COPY_END, EV_ORTHO_END);
      out->ortho_h2d_bytes = ctx->ortho_h2d_bytes;
    } else {
      out->ortho_h2d_ms = 0.f;
      out->ortho_kernel_ms = ms(EV_ORTHO_H2D_END, EV_ORTHO_END);
    }
    out->ortho_total_ms = ms(EV_ORTHO_BEGIN, EV_ORTHO_END);
    out->ortho_kernel_launches = ctx->ortho_launches;
  }
  return AMB_OK;
}

int amb_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return AMB_ERR_INVALID_ARGUMENT;
  cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return AMB_ERR_CUDA;
  }
  return AMB_OK;
}

int amb_host_free(void* ptr) {
  if (!ptr) return AMB_OK;
  cudaError_t e = cudaFreeHost(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return AMB_ERR_CUDA;
  }
  return AMB_OK;
}

}  // extern "C"
