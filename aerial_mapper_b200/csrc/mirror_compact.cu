// mirror_compact.cu — opt-in narrow transport of two result layers to their host mirrors (amb_set_host_mirror_compact).
//
// End to end the path is bound by PCIe: 1.6 GB of float32 result layers travel to the host per step at
// `joint_10k`.  Two of the four layers hold small integers by construction — `ortho` (gray values 0..255,
// ortho-backward-grid.cc:203-206) and `observation_index` (frame numbers, or NaN where no frame saw the cell, :182) —
// so they can cross the bus as one byte per cell and be widened to the float32 the caller's grid_map layer holds by host
// threads while later chunks are still in flight: 0.2 GB instead of 0.8 GB for the two.
//
//   pack_codes_kernel   float32 layer -> uint8 codes; any value that is not representable (a non-integer, a value
//                       outside the code range, a NaN where the layer has no NaN code) raises a flag and the layer
//                       travels as plain float32 instead — the mirror always receives exactly the layer's bits
//   chunked D2H         of the codes on the copy stream, one event per chunk
//   expander threads    wait for their chunks' events and write the float32 values into the caller's mirror
// amb_sync (and a later compact mirror of the same layer) join the threads.
#include <condition_variable>
#include <deque>
#include <exception>
#include <mutex>
#include <thread>

#include "amb_context.h"

namespace amb {
void expand_codes(const uint8_t* src, float* dst, size_t n, int nan_code);  // mirror_expand.cc
namespace {

#ifdef AMB_CUDA_EMU  // tests/emu: small chunks so that small maps exercise the multi-chunk / ragged-end logic
constexpr size_t kChunkCells = 3000;
#else
constexpr size_t kChunkCells = 4u << 20;  // 4 M cells: 4 MB of codes, 16 MB of float32
#endif
constexpr int kExpanders = 16;            // host threads widening codes (one pool per process)

// nan_code < 0: the layer has no NaN code (every byte value is a number)
__global__ void __launch_bounds__(256) pack_codes_kernel(const float* __restrict__ layer, uint8_t* __restrict__ codes,
                                                         size_t n, int nan_code, unsigned int* __restrict__ bad) {
  bool any_bad = false;
  for (size_t k = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = layer[k];
    const unsigned int bits = __float_as_uint(v);
    int code;
    if (v != v) {
      code = nan_code;
      if (nan_code < 0 || bits != 0x7fc00000u) {  // only the canonical quiet NaN the expanders write back has a code
        any_bad = true;
        code = 0;
      }
    } else {
      const int hi = nan_code >= 0 ? nan_code - 1 : 255;  // largest numeric code
      const int iv = static_cast<int>(fminf(fmaxf(v, -1.0f), 256.0f));
      code = iv;
      // exact-bits criterion: the widened code must reproduce the layer's bit pattern (rules out -0.0 as well)
      if (iv < 0 || iv > hi || __float_as_uint(static_cast<float>(iv)) != bits) {
        any_bad = true;
        code = 0;
      }
    }
    codes[k] = static_cast<uint8_t>(code);
  }
  if (any_bad) *bad = 1u;  // benign race: every writer stores the same value
}

// One expander pool per process: kExpanders threads, started on first use, fed (layer round, chunk) jobs.  A job waits for
// its chunk's copy event and widens the chunk with expand_codes (mirror_expand.cc: AVX2, streaming stores).
struct ExpandJob {
  amb_ctx* ctx;
  int layer, chunk;
  size_t lo, hi;
  int nan_code;
};

class ExpanderPool {
 public:
  static ExpanderPool& instance() {
    // intentionally never destroyed: the detached workers block on its condition variable until the process ends
    static ExpanderPool* pool = new ExpanderPool;
    return *pool;
  }
  // false if the workers could not be started (the caller then falls back to the float32 download)
  bool submit(const ExpandJob& job) {
    std::unique_lock<std::mutex> lock(mu_);
    if (!started_ && !start_locked()) return false;
    queue_.push_back(job);
    ++job.ctx->compact[job.layer].jobs_in_flight;
    lock.unlock();
    cv_.notify_one();
    return true;
  }
  void wait_layer(amb_ctx* ctx, int layer) {
    std::unique_lock<std::mutex> lock(mu_);
    done_cv_.wait(lock, [&] { return ctx->compact[layer].jobs_in_flight == 0; });
  }

 private:
  bool start_locked() {
    try {
      for (int t = 0; t < kExpanders; ++t) std::thread([this] { run(); }).detach();
    } catch (const std::exception&) {
      return false;
    }
    started_ = true;
    return true;
  }
  void run() {
    for (;;) {
      ExpandJob job;
      {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [&] { return !queue_.empty(); });
        job = queue_.front();
        queue_.pop_front();
      }
      CompactMirror& cm = job.ctx->compact[job.layer];
      cudaSetDevice(job.ctx->device);
      if (cudaEventSynchronize(cm.chunk_events[job.chunk]) != cudaSuccess) {
        cm.failed = true;
      } else {
        expand_codes(cm.host_codes + job.lo, job.ctx->host_mirror[job.layer] + job.lo, job.hi - job.lo, job.nan_code);
      }
      {
        std::lock_guard<std::mutex> lock(mu_);
        --cm.jobs_in_flight;
      }
      done_cv_.notify_all();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::deque<ExpandJob> queue_;
  bool started_ = false;
};

}  // namespace


void wait_compact_layer(amb_ctx* ctx, int layer) {
  if (ctx->compact[layer].jobs_in_flight.load() != 0) ExpanderPool::instance().wait_layer(ctx, layer);
}

int join_compact_mirrors(amb_ctx* ctx) {
  int status = AMB_OK;
  for (int l = 0; l < AMB_NUM_LAYERS; ++l) {
    CompactMirror& cm = ctx->compact[l];
    wait_compact_layer(ctx, l);
    if (cm.failed.exchange(false)) {
      // an expander gave up (its chunk event could not be waited for): the mirror is only partly widened.  Repair it with
      // a plain synchronous float32 download and tell the caller.
      status = AMB_ERR_CUDA;
      ctx->last_error = "compact mirror: an expander thread failed; layer re-downloaded as float32";
      if (ctx->host_mirror[l] && ctx->layers[l])
        cudaMemcpy(ctx->host_mirror[l], ctx->layers[l], ctx->slab_cells() * sizeof(float), cudaMemcpyDeviceToHost);
    }
  }
  return status;
}

void release_compact_mirrors(amb_ctx* ctx) {
  for (int l = 0; l < AMB_NUM_LAYERS; ++l) {
    CompactMirror& cm = ctx->compact[l];
    wait_compact_layer(ctx, l);
    cm.codes.release();
    if (cm.host_codes) cudaFreeHost(cm.host_codes);
    if (cm.host_flag) cudaFreeHost(cm.host_flag);
    for (cudaEvent_t e : cm.chunk_events)
      if (e) cudaEventDestroy(e);
    cm.chunk_events.clear();
    cm.host_codes = nullptr;
    cm.host_flag = nullptr;
  }
}

// Called by mirror_layer() for a layer whose compact transport is enabled.  Returns AMB_OK after the codes' copies and
// the expander threads have been started — or after the plain float32 download has been enqueued instead.
int mirror_layer_compact(amb_ctx* ctx, int layer) {
  CompactMirror& cm = ctx->compact[layer];
  const int nan_code = layer == AMB_LAYER_OBSERVATION_INDEX ? 255 : -1;
  const size_t cells = ctx->slab_cells();
  wait_compact_layer(ctx, layer);  // the previous round's expansion of this layer
  AMB_CUDA(ctx, cm.codes.reserve(cells));
  if (cm.host_bytes < cells) {
    if (cm.host_codes) cudaFreeHost(cm.host_codes);
    cm.host_codes = nullptr;
    AMB_CUDA(ctx, cudaHostAlloc(&cm.host_codes, cells, cudaHostAllocDefault));
    cm.host_bytes = cells;
  }
  if (!cm.host_flag) AMB_CUDA(ctx, cudaHostAlloc(&cm.host_flag, sizeof(unsigned int), cudaHostAllocMapped));
  *cm.host_flag = 0u;
  unsigned int* d_flag = nullptr;
  AMB_CUDA(ctx, cudaHostGetDevicePointer(&d_flag, cm.host_flag, 0));
  const float* d_layer = ctx->layers[layer];
  uint8_t* d_codes = cm.codes.as<uint8_t>();
  pack_codes_kernel<<<kNumSMsB200 * 8, 256, 0, ctx->stream>>>(d_layer, d_codes, cells, nan_code, d_flag);
  AMB_CUDA(ctx, cudaGetLastError());
  AMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the flag decides the transport (the layer is final here)
  if (*cm.host_flag || cm.failed.load()) {
    // not representable (e.g. `ortho` written by OrthoFromPcl, or uploaded by the caller): plain float32 transport
    return enqueue_layer_download(ctx, layer, ctx->host_mirror[layer]);
  }
  const int n_chunks = static_cast<int>((cells + kChunkCells - 1) / kChunkCells);
  while (static_cast<int>(cm.chunk_events.size()) < n_chunks) {
    cudaEvent_t e = nullptr;
    AMB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cm.chunk_events.push_back(e);
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t lo = static_cast<size_t>(c) * kChunkCells, hi = lo + kChunkCells < cells ? lo + kChunkCells : cells;
    AMB_CUDA(ctx, cudaMemcpyAsync(cm.host_codes + lo, cm.codes.as<uint8_t>() + lo, hi - lo, cudaMemcpyDeviceToHost,
                                  ctx->copy_stream));
    AMB_CUDA(ctx, cudaEventRecord(cm.chunk_events[c], ctx->copy_stream));
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t lo = static_cast<size_t>(c) * kChunkCells, hi = lo + kChunkCells < cells ? lo + kChunkCells : cells;
    if (!ExpanderPool::instance().submit(ExpandJob{ctx, layer, c, lo, hi, nan_code})) {
      // no worker threads: wait for what was queued, then the plain float32 transport (overwrites the partial result)
      wait_compact_layer(ctx, layer);
      return enqueue_layer_download(ctx, layer, ctx->host_mirror[layer]);
    }
  }
  return AMB_OK;
}

}  // namespace amb

extern "C" int amb_set_host_mirror_compact(amb_ctx* ctx, int layer, int enable) {
  if (!ctx || (layer != AMB_LAYER_ORTHO && layer != AMB_LAYER_OBSERVATION_INDEX)) return AMB_ERR_INVALID_ARGUMENT;
  ctx->compact[layer].enabled = enable != 0;
  return AMB_OK;
}
