// host_staging.cu — PAGEABLE host memory on the caller's side of the C ABI.
//
// The reference's callers hand over std::vector<Eigen::Vector3d>, cv::Mat and Eigen::MatrixXf storage: ordinary pageable
// memory.  cudaMemcpy from / to such memory is staged by the driver through one small bounce buffer by ONE thread
// (measured on this image's hosts: ~6-12 GB/s), i.e. 100-200 ms for the 1.2 GB cloud of the benchmark, several times
// the whole GPU job.  Here the library stages itself: a process-wide pool of worker threads copies the caller's memory
// into three pinned slots in 32 MB chunks (streaming loads + stores), each chunk's DMA is enqueued as soon as it is
// complete, and the workers fill the next slot meanwhile — the PCIe copy of chunk k overlaps the memcpy of chunk k + 1,
// so the transfer runs at the host's memcpy rate (tens of GB/s) instead of the driver's single-thread rate.  The other
// direction mirrors it (DMA into a slot, workers move the slot into the caller's memory).  Pinned or registered
// memory takes the direct cudaMemcpyAsync path as before: nothing changes for callers that already pin.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "amb_context.h"

namespace amb {

namespace {

// A tiny fork-join pool: run(n, fn) executes fn(0..n-1) on the workers and the calling thread, returns when all are done.
class HostWorkers {
 public:
  static HostWorkers& get() {
    static HostWorkers* w = new HostWorkers;  // (leaked on purpose: no static-destruction order problems at exit)
    return *w;
  }
  int size() const { return static_cast<int>(threads_.size()) + 1; }
  void run(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    std::unique_lock<std::mutex> call_lock(call_mu_);  // one fork-join at a time
    {
      std::lock_guard<std::mutex> g(mu_);
      fn_ = &fn;
      next_ = 0;
      total_ = n;
      pending_ = n;
      ++generation_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> g(mu_);
    done_cv_.wait(g, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  HostWorkers() {
    unsigned int hw = std::thread::hardware_concurrency();
    int n = static_cast<int>(std::min(15u, hw > 1 ? hw - 1 : 0u));
    if (const char* e = std::getenv("AMB_STAGING_THREADS")) n = std::max(0, std::atoi(e) - 1);
    for (int k = 0; k < n; ++k) {
      try {
        threads_.emplace_back([this] { loop(); });
      } catch (...) {
        break;  // fewer workers: the calling thread always takes part
      }
    }
    for (std::thread& t : threads_) t.detach();
  }
  void work() {
    for (;;) {
      int item;
      const std::function<void(int)>* fn;
      {
        std::lock_guard<std::mutex> g(mu_);
        if (!fn_ || next_ >= total_) return;
        item = next_++;
        fn = fn_;
      }
      (*fn)(item);
      std::lock_guard<std::mutex> g(mu_);
      if (--pending_ == 0) done_cv_.notify_all();
    }
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return generation_ != seen; });
        seen = generation_;
      }
      work();
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, pending_ = 0;
  unsigned long long generation_ = 0;
  std::vector<std::thread> threads_;
};

constexpr size_t kChunk = size_t(32) << 20;
constexpr int kSlots = 3;

struct Slots {  // per device: pinned bounce buffers + the event of each slot's last DMA
  unsigned char* buf[kSlots] = {};
  cudaEvent_t ev[kSlots] = {};
  bool ok = false;
};

Slots* slots_for(int device) {
  static std::mutex mu;
  static Slots table[64];
  std::lock_guard<std::mutex> g(mu);
  Slots& s = table[device & 63];
  if (!s.ok) {
    bool good = true;
    for (int k = 0; k < kSlots && good; ++k) {
      good = cudaHostAlloc(reinterpret_cast<void**>(&s.buf[k]), kChunk, cudaHostAllocPortable) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.ev[k], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!good) {
      cudaGetLastError();
      return nullptr;
    }
    s.ok = true;
  }
  return &s;
}

std::mutex& device_mutex(int device) {  // the slots of a device serve one staged transfer at a time
  static std::mutex mu[64];
  return mu[device & 63];
}

// dst[0, bytes) = src[0, bytes) with all workers (each takes a contiguous slice; slices are multiples of 4 KB)
void parallel_copy(void* dst, const void* src, size_t bytes) {
  HostWorkers& w = HostWorkers::get();
  const int parts = static_cast<int>(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(w.size()), bytes >> 20)));
  const size_t per = (((bytes + parts - 1) / parts) + 4095) & ~static_cast<size_t>(4095);
  w.run(parts, [&](int k) {
    const size_t lo = std::min(bytes, per * static_cast<size_t>(k)), hi = std::min(bytes, lo + per);
    if (hi > lo) std::memcpy(static_cast<unsigned char*>(dst) + lo, static_cast<const unsigned char*>(src) + lo, hi - lo);
  });
}

}  // namespace

bool host_memory_is_pageable(const void* p) {
  // development switch: AMB_STAGING_OFF=1 leaves pageable memory to the driver's own staging (for A/B timing)
  static const bool off = [] {
    const char* e = std::getenv("AMB_STAGING_OFF");
    return e && e[0] == '1';
  }();
  if (off) return false;
  cudaPointerAttributes attr;
  const cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered;
}

// Host -> device on `s`.  Pinned source: one asynchronous copy.  Pageable source: staged as described above; returns
// after the last chunk's DMA has been ENQUEUED (the caller's memory has been read completely by then).
int staged_h2d(amb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return AMB_OK;
  if (!host_memory_is_pageable(src_host) || bytes < (size_t(4) << 20)) {
    AMB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, s));
    return AMB_OK;
  }
  Slots* sl = slots_for(ctx->device);
  if (!sl) {  // no pinned memory to be had: the driver's own staging
    AMB_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, s));
    return AMB_OK;
  }
  std::lock_guard<std::mutex> g(device_mutex(ctx->device));
  size_t off = 0;
  for (int k = 0; off < bytes; ++k) {
    const int slot = k % kSlots;
    const size_t len = std::min(kChunk, bytes - off);
    AMB_CUDA(ctx, cudaEventSynchronize(sl->ev[slot]));  // the slot's previous DMA has left it
    parallel_copy(sl->buf[slot], static_cast<const unsigned char*>(src_host) + off, len);
    AMB_CUDA(ctx, cudaMemcpyAsync(static_cast<unsigned char*>(dst_dev) + off, sl->buf[slot], len, cudaMemcpyHostToDevice, s));
    AMB_CUDA(ctx, cudaEventRecord(sl->ev[slot], s));
    off += len;
  }
  return AMB_OK;
}

// Device -> host on `s`; returns when the caller's memory holds the data (synchronous, like cudaMemcpy to pageable memory).
int staged_d2h(amb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return AMB_OK;
  Slots* sl = (host_memory_is_pageable(dst_host) && bytes >= (size_t(4) << 20)) ? slots_for(ctx->device) : nullptr;
  if (!sl) {
    AMB_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, s));
    AMB_CUDA(ctx, cudaStreamSynchronize(s));
    return AMB_OK;
  }
  std::lock_guard<std::mutex> g(device_mutex(ctx->device));
  const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
  // DMA of chunk k + 1 (and k + 2) is in flight while chunk k is moved out of its slot
  for (int k = 0; k < kSlots; ++k) AMB_CUDA(ctx, cudaEventSynchronize(sl->ev[k]));
  auto enqueue = [&](size_t k) -> cudaError_t {
    const size_t off = k * kChunk, len = std::min(kChunk, bytes - off);
    cudaError_t e = cudaMemcpyAsync(sl->buf[k % kSlots], static_cast<const unsigned char*>(src_dev) + off, len,
                                    cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaEventRecord(sl->ev[k % kSlots], s);
    return e;
  };
  for (size_t k = 0; k < std::min<size_t>(kSlots - 1, n_chunks); ++k) AMB_CUDA(ctx, enqueue(k));
  for (size_t k = 0; k < n_chunks; ++k) {
    if (k + kSlots - 1 < n_chunks) AMB_CUDA(ctx, enqueue(k + kSlots - 1));   // its slot was emptied in iteration k - 1
    const size_t off = k * kChunk, len = std::min(kChunk, bytes - off);
    AMB_CUDA(ctx, cudaEventSynchronize(sl->ev[k % kSlots]));
    parallel_copy(static_cast<unsigned char*>(dst_host) + off, sl->buf[k % kSlots], len);
  }
  return AMB_OK;
}

// Rows [0, h) of `w` bytes each, source pitch `src_pitch`, to a device region with pitch `dst_pitch` (host frames of the
// orthomosaic: the winners' sub-rectangles).  Pageable sources are packed by the workers into pinned slots with the
// DEVICE pitch, so each slot leaves as one linear copy.
int staged_h2d_2d(amb_ctx* ctx, void* dst_dev, size_t dst_pitch, const void* src_host, size_t src_pitch, size_t w,
                  size_t h, bool pageable, cudaStream_t s) {
  if (w == 0 || h == 0) return AMB_OK;
  Slots* sl = (pageable && dst_pitch <= kChunk) ? slots_for(ctx->device) : nullptr;
  if (!sl) {
    AMB_CUDA(ctx, cudaMemcpy2DAsync(dst_dev, dst_pitch, src_host, src_pitch, w, h, cudaMemcpyHostToDevice, s));
    return AMB_OK;
  }
  std::lock_guard<std::mutex> g(device_mutex(ctx->device));
  const size_t rows_per = std::max<size_t>(1, kChunk / dst_pitch);
  int k = 0;
  for (size_t r0 = 0; r0 < h; r0 += rows_per, ++k) {
    const int slot = k % kSlots;
    const size_t rows = std::min(rows_per, h - r0);
    AMB_CUDA(ctx, cudaEventSynchronize(sl->ev[slot]));
    unsigned char* stage = sl->buf[slot];
    const unsigned char* src = static_cast<const unsigned char*>(src_host) + r0 * src_pitch;
    HostWorkers& wk = HostWorkers::get();
    const int parts = static_cast<int>(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(wk.size()), rows / 64)));
    const size_t per = (rows + parts - 1) / parts;
    wk.run(parts, [&](int p) {
      const size_t lo = std::min(rows, per * static_cast<size_t>(p)), hi = std::min(rows, lo + per);
      for (size_t r = lo; r < hi; ++r) std::memcpy(stage + r * dst_pitch, src + r * src_pitch, w);
    });
    AMB_CUDA(ctx, cudaMemcpyAsync(static_cast<unsigned char*>(dst_dev) + r0 * dst_pitch, stage, rows * dst_pitch,
                                  cudaMemcpyHostToDevice, s));
    AMB_CUDA(ctx, cudaEventRecord(sl->ev[slot], s));
  }
  return AMB_OK;
}

}  // namespace amb
