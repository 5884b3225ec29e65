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
#ifdef AMB_CUDA_EMU  // tests/emu only: the staged path ran (tests/test_emulated_kernels.py)
    emu::note("staged_h2d_chunks", 1);
#endif
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
#ifdef AMB_CUDA_EMU
    emu::note("staged_d2h_chunks", 1);
#endif
  }
  return AMB_OK;
}

// Host frames of the orthomosaic: rectangle k = rows [0, h) of `w` bytes at src (row pitch src_pitch) -> device region at
// dst (row pitch dst_pitch); the device regions of consecutive rectangles are ADJACENT (rect k + 1 starts where rect k
// ends), which is how ortho_run lays them out.  Pinned sources: one cudaMemcpy2DAsync each.  Pageable sources (cv::Mat
// storage): the rows of as many rectangles as fit a pinned slot are packed by the whole worker pool in ONE fork-join
// (with the device pitch, so the slot is a verbatim image of a contiguous device range) and leave as one linear copy —
// a fork-join and a DMA per 32 MB, not per 2 MB rectangle (measured at joint_10k: 125 ms with one fork-join per
// rectangle, 100 ms with the driver's own staging of 250 cudaMemcpy2DAsync calls).
int staged_h2d_rects(amb_ctx* ctx, const StagedRect* rects, size_t n, cudaStream_t s) {
  if (n == 0) return AMB_OK;
  bool pageable = false;
  for (size_t k = 0; k < n && !pageable; ++k) pageable = rects[k].h > 0 && host_memory_is_pageable(rects[k].src);
  Slots* sl = pageable ? slots_for(ctx->device) : nullptr;
  bool adjacent = true;
  for (size_t k = 0; k + 1 < n; ++k)
    adjacent = adjacent && rects[k + 1].dst == rects[k].dst + rects[k].dst_pitch * rects[k].h;
  for (size_t k = 0; k < n; ++k) adjacent = adjacent && rects[k].dst_pitch <= kChunk;
  if (!sl || !adjacent) {
    for (size_t k = 0; k < n; ++k)
      if (rects[k].h)
        AMB_CUDA(ctx, cudaMemcpy2DAsync(rects[k].dst, rects[k].dst_pitch, rects[k].src, rects[k].src_pitch, rects[k].w,
                                        rects[k].h, cudaMemcpyHostToDevice, s));
    return AMB_OK;
  }
  std::lock_guard<std::mutex> g(device_mutex(ctx->device));
  struct Seg {  // rows [r0, r0 + rows) of rectangle `k`, packed at byte offset `at` of the slot
    size_t k, r0, rows, at;
  };
  std::vector<Seg> segs;
  size_t k = 0, r0 = 0;
  int turn = 0;
  while (k < n) {
    // fill one slot
    segs.clear();
    size_t used = 0, total_rows = 0;
    unsigned char* dev_begin = rects[k].dst + r0 * rects[k].dst_pitch;
    while (k < n) {
      const StagedRect& r = rects[k];
      const size_t fit = (kChunk - used) / r.dst_pitch;
      const size_t rows = std::min(fit, r.h - r0);
      if (rows == 0 && r.h > r0) break;  // slot full
      if (rows) {
        segs.push_back(Seg{k, r0, rows, used});
        used += rows * r.dst_pitch;
        total_rows += rows;
      }
      r0 += rows;
      if (r0 >= r.h) {
        ++k;
        r0 = 0;
      } else {
        break;  // the rest of this rectangle goes to the next slot
      }
    }
    if (used == 0) continue;
    const int slot = turn++ % kSlots;
    AMB_CUDA(ctx, cudaEventSynchronize(sl->ev[slot]));
    unsigned char* stage = sl->buf[slot];
    HostWorkers& wk = HostWorkers::get();
    const int parts = static_cast<int>(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(wk.size()), total_rows / 256)));
    const size_t per = (total_rows + parts - 1) / parts;
    wk.run(parts, [&](int p) {
      size_t lo = std::min(total_rows, per * static_cast<size_t>(p)), hi = std::min(total_rows, lo + per);
      size_t first = 0;  // global row index of the segment's first row
      for (size_t q = 0; q < segs.size() && lo < hi; ++q) {
        const Seg& sg = segs[q];
        if (lo < first + sg.rows) {
          const StagedRect& r = rects[sg.k];
          const size_t a = lo - first, b = std::min(sg.rows, hi - first);
          for (size_t row = a; row < b; ++row)
            std::memcpy(stage + sg.at + row * r.dst_pitch, r.src + (sg.r0 + row) * r.src_pitch, r.w);
          lo = first + b;
        }
        first += sg.rows;
      }
    });
    AMB_CUDA(ctx, cudaMemcpyAsync(dev_begin, stage, used, cudaMemcpyHostToDevice, s));
    AMB_CUDA(ctx, cudaEventRecord(sl->ev[slot], s));
#ifdef AMB_CUDA_EMU
    emu::note("staged_rect_slots", 1);
#endif
  }
  return AMB_OK;
}

}  // namespace amb
