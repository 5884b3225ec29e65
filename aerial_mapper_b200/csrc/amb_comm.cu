// amb_comm.cu — the one exchange step of the path inside the library: a context can join an NCCL
// communicator (one process per GPU, or several contexts of one process), and amb_dsm_process_sharded_device runs
//     border-halo compaction  ->  ONE ncclAllGather of the halos  ->  binning over [own points | neighbours' halos]
// entirely on the context's stream: no host synchronisation, no framework plumbing, device-side counts.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): the library keeps depending on the CUDA runtime only, a process
// that already carries an NCCL (e.g. PyTorch's bundled one) shares it, and single-GPU users never need it.
#include <dlfcn.h>

#include <algorithm>
#include <mutex>

#include "amb_context.h"
#include "dsm_plan.h"

namespace amb {

// ---- the few NCCL entry points used, declared locally (ABI-stable since NCCL 2.0) ----
struct NcclUniqueId {
  char internal[128];
};
typedef void* nccl_comm_t;
typedef int (*nccl_get_unique_id_fn)(NcclUniqueId*);
typedef int (*nccl_comm_init_rank_fn)(nccl_comm_t*, int, NcclUniqueId, int);
typedef int (*nccl_comm_destroy_fn)(nccl_comm_t);
typedef int (*nccl_all_gather_fn)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef int (*nccl_send_fn)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*nccl_recv_fn)(void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*nccl_group_fn)(void);
typedef const char* (*nccl_error_string_fn)(int);
constexpr int kNcclInt8 = 0;  // ncclInt8 / ncclChar

struct NcclApi {
  void* handle = nullptr;
  nccl_get_unique_id_fn get_unique_id = nullptr;
  nccl_comm_init_rank_fn comm_init_rank = nullptr;
  nccl_comm_destroy_fn comm_destroy = nullptr;
  nccl_all_gather_fn all_gather = nullptr;
  nccl_send_fn send = nullptr;
  nccl_recv_fn recv = nullptr;
  nccl_group_fn group_start = nullptr, group_end = nullptr;
  nccl_error_string_fn error_string = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
#ifndef AMB_CUDA_EMU
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    api.get_unique_id = reinterpret_cast<nccl_get_unique_id_fn>(dlsym(api.handle, "ncclGetUniqueId"));
    api.comm_init_rank = reinterpret_cast<nccl_comm_init_rank_fn>(dlsym(api.handle, "ncclCommInitRank"));
    api.comm_destroy = reinterpret_cast<nccl_comm_destroy_fn>(dlsym(api.handle, "ncclCommDestroy"));
    api.all_gather = reinterpret_cast<nccl_all_gather_fn>(dlsym(api.handle, "ncclAllGather"));
    api.send = reinterpret_cast<nccl_send_fn>(dlsym(api.handle, "ncclSend"));
    api.recv = reinterpret_cast<nccl_recv_fn>(dlsym(api.handle, "ncclRecv"));
    api.group_start = reinterpret_cast<nccl_group_fn>(dlsym(api.handle, "ncclGroupStart"));
    api.group_end = reinterpret_cast<nccl_group_fn>(dlsym(api.handle, "ncclGroupEnd"));
    api.error_string = reinterpret_cast<nccl_error_string_fn>(dlsym(api.handle, "ncclGetErrorString"));
    api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_gather && api.group_start &&
             api.group_end && api.send && api.recv;
#endif
  });
  return api;
}

static int nccl_fail(amb_ctx* ctx, int rc, const char* what) {
  if (ctx) {
    const char* msg = nccl().error_string ? nccl().error_string(rc) : "?";
    ctx->last_error = std::string(what) + ": NCCL error " + std::to_string(rc) + " (" + msg + ")";
  }
  return AMB_ERR_CUDA;
}

// Border halo of a sharded cloud as 32-byte records {x, y, z, id} behind a 32-byte header {count, ...}: exactly what the
// binning kernels of every rank read after the all-gather.  Unordered append (warp-aggregated atomic): order does not
// matter, every point carries its global id.
// seg_up / seg_down: two segments (header + records).  A point near the stripe's LOW-column border (large y) goes to
// seg_up — what the previous rank needs — one near the HIGH-column border to seg_down; with `split` = false everything
// goes to seg_up (the all-gather exchange delivers one list per rank to everybody).
__device__ __forceinline__ void halo_append(bool take, int lane, unsigned char* seg, unsigned int capacity, double x,
                                            double y, double z, unsigned long long id) {
  const unsigned int mask = __ballot_sync(0xffffffffu, take);
  if (!mask) return;
  const int leader = __ffs(mask) - 1;
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(reinterpret_cast<unsigned int*>(seg), static_cast<unsigned int>(__popc(mask)));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (take) {
    const unsigned int slot = base + __popc(mask & ((1u << lane) - 1u));
    if (slot < capacity) {
      double* r = reinterpret_cast<double*>(seg + 32) + 4 * static_cast<size_t>(slot);
      r[0] = x;
      r[1] = y;
      r[2] = z;
      r[3] = __longlong_as_double(static_cast<long long>(id));
    }
  }
}

__global__ void __launch_bounds__(256) dsm_halo_records_kernel(const double* __restrict__ xyz,
                                                               const unsigned long long* __restrict__ ids, size_t n,
                                                               double y_lo, double y_hi, double reach, double shift_y,
                                                               unsigned char* seg_up, unsigned char* seg_down, bool split,
                                                               unsigned int capacity) {
  const int lane = threadIdx.x & 31;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t n_round = ((n + stride - 1) / stride) * stride;  // whole warps stay converged for the ballots
  for (size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; t < n_round; t += stride) {
    bool up = false, down = false;
    double x = 0, y = 0, z = 0;
    unsigned long long id = 0;
    if (t < n) {
      y = xyz[3 * t + 1];
      const double ys = y - shift_y;
      up = ys > y_hi - reach;    // column 0 is the max-y side: the previous rank's stripe lies beyond y_hi
      down = ys < y_lo + reach;
      if (up || down) {
        x = xyz[3 * t + 0];
        z = xyz[3 * t + 2];
        id = ids ? ids[t] : static_cast<unsigned long long>(t);
      }
    }
    if (split) {
      halo_append(up, lane, seg_up, capacity, x, y, z, id);
      halo_append(down, lane, seg_down, capacity, x, y, z, id);
    } else {
      halo_append(up || down, lane, seg_up, capacity, x, y, z, id);
    }
  }
}

}  // namespace amb

using namespace amb;

extern "C" {

int amb_comm_unique_id(void* id128) {
  if (!id128) return AMB_ERR_INVALID_ARGUMENT;
  if (!nccl().ok) return AMB_ERR_UNSUPPORTED;
  NcclUniqueId id;
  const int rc = nccl().get_unique_id(&id);
  if (rc != 0) return AMB_ERR_CUDA;
  std::memcpy(id128, &id, sizeof(id));
  return AMB_OK;
}

int amb_comm_init(amb_ctx* ctx, int rank, int nranks, const void* id128) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return AMB_ERR_INVALID_ARGUMENT;
  if (ctx->nccl_comm) return AMB_ERR_INVALID_ARGUMENT;  // already a member
  if (!nccl().ok) {
    ctx->last_error = "libnccl.so.2 could not be loaded (dlopen)";
    return AMB_ERR_UNSUPPORTED;
  }
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  nccl_comm_t comm = nullptr;
  const int rc = nccl().comm_init_rank(&comm, nranks, id, rank);
  if (rc != 0) return nccl_fail(ctx, rc, "ncclCommInitRank");
  ctx->nccl_comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_size = nranks;
  // every member learns every stripe (decides, identically on all ranks, whether the halo step may talk to the two
  // neighbours only)
  {
    DeviceBuffer tmp;
    AMB_CUDA(ctx, tmp.reserve(2 * sizeof(int32_t) * (static_cast<size_t>(nranks) + 1)));
    int32_t mine[2] = {ctx->col_begin, ctx->col_end};
    int32_t* d = tmp.as<int32_t>();
    AMB_CUDA(ctx, cudaMemcpyAsync(d, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
    const int rc2 = nccl().all_gather(d, d + 2, sizeof(mine), kNcclInt8, comm, ctx->stream);
    if (rc2 != 0) {
      tmp.release();
      return nccl_fail(ctx, rc2, "ncclAllGather(stripes)");
    }
    ctx->comm_stripes.assign(2 * static_cast<size_t>(nranks), 0);
    AMB_CUDA(ctx, cudaMemcpyAsync(ctx->comm_stripes.data(), d + 2, 2 * sizeof(int32_t) * nranks, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    AMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    tmp.release();
  }
  return AMB_OK;
}

int amb_comm_set_exchange(amb_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return AMB_ERR_INVALID_ARGUMENT;
  ctx->halo_exchange_mode = mode;
  return AMB_OK;
}

int amb_comm_destroy(amb_ctx* ctx) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  if (ctx->nccl_comm && nccl().ok) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl().comm_destroy(static_cast<nccl_comm_t>(ctx->nccl_comm));
  }
  ctx->nccl_comm = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_size = 1;
  return AMB_OK;
}

int amb_comm_size(const amb_ctx* ctx) { return ctx ? ctx->comm_size : 0; }
int amb_comm_rank(const amb_ctx* ctx) { return ctx ? ctx->comm_rank : -1; }

int amb_dsm_process_sharded_device(amb_ctx* ctx, const double* d_xyz, const uint64_t* d_ids, size_t n_local,
                                   int32_t interpolation_radius, double center_easting, double center_northing,
                                   uint32_t halo_capacity) {
  if (!ctx) return AMB_ERR_INVALID_ARGUMENT;
  if (!d_xyz || !d_ids) return AMB_ERR_INVALID_ARGUMENT;  // a sharded cloud needs global ids
  if (interpolation_radius < 1) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->comm_size <= 1 || !ctx->nccl_comm)  // not sharded: the whole cloud is here
    return amb_dsm_process_device_ids(ctx, d_xyz, d_ids, n_local, interpolation_radius, center_easting,
                                      center_northing);
  if (halo_capacity == 0) return AMB_ERR_INVALID_ARGUMENT;
  const int nranks = ctx->comm_size, rank = ctx->comm_rank;
  const size_t seg_bytes = 32 * (static_cast<size_t>(halo_capacity) + 1);  // header + records
  cudaStream_t s = ctx->stream;
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_BEGIN], s));
  double y_lo = 0, y_hi = 0;
  int st = amb_stripe_y_interval(&ctx->geom, ctx->col_begin, ctx->col_end, &y_lo, &y_hi);
  if (st != AMB_OK) return st;
  const double reach = amb_dsm_halo_reach(&ctx->geom, interpolation_radius);

  // Who needs this rank's border points?  If the stripes are consecutive (rank r + 1 starts where rank r ends) and every
  // stripe is at least `reach` wide, only the two adjacent ranks do: two ncclSend / ncclRecv pairs in one group move each
  // half of the halo to the one rank that stages it.  Otherwise (narrow or irregular stripes) ONE ncclAllGather gives
  // every rank every halo.  The decision uses the stripes gathered at amb_comm_init: identical on all ranks.
  bool neighbours = ctx->halo_exchange_mode != 1;
  if (ctx->halo_exchange_mode == 0) {
    const double reach_cols = reach / ctx->geom.resolution;
    for (int r = 0; r < nranks && neighbours; ++r) {
      const int32_t b = ctx->comm_stripes[2 * r], e = ctx->comm_stripes[2 * r + 1];
      if (static_cast<double>(e - b) < reach_cols + 1.0) neighbours = false;
      if (r > 0 && ctx->comm_stripes[2 * r - 1] != b) neighbours = false;
    }
  }
  HaloSource halo;
  halo.capacity = halo_capacity;
  halo.seg_bytes = seg_bytes;
  if (neighbours) {
    // send: [up | down], recv: [from rank - 1 | from rank + 1]; each list holds ONE side of the halo: half the capacity
    const uint32_t side_capacity = halo_capacity / 2 + 1;
    const size_t side_bytes = 32 * (static_cast<size_t>(side_capacity) + 1);
    halo.capacity = side_capacity;
    halo.seg_bytes = side_bytes;
    AMB_CUDA(ctx, ctx->halo_send.reserve(2 * side_bytes));
    AMB_CUDA(ctx, ctx->halo_recv.reserve(2 * side_bytes));
    unsigned char* send = ctx->halo_send.as<unsigned char>();
    unsigned char* recv = ctx->halo_recv.as<unsigned char>();
    AMB_CUDA(ctx, cudaMemsetAsync(send, 0, 32, s));
    AMB_CUDA(ctx, cudaMemsetAsync(send + side_bytes, 0, 32, s));
    AMB_CUDA(ctx, cudaMemsetAsync(recv, 0, 32, s));              // a missing neighbour contributes an empty list
    AMB_CUDA(ctx, cudaMemsetAsync(recv + side_bytes, 0, 32, s));
    if (n_local > 0) {
      dsm_halo_records_kernel<<<kNumSMsB200 * 8, 256, 0, s>>>(
          d_xyz, reinterpret_cast<const unsigned long long*>(d_ids), n_local, y_lo, y_hi, reach, center_easting, send,
          send + side_bytes, true, side_capacity);
      AMB_CUDA(ctx, cudaGetLastError());
    }
    nccl_comm_t comm = static_cast<nccl_comm_t>(ctx->nccl_comm);
    int rc = nccl().group_start();
    if (rc == 0 && rank > 0) rc = nccl().send(send, side_bytes, kNcclInt8, rank - 1, comm, s);
    if (rc == 0 && rank > 0) rc = nccl().recv(recv, side_bytes, kNcclInt8, rank - 1, comm, s);
    if (rc == 0 && rank < nranks - 1) rc = nccl().send(send + side_bytes, side_bytes, kNcclInt8, rank + 1, comm, s);
    if (rc == 0 && rank < nranks - 1) rc = nccl().recv(recv + side_bytes, side_bytes, kNcclInt8, rank + 1, comm, s);
    const int rc_end = nccl().group_end();
    if (rc != 0 || rc_end != 0) return nccl_fail(ctx, rc != 0 ? rc : rc_end, "ncclSend/ncclRecv (halo)");
    halo.gathered = recv;
    halo.nranks = 2;
    halo.my_rank = -1;  // both segments are foreign
  } else {
    AMB_CUDA(ctx, ctx->halo_send.reserve(seg_bytes));
    AMB_CUDA(ctx, ctx->halo_recv.reserve(seg_bytes * nranks));
    unsigned char* send = ctx->halo_send.as<unsigned char>();
    AMB_CUDA(ctx, cudaMemsetAsync(send, 0, 32, s));
    if (n_local > 0) {
      dsm_halo_records_kernel<<<kNumSMsB200 * 8, 256, 0, s>>>(
          d_xyz, reinterpret_cast<const unsigned long long*>(d_ids), n_local, y_lo, y_hi, reach, center_easting, send,
          send, false, halo_capacity);
      AMB_CUDA(ctx, cudaGetLastError());
    }
    // the one collective of the step
    const int rc = nccl().all_gather(send, ctx->halo_recv.ptr, seg_bytes, kNcclInt8,
                                     static_cast<nccl_comm_t>(ctx->nccl_comm), s);
    if (rc != 0) return nccl_fail(ctx, rc, "ncclAllGather");
    halo.gathered = ctx->halo_recv.as<unsigned char>();
    halo.nranks = nranks;
    halo.my_rank = rank;
  }
  AMB_CUDA(ctx, cudaEventRecord(ctx->events[EV_DSM_H2D_END], s));  // "h2d" slot of the timings = the halo step
  ctx->dsm_had_h2d = true;
  st = dsm_run(ctx, d_xyz, reinterpret_cast<const unsigned long long*>(d_ids), n_local, interpolation_radius,
               center_easting, center_northing, 0, nullptr, &halo);
  ctx->dsm_timed = (st == AMB_OK);
  return st;
}

int amb_dsm_process_sharded(amb_ctx* ctx, const double* xyz, const uint64_t* ids, size_t n_local,
                            int32_t interpolation_radius, double center_easting, double center_northing,
                            uint32_t halo_capacity) {
  if (!ctx || !xyz || !ids) return AMB_ERR_INVALID_ARGUMENT;
  AMB_CUDA(ctx, cudaSetDevice(ctx->device));
  AMB_CUDA(ctx, ctx->points.reserve(std::max<size_t>(n_local, 1) * 3 * sizeof(double)));
  AMB_CUDA(ctx, ctx->point_ids.reserve(std::max<size_t>(n_local, 1) * sizeof(uint64_t)));
  if (n_local) {
    AMB_CUDA(ctx, cudaMemcpyAsync(ctx->points.ptr, xyz, n_local * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AMB_CUDA(ctx, cudaMemcpyAsync(ctx->point_ids.ptr, ids, n_local * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                  ctx->stream));
  }
  const int st = amb_dsm_process_sharded_device(ctx, ctx->points.as<double>(), ctx->point_ids.as<uint64_t>(), n_local,
                                                interpolation_radius, center_easting, center_northing, halo_capacity);
  if (st != AMB_OK) return st;
  AMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host entry points are synchronous (elevation final on the device)
  return AMB_OK;
}

}  // extern "C"
