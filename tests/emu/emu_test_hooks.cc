// emu_test_hooks.cc — TEST INFRASTRUCTURE ONLY (tests/emu): entry points of the EMULATED library that reach product
// internals no C-ABI call exposes on a single process.  Compiled into libamb_emu.so only; the product library does not
// contain this file.
//
// amb_emu_self_push: the producer side of the peer-push halo exchange (dsm_partition_kernel<PUSH>, halo_publish;
// csrc/dsm_partition.inc, csrc/halo_push.h) needs two other ranks' memory on real hardware.  Here the "neighbours'
// segments" are two local buffers: the DSM of a stripe runs through amb::dsm_run with a HaloPush whose segments point
// at them (and an empty incoming halo), and the caller gets the segments back to compare with the border sets it
// computes itself.
#include <cstring>

#include "amb_context.h"
#include "halo_push.h"

extern "C" int amb_emu_self_push(amb_ctx* ctx, const double* xyz, const unsigned long long* ids, size_t n,
                                 int32_t interpolation_radius, double center_easting, double center_northing,
                                 unsigned int capacity, int have_prev, int have_next, unsigned int stamp,
                                 unsigned char* seg_up_out, unsigned char* seg_down_out) {
  if (!ctx || !xyz || !ids || !seg_up_out || !seg_down_out) return AMB_ERR_INVALID_ARGUMENT;
  const size_t seg_bytes = 32 * (static_cast<size_t>(capacity) + 1);
  std::memset(seg_up_out, 0, seg_bytes);
  std::memset(seg_down_out, 0, seg_bytes);
  // the incoming halo: two empty segments (count 0)
  unsigned char* incoming = nullptr;
  unsigned int* counters = nullptr;
  if (cudaMalloc(&incoming, 2 * seg_bytes) != cudaSuccess || cudaMalloc(&counters, 4 * sizeof(unsigned int)) != cudaSuccess)
    return AMB_ERR_CUDA;
  std::memset(incoming, 0, 2 * seg_bytes);
  std::memset(counters, 0, 4 * sizeof(unsigned int));
  double y_lo = 0, y_hi = 0;
  int st = amb_stripe_y_interval(&ctx->geom, ctx->col_begin, ctx->col_end, &y_lo, &y_hi);
  if (st != AMB_OK) return st;
  amb::HaloSource halo;
  halo.gathered = incoming;
  halo.nranks = 2;
  halo.my_rank = -1;
  halo.capacity = capacity;
  halo.seg_bytes = seg_bytes;
  amb::HaloPush push;
  push.seg_up = have_prev ? seg_up_out : nullptr;
  push.seg_down = have_next ? seg_down_out : nullptr;
  push.counters = counters;
  push.capacity = capacity;
  push.stamp = stamp;
  push.y_lo = y_lo;
  push.y_hi = y_hi;
  push.reach = amb_dsm_halo_reach(&ctx->geom, interpolation_radius);
  push.shift_y = center_easting;
  push.wait_prev = nullptr;  // nobody to wait for: the incoming segments are final (empty)
  push.wait_next = nullptr;
  st = amb::dsm_run(ctx, xyz, ids, n, interpolation_radius, center_easting, center_northing, 0, nullptr, &halo, &push);
  const unsigned int left[3] = {counters[0], counters[1], counters[2]};
  cudaFree(incoming);
  cudaFree(counters);
  if (st != AMB_OK) return st;
  return (left[0] | left[1] | left[2]) ? AMB_ERR_CHECK_FAILED : AMB_OK;  // the last block resets the three counters
}
