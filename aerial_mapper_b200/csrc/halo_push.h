// halo_push.h — device side of the peer-push halo exchange (amb_comm.cu sets it up; the binning's partition kernel,
// dsm_partition.inc, is the producer): border records are stored straight into the adjacent ranks' receive segments over
// NVLink peer memory while the rank's own points are being binned, and the last block publishes {count, step stamp}.
#ifndef AMB_HALO_PUSH_H_
#define AMB_HALO_PUSH_H_

#include <cuda_runtime.h>

namespace amb {

struct HaloPush {
  unsigned char* seg_up = nullptr;     // PEER memory: the segment of rank - 1 that this rank fills (nullptr: no such rank)
  unsigned char* seg_down = nullptr;   // PEER memory: the segment of rank + 1 that this rank fills
  unsigned int* counters = nullptr;    // LOCAL: [0] up count, [1] down count, [2] blocks done
  unsigned int capacity = 0;           // records per segment
  unsigned int stamp = 0;              // step + 1
  double y_lo = 0, y_hi = 0, reach = 0, shift_y = 0;   // a point is a border point if y - shift_y is within reach of a border
  // consumer side: this rank's own segments of the same parity (nullptr: no such neighbour)
  const unsigned char* wait_prev = nullptr;
  const unsigned char* wait_next = nullptr;
};

__device__ __forceinline__ void store_peer_record(double* dst, double x, double y, double z, unsigned long long id) {
#ifdef AMB_CUDA_EMU  // tests/emu (the exchange itself never runs there: no NCCL, no peer memory)
  dst[0] = x;
  dst[1] = y;
  dst[2] = z;
  dst[3] = __longlong_as_double(static_cast<long long>(id));
#else
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(x), "d"(y), "d"(z),
               "d"(__longlong_as_double(static_cast<long long>(id)))
               : "memory");
#endif
}
__device__ __forceinline__ void store_release_sys(unsigned char* p, unsigned long long v) {
#ifdef AMB_CUDA_EMU
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
#else
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ unsigned long long load_acquire_sys(const unsigned char* p) {
#ifdef AMB_CUDA_EMU
  return *reinterpret_cast<const volatile unsigned long long*>(p);
#else
  unsigned long long h;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(h) : "l"(p) : "memory");
  return h;
#endif
}

// Warp-aggregated append of the lanes with `take` to a peer segment (every lane of the warp must call it).
__device__ __forceinline__ void halo_push_record(bool take, int lane, unsigned char* seg, unsigned int* counter,
                                                 unsigned int capacity, double x, double y, double z,
                                                 unsigned long long id) {
  const unsigned int mask = __ballot_sync(0xffffffffu, take);
  if (!mask) return;
  const int leader = __ffs(mask) - 1;
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(counter, static_cast<unsigned int>(__popc(mask)));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (take) {
    const unsigned int slot = base + __popc(mask & ((1u << lane) - 1u));
    if (slot < capacity) store_peer_record(reinterpret_cast<double*>(seg + 32) + 4 * static_cast<size_t>(slot), x, y, z, id);
  }
}

// End of the producing kernel, called by every thread of every block: each thread's peer stores are fenced (system
// scope) before its block takes a ticket; the block that takes the last ticket has therefore observed all of them and
// releases the two headers {count | stamp << 32}.
__device__ __forceinline__ void halo_publish(const HaloPush& a) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(a.counters + 2, 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      const unsigned int c_up = atomicExch(a.counters + 0, 0u);  // (reset for the next step of this parity)
      const unsigned int c_down = atomicExch(a.counters + 1, 0u);
      atomicExch(a.counters + 2, 0u);
      __threadfence_system();
      const unsigned long long stamp = static_cast<unsigned long long>(a.stamp) << 32;
      if (a.seg_up) store_release_sys(a.seg_up, static_cast<unsigned long long>(c_up) | stamp);
      if (a.seg_down) store_release_sys(a.seg_down, static_cast<unsigned long long>(c_down) | stamp);
    }
  }
}

// amb_comm.cu: enqueue the kernel in which two lanes spin (acquire, system scope, bounded) until both of this rank's
// segments carry the step's stamp
int halo_wait_launch(struct ::amb_ctx* ctx, const HaloPush& push);

}  // namespace amb
#endif  // AMB_HALO_PUSH_H_
