// cuda_runtime.h — a CPU stand-in for the CUDA runtime and the device-side language features the product sources use.
// TEST INFRASTRUCTURE ONLY (tests/emu/README.md): lets tests/emu/build_emu.py compile aerial_mapper_b200/csrc/*.cu as
// plain C++ into tests/emu/_build/libamb_emu.so, so that the KERNEL SOURCE can be exercised on a box without a GPU.
// The product package never loads that library and the bench never does; a test has to swap it in explicitly.
//
// Execution model: one OS thread; a launch runs its blocks one after the other; the threads of a block are fibers
// (a minimal x86-64 context switch) scheduled round-robin, each running until it reaches a block barrier or a warp collective.
//   __syncthreads / __syncthreads_count     all live threads of the block
//   __shfl*_sync, __ballot_sync, __any_sync, __syncwarp   all live lanes of the warp (masks are assumed full)
// Threads that have returned no longer count.  A round in which no fiber makes progress is reported as a deadlock.
// Atomics are plain read-modify-writes (no concurrency).  "Device" memory is host memory; streams and events are
// ordering no-ops because every operation completes before the call returns.
#ifndef AMB_EMU_CUDA_RUNTIME_H_
#define AMB_EMU_CUDA_RUNTIME_H_

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>

#define AMB_CUDA_EMU 1

// ---- language ----
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__
#define __constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 {
  unsigned int x, y, z;
};
struct dim3 {
  unsigned int x, y, z;
  dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 {
  double x, y;
};
struct float2 {
  float x, y;
};
struct float4 {
  float x, y, z, w;
};
struct uint4 {
  unsigned int x, y, z, w;
};
struct int2 {
  int x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
// packed FP32 pairs (sm_100 FADD2 / FFMA2): two independent IEEE operations
inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
constexpr int warpSize = 32;

// ---- runtime API ----
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr cudaError_t cudaErrorInvalidValue = 1;
typedef void* cudaStream_t;
struct EmuEvent {
  double t;
};
typedef EmuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2,
       cudaHostAllocPortable = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };
namespace emu {
// blocks handed out by cudaHostAlloc (for AMB_EMU_PAGEABLE, below)
inline std::map<uintptr_t, size_t>& pinned_blocks() {
  static std::map<uintptr_t, size_t> m;
  return m;
}
inline std::mutex& pinned_mutex() {
  static std::mutex mu;
  return mu;
}
inline bool is_pinned(const void* p) {
  std::lock_guard<std::mutex> g(pinned_mutex());
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  std::map<uintptr_t, size_t>::const_iterator it = pinned_blocks().upper_bound(a);
  if (it == pinned_blocks().begin()) return false;
  --it;
  return a < it->first + it->second;
}
}  // namespace emu
// every host pointer counts as pinned on the emulation, so the staged (pageable) copy path of host_staging.cu stays off —
// unless AMB_EMU_PAGEABLE=1: then every pointer that did not come from cudaHostAlloc / cudaMallocHost counts as pageable
// and the chunking / slot reuse / rectangle packing logic of that file runs on the CPU (tests/test_emulated_kernels.py)
inline int cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  static const bool pageable = [] {
    const char* e = std::getenv("AMB_EMU_PAGEABLE");
    return e && e[0] == '1';
  }();
  a->type = (pageable && !emu::is_pinned(p)) ? cudaMemoryTypeUnregistered : cudaMemoryTypeHost;
  return 0;
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };

namespace emu {
double now();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
unsigned char* dyn_smem();
void note(const char* key, long long value);  // test-only statistics (amb_emu_counter)
// collectives (called from fibers)
void sync_block();
int sync_block_count(int pred);
unsigned long long warp_exchange(unsigned long long v, int src_lane);  // value of lane src_lane (all lanes call)
unsigned int warp_ballot(bool pred);
void sync_warp();
}  // namespace emu

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaGetDeviceCount(int* n) {
  *n = 1;
  return cudaSuccess;
}
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
  *v = 148;
  return cudaSuccess;
}
// Fresh "device" and pinned memory is POISONED (0xA5...): code that relies on cudaMalloc returning zeros — it often
// does on a real GPU, by accident — shows up here.
template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  *p = static_cast<T*>(std::malloc(bytes ? bytes : 1));
  if (*p) std::memset(*p, 0xA5, bytes ? bytes : 1);
  return *p ? cudaSuccess : cudaErrorInvalidValue;
}
inline cudaError_t cudaFree(void* p) {
  std::free(p);
  return cudaSuccess;
}
template <typename T>
inline cudaError_t cudaHostAlloc(T** p, size_t bytes, unsigned) {
  *p = static_cast<T*>(std::malloc(bytes ? bytes : 1));
  if (*p) {
    std::memset(*p, 0xA5, bytes ? bytes : 1);
    std::lock_guard<std::mutex> g(emu::pinned_mutex());
    emu::pinned_blocks()[reinterpret_cast<uintptr_t>(*p)] = bytes ? bytes : 1;
  }
  return *p ? cudaSuccess : cudaErrorInvalidValue;
}
inline cudaError_t cudaFreeHost(void* p) {
  {
    std::lock_guard<std::mutex> g(emu::pinned_mutex());
    emu::pinned_blocks().erase(reinterpret_cast<uintptr_t>(p));
  }
  std::free(p);
  return cudaSuccess;
}
template <typename T>
inline cudaError_t cudaHostGetDevicePointer(T** dev, void* host, unsigned) {
  *dev = static_cast<T*>(host);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height,
                                     cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t r = 0; r < height; ++r)
    std::memmove(static_cast<char*>(d) + r * dpitch, static_cast<const char*>(s) + r * spitch, width);
  return cudaSuccess;
}
template <typename T>
inline cudaError_t cudaMemcpyToSymbolAsync(T& symbol, const void* s, size_t n, size_t offset, cudaMemcpyKind,
                                           cudaStream_t = nullptr) {
  std::memcpy(reinterpret_cast<char*>(&symbol) + offset, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
  std::memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) {
  std::memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = std::malloc(1);
  return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
  std::free(s);
  return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
  *e = new EmuEvent{0.0};
  return cudaSuccess;
}
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
  if (e) e->t = emu::now();
  return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = static_cast<float>((b->t - a->t) * 1e3);
  return cudaSuccess;
}
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
  return cudaSuccess;
}

// ---- device intrinsics ----
using std::isinf;
using std::isnan;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
inline size_t max(size_t a, size_t b) { return a > b ? a : b; }

template <typename T>
inline T __ldg(const T* p) {
  return *p;
}
inline double __dadd_rn(double a, double b) { return a + b; }  // build with -ffp-contract=off
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __double2float_rn(double a) { return static_cast<float>(a); }
inline int __popc(unsigned int v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline double __longlong_as_double(long long v) {
  double d;
  std::memcpy(&d, &v, 8);
  return d;
}
inline long long __double_as_longlong(double d) {
  long long v;
  std::memcpy(&v, &d, 8);
  return v;
}
inline unsigned int __float_as_uint(float f) {
  unsigned int v;
  std::memcpy(&v, &f, 4);
  return v;
}
inline float __uint_as_float(unsigned int v) {
  float f;
  std::memcpy(&f, &v, 4);
  return f;
}
inline float __int_as_float(int v) {
  float f;
  std::memcpy(&f, &v, 4);
  return f;
}
inline void __threadfence_system() {}
inline long long clock64() { return 0; }
inline void __nanosleep(unsigned int) {}
// inter-process peer memory: not available on the emulation (the peer-push halo exchange of amb_comm.cu stays off)
struct cudaIpcMemHandle_t { char reserved[64]; };
constexpr unsigned int cudaIpcMemLazyEnablePeerAccess = 1;
constexpr cudaError_t cudaErrorPeerAccessAlreadyEnabled = 704;
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorInvalidValue; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned int) { return cudaErrorInvalidValue; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned int) { return cudaErrorInvalidValue; }
inline void __threadfence() {}

inline void __syncthreads() { emu::sync_block(); }
inline int __syncthreads_count(int pred) { return emu::sync_block_count(pred); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
inline unsigned int __ballot_sync(unsigned, bool pred) { return emu::warp_ballot(pred); }
inline bool __any_sync(unsigned, bool pred) { return emu::warp_ballot(pred) != 0u; }

namespace emu {
template <typename T>
inline T shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  bits = warp_exchange(bits, src_lane);
  T out;
  std::memcpy(&out, &bits, sizeof(T));
  return out;
}
inline int lane_id() { return static_cast<int>(threadIdx.x & 31u); }
}  // namespace emu
template <typename T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  return emu::shfl(v, src_lane & 31);
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return emu::shfl(v, emu::lane_id() ^ lane_mask);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  const int src = emu::lane_id() + static_cast<int>(delta);
  return emu::shfl(v, src < 32 ? src : emu::lane_id());
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  const int src = emu::lane_id() - static_cast<int>(delta);
  return emu::shfl(v, src >= 0 ? src : emu::lane_id());
}

template <typename T>
inline T atomicAdd(T* p, T v) {
  const T old = *p;
  *p = old + v;
  return old;
}
template <typename T>
inline T atomicExch(T* p, T v) {
  const T old = *p;
  *p = v;
  return old;
}
template <typename T>
inline T atomicMin(T* p, T v) {
  const T old = *p;
  if (v < old) *p = v;
  return old;
}
template <typename T>
inline T atomicMax(T* p, T v) {
  const T old = *p;
  if (v > old) *p = v;
  return old;
}

#endif  // AMB_EMU_CUDA_RUNTIME_H_
