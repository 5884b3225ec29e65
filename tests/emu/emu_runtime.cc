// emu_runtime.cc — the fiber scheduler behind tests/emu/include/cuda_runtime.h (TEST INFRASTRUCTURE ONLY).
// Fibers switch with a dozen instructions of x86-64 assembly (callee-saved registers + stack pointer); ucontext's
// swapcontext costs a signal-mask system call per switch, and a warp-per-cell kernel switches ~10^7 times.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <string>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

namespace {

constexpr size_t kStackBytes = 256 * 1024;

struct Barrier {
  int alive = 0, arrived = 0;
  unsigned long long generation = 0;
};
struct Warp {
  Barrier bar;
  unsigned long long slot[32];
  bool pred[32];
};
struct Fiber {
  void* sp = nullptr;  // saved stack pointer while the fiber is not running
  char* stack = nullptr;
  bool done = false;
};

std::vector<Fiber> g_fibers;  // pool; the first g_n belong to the running block
int g_n = 0;
int g_sched_mode = -1;  // 0 ascending, 1 reverse, 2 pseudo-random permutation (changes every round)
unsigned long long g_sched_stride = 1, g_sched_offset = 0, g_sched_state = 0x9e3779b97f4a7c15ull;
std::vector<Warp> g_warps;
Barrier g_block;
int g_block_count_acc = 0, g_block_count_result = 0;
void* g_scheduler_sp = nullptr;
int g_current = -1;
unsigned long long g_progress = 0;  // bumped whenever a barrier releases or a fiber finishes
const std::function<void()>* g_body = nullptr;
std::vector<unsigned char> g_dyn_smem;
unsigned char* g_dyn_ptr = nullptr;

extern "C" void amb_emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl amb_emu_switch
.type amb_emu_switch,@function
amb_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size amb_emu_switch,.-amb_emu_switch
)");

void yield() { amb_emu_switch(&g_fibers[g_current].sp, g_scheduler_sp); }

void arrive_and_wait(Barrier& b) {
  const unsigned long long gen = b.generation;
  if (++b.arrived >= b.alive) {
    b.arrived = 0;
    ++b.generation;
    ++g_progress;
    return;
  }
  while (b.generation == gen) yield();
}

void leave(Barrier& b) {  // a thread that returned no longer takes part
  --b.alive;
  if (b.alive > 0 && b.arrived >= b.alive) {
    b.arrived = 0;
    ++b.generation;
    ++g_progress;
  }
}

void fiber_main() {
  (*g_body)();
  Fiber& f = g_fibers[g_current];
  f.done = true;
  ++g_progress;
  leave(g_block);
  leave(g_warps[g_current >> 5].bar);
  amb_emu_switch(&f.sp, g_scheduler_sp);  // never resumed
  std::abort();
}

void prepare(Fiber& f) {  // a fresh stack whose first switch-in "returns" into fiber_main
  uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack) + kStackBytes) & ~uintptr_t(15);
  void** sp = reinterpret_cast<void**>(top - 64);
  for (int k = 0; k < 6; ++k) sp[k] = nullptr;           // r15 r14 r13 r12 rbx rbp
  sp[6] = reinterpret_cast<void*>(&fiber_main);          // return address of the first switch
  sp[7] = nullptr;                                       // fiber_main's own (never used) return address
  f.sp = sp;
}

}  // namespace

unsigned char* dyn_smem() { return g_dyn_ptr; }

static std::map<std::string, std::pair<long long, long long> > g_notes;  // key -> (sum, count)
void note(const char* key, long long value) {
  std::pair<long long, long long>& e = g_notes[key];
  e.first += value;
  e.second += 1;
}

void sync_block() { arrive_and_wait(g_block); }

int sync_block_count(int pred) {
  g_block_count_acc += pred ? 1 : 0;
  const unsigned long long gen = g_block.generation;
  if (++g_block.arrived >= g_block.alive) {
    g_block.arrived = 0;
    g_block_count_result = g_block_count_acc;
    g_block_count_acc = 0;
    ++g_block.generation;
    ++g_progress;
  } else {
    while (g_block.generation == gen) yield();
  }
  const int r = g_block_count_result;
  arrive_and_wait(g_block);  // nobody starts the next count before everyone has read this one
  return r;
}

void sync_warp() { arrive_and_wait(g_warps[g_current >> 5].bar); }

unsigned long long warp_exchange(unsigned long long v, int src_lane) {
  Warp& w = g_warps[g_current >> 5];
  const int lane = g_current & 31;
  w.slot[lane] = v;
  arrive_and_wait(w.bar);
  const int src = (g_current & ~31) + src_lane;
  // a lane that has exited (or does not exist) has no value: CUDA leaves the result undefined; return the caller's own
  const unsigned long long out =
      (src < g_n && !g_fibers[src].done) ? w.slot[src_lane] : v;
  arrive_and_wait(w.bar);
  return out;
}

unsigned int warp_ballot(bool pred) {
  Warp& w = g_warps[g_current >> 5];
  const int lane = g_current & 31;
  w.pred[lane] = pred;
  arrive_and_wait(w.bar);
  unsigned int mask = 0;
  const int base = g_current & ~31;
  for (int l = 0; l < 32; ++l) {
    const int t = base + l;
    if (t < g_n && !g_fibers[t].done && w.pred[l]) mask |= 1u << l;
  }
  arrive_and_wait(w.bar);
  return mask;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int n = static_cast<int>(block.x * block.y * block.z);
  if (n <= 0 || grid.x == 0) return;
  if (g_sched_mode < 0) {
    const char* m = std::getenv("AMB_EMU_SCHED");
    g_sched_mode = (m && std::string(m) == "reverse") ? 1 : (m && std::string(m).rfind("random", 0) == 0) ? 2 : 0;
    if (g_sched_mode == 2 && std::strlen(m) > 7) g_sched_state ^= std::strtoull(m + 7, nullptr, 10) * 0x2545f4914f6cdd1dull;
  }
  if (static_cast<int>(g_fibers.size()) < n) {
    const size_t old = g_fibers.size();
    g_fibers.resize(n);
    for (size_t k = old; k < g_fibers.size(); ++k) g_fibers[k].stack = static_cast<char*>(std::malloc(kStackBytes));
  }
  g_n = n;
  g_dyn_smem.assign(smem + 64, 0);
  g_dyn_ptr = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(g_dyn_smem.data()) + 63) & ~uintptr_t(63));
  g_body = &body;
  blockDim = block;
  gridDim = grid;
  const int n_warps = (n + 31) / 32;
  for (unsigned int b = 0; b < grid.x; ++b) {
    g_warps.assign(n_warps, Warp());
    g_block = Barrier();
    g_block.alive = n;
    g_block_count_acc = 0;
    for (int t = 0; t < n; ++t) {
      Fiber& f = g_fibers[t];
      f.done = false;
      prepare(f);
      ++g_warps[t >> 5].bar.alive;
    }
    int remaining = n;
    while (remaining > 0) {
      const unsigned long long before = g_progress;
      remaining = 0;
      if (g_sched_mode == 2) {  // t = (slot * stride + offset) mod n with an odd stride coprime to n: a permutation
        g_sched_state = g_sched_state * 6364136223846793005ull + 1442695040888963407ull;
        g_sched_stride = ((g_sched_state >> 33) % static_cast<unsigned long long>(n)) | 1ull;
        while (std::__gcd(g_sched_stride, static_cast<unsigned long long>(n)) != 1ull) g_sched_stride += 2;
        g_sched_offset = (g_sched_state >> 13) % static_cast<unsigned long long>(n);
      }
      for (int slot = 0; slot < n; ++slot) {
        // AMB_EMU_SCHED: the order in which the threads of a block get their turn (default ascending).  A kernel must not
        // care; a missing barrier between a write and another thread's read shows up as a changed result under
        // `reverse` or `random` (tests/test_emulated_kernels.py runs the pending kernels under all three).
        int t = slot;
        if (g_sched_mode == 1) {
          t = n - 1 - slot;
        } else if (g_sched_mode == 2) {
          t = static_cast<int>((static_cast<unsigned long long>(slot) * g_sched_stride + g_sched_offset) % n);
        }
        if (g_fibers[t].done) continue;
        g_current = t;
        threadIdx = uint3{static_cast<unsigned int>(t) % block.x, (static_cast<unsigned int>(t) / block.x) % block.y,
                          static_cast<unsigned int>(t) / (block.x * block.y)};
        blockIdx = uint3{b, 0, 0};
        amb_emu_switch(&g_scheduler_sp, g_fibers[t].sp);
        if (!g_fibers[t].done) ++remaining;
      }
      if (remaining > 0 && g_progress == before) {
        std::fprintf(stderr, "cuda emulation: deadlock in block %u (%d threads wait at collectives nobody completes)\n", b,
                     remaining);
        std::abort();
      }
    }
  }
  g_current = -1;
  g_body = nullptr;
  g_n = 0;
}

}  // namespace emu

// Test-only: read and reset a statistic recorded by emu::note() inside a kernel.
extern "C" int amb_emu_counter(const char* key, long long* sum, long long* count) {
  std::pair<long long, long long>& e = emu::g_notes[key];
  *sum = e.first;
  *count = e.second;
  e = std::make_pair(0ll, 0ll);
  return 0;
}
