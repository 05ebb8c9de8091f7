// Minimal CPU emulation of the CUDA execution model, for index/logic checks of the kernels in
// free_hunch_b200/csrc WITHOUT a GPU (this container has none).  TEST INFRASTRUCTURE ONLY: the product
// library is built by nvcc for sm_100a and never includes this file.
//
// Model (emu::launch): blocks run one after another; the threads of a block are fibers (see below) that meet at
// __syncthreads()/__syncwarp() barriers; warp shuffles exchange through a per-warp mailbox.  Static
// `__shared__` variables become `static` (safe because blocks are sequential); dynamic shared memory is a
// per-launch buffer.  Device memory is host memory; cudaMalloc/cudaMemcpy are thin wrappers.
// emu::launch_coop runs ALL blocks of a grid "concurrently" for kernels with grid-wide synchronisation (such kernels must
// not use static __shared__ variables).  Every emulated thread is a fiber on the calling OS thread (own 64 KB stack,
// register-only context switch on x86-64, ucontext elsewhere): a 128 x 256 cooperative grid is 32 768 of them -- more
// than kernel.pid_max allows as real threads, and 10x faster than real threads meeting at futex barriers.  Barriers
// yield to a round-robin scheduler that skips fibers whose barrier has not been released; polling loops
// (fh_ll_relax) yield as well.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <pthread.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <thread>
#include <vector>

#define FH_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ { unsigned x, y, z; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };

namespace emu {
struct Barrier;
// Context switch: on x86-64 a 12-instruction stack switch (tests/emu/emu_globals.cpp: callee-saved registers only);
// elsewhere ucontext, whose swapcontext makes two signal-mask system calls per switch (20x slower).
#if defined(__x86_64__)
#define FH_EMU_FAST_SWITCH 1
extern "C" void fh_emu_switch(void** save_sp, void* load_sp);
#endif
struct Fiber {
  ucontext_t ctx;
  void* sp = nullptr;               // saved stack pointer (fast switch)
  bool done = false;
  Barrier* wait_bar = nullptr;      // barrier (and generation) the fiber sleeps on; null = runnable (or polling)
  unsigned wait_gen = 0;
  int block = 0, t = 0;
};
extern bool g_fiber_mode;           // true inside launch_coop: single OS thread, no locking needed
extern Fiber* g_cur;
extern ucontext_t g_sched;
extern void* g_sched_sp;
inline void fiber_yield() {
#ifdef FH_EMU_FAST_SWITCH
  fh_emu_switch(&g_cur->sp, g_sched_sp);
#else
  swapcontext(&g_cur->ctx, &g_sched);
#endif
}
inline void relax() {               // inside a polling loop
  if (g_fiber_mode) { g_cur->wait_bar = nullptr; fiber_yield(); }
  else std::this_thread::yield();
}
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, n = 0;
  unsigned gen = 0;
  void reset(int n_) { n = n_; count = 0; gen = 0; }
  void wait() {
    if (g_fiber_mode) {
      const unsigned g = gen;
      if (++count == n) { count = 0; ++gen; return; }
      g_cur->wait_bar = this; g_cur->wait_gen = g;
      while (g == gen) fiber_yield();
      g_cur->wait_bar = nullptr;
      return;
    }
    std::unique_lock<std::mutex> lk(m);
    unsigned g = gen;
    if (++count == n) { count = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return g != gen; });
  }
};
struct BlockCtx {
  Barrier bar;
  std::vector<Barrier> warp_bar;
  std::vector<uint64_t> mailbox;   // 32 slots per warp
  std::vector<char> dyn_smem;
};
extern thread_local uint3_ t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern thread_local BlockCtx* g_block;
extern Barrier g_grid_bar;
extern thread_local int t_linear;
extern const std::function<void()>* g_body;
inline void fiber_entry() {
  (*g_body)();
  g_cur->done = true;
#ifdef FH_EMU_FAST_SWITCH
  fh_emu_switch(&g_cur->sp, g_sched_sp);      // never resumed
  std::abort();
#endif                               // ucontext: returning resumes uc_link = the scheduler
}
// Fiber stacks: one persistent mapping, grown on demand (64 KB per emulated thread, touched pages only).
inline char* fiber_stacks(size_t nfib, size_t stack_bytes) {
  static char* pool = nullptr;
  static size_t pool_bytes = 0;
  if (nfib * stack_bytes > pool_bytes) {
    if (pool) munmap(pool, pool_bytes);
    pool_bytes = nfib * stack_bytes;
    pool = (char*)mmap(nullptr, pool_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (pool == (char*)MAP_FAILED) { std::fprintf(stderr, "emu: cannot map fiber stacks\n"); std::abort(); }
  }
  return pool;
}
// Runs the blocks `ids` (one BlockCtx each) to completion, all of them live at the same time.
inline void run_group(const std::function<void()>& body, std::vector<BlockCtx>& ctxs, const std::vector<uint3_>& ids, dim3 block, size_t smem) {
  const int nthreads = block.x * block.y * block.z, nwarps = (nthreads + 31) / 32;
  for (auto& ctx : ctxs) {
    ctx.bar.reset(nthreads);
    if ((int)ctx.warp_bar.size() != nwarps) ctx.warp_bar = std::vector<Barrier>(nwarps);
    for (int w = 0; w < nwarps; ++w) ctx.warp_bar[w].reset(std::min(32, nthreads - 32 * w));
    ctx.mailbox.assign(nwarps * 32, 0);
    ctx.dyn_smem.assign(smem + 16, 0);
  }
  const size_t nfib = ids.size() * (size_t)nthreads, stack_bytes = 64 * 1024;
  char* stacks = fiber_stacks(nfib, stack_bytes);
  std::vector<Fiber> fibers(nfib);
  const std::function<void()>* saved_body = g_body;
  g_body = &body;
  for (size_t i = 0; i < nfib; ++i) {
    Fiber& f = fibers[i];
    f.block = (int)(i / nthreads); f.t = (int)(i % nthreads);
#ifdef FH_EMU_FAST_SWITCH
    // initial frame for fh_emu_switch: six zeroed callee-saved registers, then the entry point as return address;
    // sp = 0 (mod 16) so that the entry function starts with the ABI's call-site alignment
    void** sp = (void**)(((uintptr_t)(stacks + (i + 1) * stack_bytes) & ~(uintptr_t)15) - 128);
    for (int k = 0; k < 6; ++k) sp[k] = nullptr;
    sp[6] = (void*)fiber_entry;
    sp[7] = nullptr;
    f.sp = sp;
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stacks + i * stack_bytes;
    f.ctx.uc_stack.ss_size = stack_bytes;
    f.ctx.uc_link = &g_sched;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
#endif
  }
  g_fiber_mode = true;
  size_t remaining = nfib;
  while (remaining) {
    bool progressed = false;
    for (size_t i = 0; i < nfib; ++i) {
      Fiber& f = fibers[i];
      if (f.done || (f.wait_bar && f.wait_bar->gen == f.wait_gen)) continue;
      g_cur = &f;
      g_block = &ctxs[f.block];
      t_linear = f.t;
      t_threadIdx = {unsigned(f.t % block.x), unsigned((f.t / block.x) % block.y), unsigned(f.t / (block.x * block.y))};
      t_blockIdx = ids[f.block];
#ifdef FH_EMU_FAST_SWITCH
      fh_emu_switch(&g_sched_sp, f.sp);
#else
      swapcontext(&g_sched, &f.ctx);
#endif
      if (f.done) --remaining;
      progressed = true;
    }
    if (!progressed) { std::fprintf(stderr, "emu: deadlock (%zu emulated threads left)\n", remaining); std::abort(); }
  }
  g_fiber_mode = false;
  g_cur = nullptr;
  g_block = nullptr;
  g_body = saved_body;
}
// Ordinary launch: blocks one after another (static __shared__ variables are safe).
inline void launch(const std::function<void()>& body, dim3 grid, dim3 block, size_t smem) {
  g_blockDim = block; g_gridDim = grid;
  std::vector<BlockCtx> ctxs(1);
  std::vector<uint3_> ids(1);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        ids[0] = {bx, by, bz};
        run_group(body, ctxs, ids, block, smem);
      }
}
// Cooperative launch: every block of the (1-D) grid is live at the same time; fh_grid_sync meets at g_grid_bar.
inline void launch_coop(const std::function<void()>& body, dim3 grid, dim3 block, size_t smem) {
  g_blockDim = block; g_gridDim = grid;
  g_grid_bar.reset((int)grid.x);
  std::vector<BlockCtx> ctxs(grid.x);
  std::vector<uint3_> ids(grid.x);
  for (unsigned b = 0; b < grid.x; ++b) ids[b] = {b, 0u, 0u};
  run_group(body, ctxs, ids, block, smem);
}
template <typename T>
inline T shfl_generic(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl size");
  int warp = t_linear / 32, lane = t_linear % 32;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  g_block->mailbox[warp * 32 + lane] = bits;
  g_block->warp_bar[warp].wait();
  int n = g_block->warp_bar[warp].n;
  uint64_t r = g_block->mailbox[warp * 32 + ((src_lane >= 0 && src_lane < n) ? src_lane : lane)];
  g_block->warp_bar[warp].wait();
  T out;
  std::memcpy(&out, &r, sizeof(T));
  return out;
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline void __syncthreads() { emu::g_block->bar.wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_block->warp_bar[emu::t_linear / 32].wait(); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int mask, int = 32) { return emu::shfl_generic(v, (emu::t_linear % 32) ^ mask); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) { return emu::shfl_generic(v, (emu::t_linear % 32) + d); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::shfl_generic(v, src); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
static inline double atomicAdd(double* p, double v) { static std::mutex m; std::lock_guard<std::mutex> g(m); double o = *p; *p = o + v; return o; }
static inline float atomicAdd(float* p, float v) { static std::mutex m; std::lock_guard<std::mutex> g(m); float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { static std::mutex m; std::lock_guard<std::mutex> g(m); int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { static std::mutex m; std::lock_guard<std::mutex> g(m); unsigned o = *p; *p = o + v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline char* emu_dyn_smem() { return (char*)(((uintptr_t)emu::g_block->dyn_smem.data() + 15) & ~uintptr_t(15)); }

// runtime shims
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::calloc(1, n ? n : 1); return 0; }
enum { cudaHostAllocMapped = 2 };
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::calloc(1, n ? n : 1); return 0; }
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return 0; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t = 0) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

#define FH_EMU_DEFINE_GLOBALS                                   \
  namespace emu {                                               \
  thread_local uint3_ t_threadIdx, t_blockIdx;                  \
  dim3 g_blockDim, g_gridDim;                                   \
  thread_local BlockCtx* g_block = nullptr;                     \
  Barrier g_grid_bar;                                           \
  bool g_fiber_mode = false;                                    \
  Fiber* g_cur = nullptr;                                       \
  ucontext_t g_sched;                                           \
  void* g_sched_sp = nullptr;                                   \
  const std::function<void()>* g_body = nullptr;                \
  thread_local int t_linear = 0;                                \
  }
