// Minimal CPU emulation of the CUDA execution model, for index/logic checks of the kernels in
// free_hunch_b200/csrc WITHOUT a GPU (this container has none).  TEST INFRASTRUCTURE ONLY: the product
// library is built by nvcc for sm_100a and never includes this file.
//
// Model (emu::launch): blocks run one after another; the threads of a block are real std::threads that meet at
// __syncthreads()/__syncwarp() barriers; warp shuffles exchange through a per-warp mailbox.  Static
// `__shared__` variables become `static` (safe because blocks are sequential); dynamic shared memory is a
// per-launch buffer.  Device memory is host memory; cudaMalloc/cudaMemcpy are thin wrappers.
// emu::launch_coop runs ALL blocks of a grid concurrently (one BlockCtx each, small thread stacks) for kernels with
// grid barriers; such kernels must not use static __shared__ variables.
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
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };

namespace emu {
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, n = 0;
  unsigned gen = 0;
  void reset(int n_) { n = n_; count = 0; gen = 0; }
  void wait() {
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
inline void launch(const std::function<void()>& body, dim3 grid, dim3 block, size_t smem) {
  BlockCtx ctx;
  int nthreads = block.x * block.y * block.z;
  int nwarps = (nthreads + 31) / 32;
  ctx.bar.reset(nthreads);
  ctx.warp_bar = std::vector<Barrier>(nwarps);
  for (int w = 0; w < nwarps; ++w) ctx.warp_bar[w].reset(std::min(32, nthreads - 32 * w));
  ctx.mailbox.assign(nwarps * 32, 0);
  ctx.dyn_smem.assign(smem + 16, 0);
  g_blockDim = block; g_gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::vector<std::thread> ts;
        ts.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t)
          ts.emplace_back([&, t] {
            g_block = &ctx;
            t_linear = t;
            t_threadIdx = {unsigned(t % block.x), unsigned((t / block.x) % block.y), unsigned(t / (block.x * block.y))};
            t_blockIdx = {bx, by, bz};
            body();
          });
        for (auto& th : ts) th.join();
      }
}
// Cooperative launch: every block of the (1-D) grid runs at the same time; fh_grid_sync meets at g_grid_bar.
inline void launch_coop(const std::function<void()>& body, dim3 grid, dim3 block, size_t smem) {
  const int nthreads = block.x * block.y * block.z, nwarps = (nthreads + 31) / 32, nblocks = grid.x;
  std::vector<BlockCtx> ctxs(nblocks);
  for (auto& ctx : ctxs) {
    ctx.bar.reset(nthreads);
    ctx.warp_bar = std::vector<Barrier>(nwarps);
    for (int w = 0; w < nwarps; ++w) ctx.warp_bar[w].reset(std::min(32, nthreads - 32 * w));
    ctx.mailbox.assign(nwarps * 32, 0);
    ctx.dyn_smem.assign(smem + 16, 0);
  }
  g_blockDim = block; g_gridDim = grid;
  g_grid_bar.reset(nblocks);
  pthread_attr_t attr, saved;
  pthread_attr_init(&attr); pthread_attr_init(&saved);
  pthread_getattr_default_np(&saved);
  pthread_attr_setstacksize(&attr, 256 * 1024);
  pthread_attr_setguardsize(&attr, 0);         // one mapping per thread: 32768 threads stay under vm.max_map_count
  pthread_setattr_default_np(&attr);
  std::vector<std::thread> ts;
  ts.reserve((size_t)nblocks * nthreads);
  for (int b = 0; b < nblocks; ++b)
    for (int t = 0; t < nthreads; ++t)
      ts.emplace_back([&, b, t] {
        g_block = &ctxs[b];
        t_linear = t;
        t_threadIdx = {unsigned(t % block.x), unsigned((t / block.x) % block.y), unsigned(t / (block.x * block.y))};
        t_blockIdx = {unsigned(b), 0u, 0u};
        body();
      });
  for (auto& th : ts) th.join();
  pthread_setattr_default_np(&saved);
  pthread_attr_destroy(&attr); pthread_attr_destroy(&saved);
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
  thread_local int t_linear = 0;                                \
  }
