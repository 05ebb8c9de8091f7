// Shared helpers for the free_hunch_b200 kernels (sm_100a).  Compiles under nvcc (product) and, with
// -DFH_EMU and tools/cuda_emu.h force-included, under g++ for CPU logic checks (tests only).
#pragma once
#ifndef FH_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#ifdef FH_EMU
#define FH_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu_dyn_smem())
#define FH_LAUNCH(kern, grid, block, smem, stream, ...)                    \
  do {                                                                     \
    emu::launch([&] { kern(__VA_ARGS__); }, grid, block, smem);            \
    ++fh_g_launches;                                                       \
  } while (0)
#else
#define FH_DYN_SMEM(type, name)                         \
  extern __shared__ __align__(16) unsigned char fh_dyn_smem_raw[]; \
  type* name = reinterpret_cast<type*>(fh_dyn_smem_raw)
// Per-kernel device timing for bench.py (fh_prof_enable): one cudaEvent pair around every launch, on the
// launching stream.  Off by default; when off the cost is one predictable branch.
extern int fh_g_prof_on;
void fh_prof_begin(const char* name, cudaStream_t stream);
void fh_prof_end(cudaStream_t stream);
#define FH_LAUNCH(kern, grid, block, smem, stream, ...)                    \
  do {                                                                     \
    if (fh_g_prof_on) fh_prof_begin(#kern, (cudaStream_t)(stream));        \
    kern<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__);      \
    if (fh_g_prof_on) fh_prof_end((cudaStream_t)(stream));                 \
    ++fh_g_launches;                                                       \
  } while (0)
#endif

extern long long fh_g_launches;

// ---- programmatic dependent launch -------------------------------------------------------------------
// The CG iteration is a chain of short dependent kernels.  Launched with the programmatic-stream-serialization
// attribute, kernel N+1 may start while kernel N is still running; it does everything that does not depend on N
// (twiddle tables, barrier init, the TMA prefetch of Z) and then blocks in fh_pdl_wait() until N has completed and
// its writes are visible.  fh_pdl_trigger() (first statement of every chained kernel) is what allows N+1 to start.
#ifdef FH_EMU
#define FH_LAUNCH_PDL FH_LAUNCH
static inline void fh_pdl_wait() {}
static inline void fh_pdl_trigger() {}
#else
#include <utility>
template <typename... KArgs, typename... Args>
static inline void fh_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}
#define FH_LAUNCH_PDL(kern, grid, block, smem, stream, ...)                              \
  do {                                                                                   \
    if (fh_g_prof_on) fh_prof_begin(#kern, (cudaStream_t)(stream));                      \
    fh_launch_pdl(kern, grid, block, smem, (cudaStream_t)(stream), __VA_ARGS__);         \
    if (fh_g_prof_on) fh_prof_end((cudaStream_t)(stream));                               \
    ++fh_g_launches;                                                                     \
  } while (0)
__device__ __forceinline__ void fh_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fh_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#define FH_MAXM 64          // max rows of Z (2 per space update): 32 BFGS pairs
#define FH_N 256            // DCT size (H = W = 256 in the DCT basis)

// Per-image scalars of a CG solve (device resident; the host only polls `done`).
struct FhCgScalars {
  double rr[2];        // r.r, ping-pong by iteration parity
  double bnorm2;       // |b|^2
  double stop2;        // (rtol*|b|)^2, or -1 when the residual test is disabled
  double resid2;       // last |r|^2
  double alpha, beta;  // (legacy path) step lengths of the current iteration
  int done;            // 1 once the loop has broken for this image   } one aligned 8-byte word: written and read
  int done_tag;        // which launch set it (+it: K1 / p kernel, -it: xr) } atomically (fh_cg.cuh: stopped-earlier test)
  int optimal;         // 1 if broken by the residual test
  int niter;           // k of cg.py:244 at the break (or iterations run so far)
};

__device__ __forceinline__ double fh_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of one double per thread; result valid in every thread.  `scratch` >= 33 doubles.
__device__ __forceinline__ double fh_block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x + 31) >> 5;
  v = fh_warp_sum(v);
  __syncthreads();                       // protect scratch reuse across calls
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = (lane < nwarp) ? scratch[lane] : 0.0;
    t = fh_warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// Sum a small array of per-block partial doubles (every thread gets the result; deterministic order).
__device__ __forceinline__ double fh_sum_partials(const double* p, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += p[i];
  return s;
}

// ---- TMA (1-D bulk copy) + mbarrier helpers ---------------------------------------------------------
// cp.async.bulk moves a contiguous, 16-byte aligned block global -> shared with one instruction issued by
// one thread; completion is signalled on an mbarrier by transaction bytes.  Under FH_EMU the copy is a
// memcpy by the issuing thread and the wait is a CTA barrier.
#ifdef FH_EMU
typedef unsigned long long fh_mbar_t;
static inline void fh_mbar_init(fh_mbar_t*, unsigned) {}
static inline void fh_mbar_fence_init() {}
static inline void fh_mbar_expect_tx(fh_mbar_t*, unsigned) {}
static inline void fh_tma_load_1d(void* dst, const void* src, unsigned bytes, fh_mbar_t*) { std::memcpy(dst, src, bytes); }
static inline void fh_mbar_wait(fh_mbar_t*, unsigned) { __syncthreads(); }
#else
typedef unsigned long long fh_mbar_t;
__device__ __forceinline__ unsigned fh_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fh_mbar_init(fh_mbar_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fh_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fh_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fh_mbar_expect_tx(fh_mbar_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fh_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fh_tma_load_1d(void* dst, const void* src, unsigned bytes, fh_mbar_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(fh_smem_u32(dst)), "l"(src), "r"(bytes), "r"(fh_smem_u32(bar)) : "memory");
}
// Bounded wait: a lost transaction traps the kernel (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void fh_mbar_wait(fh_mbar_t* bar, unsigned parity) {
  unsigned ok = 0;
  for (unsigned spins = 0; !ok; ++spins) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(fh_smem_u32(bar)), "r"(parity) : "memory");
    if (!ok && spins > 4000000u) __trap();
  }
}
#endif
