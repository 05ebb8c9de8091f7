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
#define FH_LAUNCH(kern, grid, block, smem, stream, ...)                    \
  do {                                                                     \
    kern<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__);      \
    ++fh_g_launches;                                                       \
  } while (0)
#endif

extern long long fh_g_launches;

#define FH_MAXM 64          // max rows of Z (2 per space update): 32 BFGS pairs
#define FH_N 256            // DCT size (H = W = 256 in the DCT basis)

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
