// Microbenchmark: dependent-issue latency of DFMA / DADD / DMUL on this GPU and how many independent chains one warp (or two
// warps of the same SM sub-partition) need to reach the FP64 issue rate.  One CTA per launch, clock64 around the loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int K>
__global__ void k(double* out, long long* cyc, const double* in, int iters) {
  double a[K];
#pragma unroll
  for (int i = 0; i < K; ++i) a[i] = in[threadIdx.x + i];
  const double m = in[64], c = in[65];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < K; ++i) {
        if (OP == 0) a[i] = fma(a[i], m, c);
        else if (OP == 1) a[i] = a[i] + c;
        else a[i] = a[i] * m;
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) s += a[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP, int K> void run(const char* name, int threads, double* out, long long* cyc, double* in) {
  const int iters = 2000;
  k<OP, K><<<1, threads>>>(out, cyc, in, 10);
  k<OP, K><<<1, threads>>>(out, cyc, in, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-5s chains/thread %d  warps %d : %6.2f clk per dependent step (%d ops per step per warp), %5.2f clk per warp-instruction\n", name, K, threads / 32,
         (double)h / (iters * 8.0), K, (double)h / (iters * 8.0 * K));
}
int main() {
  double* out; long long* cyc; double* in;
  cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 64); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0, 1 << 16);
  for (int threads : {32, 160, 256}) {   // 1 warp; 5 warps (warps 0 and 4 share a sub-partition); 8 warps (2 per sub-partition)
    run<0, 1>("DFMA", threads, out, cyc, in); run<0, 2>("DFMA", threads, out, cyc, in); run<0, 4>("DFMA", threads, out, cyc, in);
    run<0, 8>("DFMA", threads, out, cyc, in); run<0, 16>("DFMA", threads, out, cyc, in);
    run<1, 1>("DADD", threads, out, cyc, in); run<1, 8>("DADD", threads, out, cyc, in);
    run<2, 1>("DMUL", threads, out, cyc, in); run<2, 8>("DMUL", threads, out, cyc, in);
  }
  return 0;
}
