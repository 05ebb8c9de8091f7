// Microbenchmark: per-SM issue rates of DFMA, FFMA, F2F.F64.F32 and the integer float->double widening on this GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, const float* in, int iters) {
  double a0 = in[threadIdx.x], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float f0 = in[threadIdx.x + 1], f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {        // DFMA
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    } else if (MODE == 1) { // FFMA
      f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f);
      f4 = fmaf(f4, 1.0000001f, 1e-9f); f5 = fmaf(f5, 1.0000001f, 1e-9f); f6 = fmaf(f6, 1.0000001f, 1e-9f); f7 = fmaf(f7, 1.0000001f, 1e-9f);
    } else if (MODE == 2) { // F2F f32->f64 + DFMA (dependent through the float)
      a0 = fma((double)f0, m, a0); a1 = fma((double)f1, m, a1); a2 = fma((double)f2, m, a2); a3 = fma((double)f3, m, a3);
      a4 = fma((double)f4, m, a4); a5 = fma((double)f5, m, a5); a6 = fma((double)f6, m, a6); a7 = fma((double)f7, m, a7);
      f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; f4 += 1.f; f5 += 1.f; f6 += 1.f; f7 += 1.f;
    } else {                // integer widening + DFMA
#define W(f) __hiloint2double((int)(((unsigned)((int)__float_as_uint(f) >> 3) & 0x8fffffffu) + 0x38000000u), (int)(__float_as_uint(f) << 29))
      a0 = fma(W(f0), m, a0); a1 = fma(W(f1), m, a1); a2 = fma(W(f2), m, a2); a3 = fma(W(f3), m, a3);
      a4 = fma(W(f4), m, a4); a5 = fma(W(f5), m, a5); a6 = fma(W(f6), m, a6); a7 = fma(W(f7), m, a7);
      f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; f4 += 1.f; f5 += 1.f; f6 += 1.f; f7 += 1.f;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}
template <int MODE> void run(const char* name, double* out, float* in, int sms, double ghz) {
  const int iters = 20000, threads = 256, blocks = sms * 4;
  k<MODE><<<blocks, threads>>>(out, in, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * threads * iters * 8;
  printf("%-28s %8.3f ms  %7.2f Gop/s  %6.1f op/clk/SM (at %.3f GHz)\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / (sms * ghz * 1e9), ghz);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; float* in; cudaMalloc(&out, 1 << 24); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0, 1 << 16);
  const double ghz = p.clockRate * 1e-6;
  printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  run<0>("DFMA", out, in, p.multiProcessorCount, ghz);
  run<1>("FFMA", out, in, p.multiProcessorCount, ghz);
  run<2>("F2F.F64.F32 + DFMA + FADD", out, in, p.multiProcessorCount, ghz);
  run<3>("int widen + DFMA + FADD", out, in, p.multiProcessorCount, ghz);
  return 0;
}
