// Element-wise float64 glue of one mechanism call and one sampler step, fused: what the reference (and round 1 here) spends
// ~40 separate ATen launches per call on (conditioning_mechanisms.py:243, 250, 283-288; generate_conditional.py:146-167).
// Same operations in the same order as the reference's expressions, so the float64 results are bit-identical to PyTorch's.
#pragma once
#include "fh_common.cuh"

#ifdef FH_EMU
static inline double fh_mul(double a, double b) { return a * b; }
static inline double fh_add(double a, double b) { return a + b; }
#else
// separate roundings, never contracted into an FMA: PyTorch evaluates a + b * c as two kernels
__device__ __forceinline__ double fh_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fh_add(double a, double b) { return __dadd_rn(a, b); }
#endif
#define FH_MECH_THREADS 256
#define FH_MECH_PER_THREAD 8
#define FH_MECH_CHUNK (FH_MECH_THREADS * FH_MECH_PER_THREAD)

// Per-CTA statistics of v = g * sigma^2 over its chunk of image b: (count, mean, M2) by a two-pass evaluation on registers;
// the consumer merges the chunks pairwise in a fixed order (Chan et al.), which is what makes the standard deviation of
// conditioning_mechanisms.py:283 robust where sum-of-squares minus squared-sum would cancel.
__global__ void __launch_bounds__(FH_MECH_THREADS)
fh_k_mech_stats(const double* __restrict__ g, double sigma2, double* __restrict__ part /* (B, nblk, 3) */, long d) {
  __shared__ double red[33];
  const size_t b = blockIdx.y;
  const long i0 = (long)blockIdx.x * FH_MECH_CHUNK;
  double v[FH_MECH_PER_THREAD];
  double s = 0.0;
  int n = 0;
#pragma unroll
  for (int e = 0; e < FH_MECH_PER_THREAD; ++e) {
    const long i = i0 + threadIdx.x + (long)e * FH_MECH_THREADS;
    v[e] = 0.0;
    if (i < d) { v[e] = fh_mul(g[b * d + i], sigma2); s += v[e]; ++n; }
  }
  const double tot = fh_block_sum(s, red);
  const long cnt = min((long)FH_MECH_CHUNK, d - i0);
  const double mean = tot / (double)cnt;
  double m2 = 0.0;
#pragma unroll
  for (int e = 0; e < FH_MECH_PER_THREAD; ++e) {
    const long i = i0 + threadIdx.x + (long)e * FH_MECH_THREADS;
    if (i < d) { const double t = v[e] - mean; m2 = fma(t, t, m2); }
  }
  const double M2 = fh_block_sum(m2, red);
  if (threadIdx.x == 0) {
    double* p = part + (b * gridDim.x + blockIdx.x) * 3;
    p[0] = (double)cnt; p[1] = mean; p[2] = M2;
  }
  (void)n;
}
// out = mu + (std(v) > threshold ? sigv * scale : g * scale * sigma^2)   per image (conditioning_mechanisms.py:283-288);
// flag_out[b] = the decision.  Every CTA merges the chunk statistics itself (same order, same result everywhere).
// `flag_in` != null: the decision was taken before (fh_k_mech_flag) -- then sigv is only read where the flag is set, which is
// what lets the caller skip the Sigma.mat product for the other images.
__global__ void __launch_bounds__(FH_MECH_THREADS)
fh_k_mech_combine(const double* __restrict__ mu, const double* __restrict__ g, const double* __restrict__ sigv, const double* __restrict__ part,
                  int nblk, double sigma2, double threshold, double scale, double* __restrict__ out, int* __restrict__ flag_out, long d,
                  const int* __restrict__ flag_in) {
  __shared__ int s_flag;
  __shared__ double red[33];
  const size_t b = blockIdx.y;
  if (flag_in) {
    if (threadIdx.x == 0) s_flag = flag_in[b];
  } else {
    // merge of the chunk statistics (Chan et al.): N = sum n_i, mean = sum n_i mean_i / N, M2 = sum (M2_i + n_i (mean_i - mean)^2)
    // as three fixed-order block sums (one chunk per thread; a serial merge by one thread took 32 us per launch)
    const double* p = part + b * nblk * 3;
    const bool has = (int)threadIdx.x < nblk;
    const double ni = has ? p[3 * threadIdx.x] : 0.0, mi = has ? p[3 * threadIdx.x + 1] : 0.0, Mi = has ? p[3 * threadIdx.x + 2] : 0.0;
    double n = 0.0, wsum = 0.0;
    for (int k = threadIdx.x; k < nblk; k += FH_MECH_THREADS) { n += p[3 * k]; wsum += p[3 * k] * p[3 * k + 1]; }
    n = fh_block_sum(n, red);
    wsum = fh_block_sum(wsum, red);
    const double mean = wsum / n;
    double m2 = 0.0;
    for (int k = threadIdx.x; k < nblk; k += FH_MECH_THREADS) { const double dlt = p[3 * k + 1] - mean; m2 += p[3 * k + 2] + p[3 * k] * dlt * dlt; }
    (void)ni; (void)mi; (void)Mi;
    const double M2 = fh_block_sum(m2, red);
    if (threadIdx.x == 0) {
      const double sd = sqrt(M2 / (n - 1.0));                               // torch.std: unbiased
      s_flag = sd > threshold ? 1 : 0;
      if (blockIdx.x == 0) flag_out[b] = s_flag;
    }
  }
  __syncthreads();
  const int use_cov = s_flag;
  const long i0 = (long)blockIdx.x * FH_MECH_CHUNK;
#pragma unroll
  for (int e = 0; e < FH_MECH_PER_THREAD; ++e) {
    const long i = i0 + threadIdx.x + (long)e * FH_MECH_THREADS;
    if (i < d) {
      const size_t j = b * d + i;
      // reference: p = Sigma.mat * scale / sigma^2 ; mu + p * sigma^2     |     p = g * scale ; mu + p * sigma^2
      const double pg = use_cov ? fh_mul(sigv[j], scale) / sigma2 : fh_mul(g[j], scale);
      out[j] = fh_add(mu[j], fh_mul(pg, sigma2));
    }
  }
}
// The decision alone, one CTA per image: flag[b] = std(g sigma^2) > threshold from the chunk statistics (the same merge, in
// the same order, as in fh_k_mech_combine).
__global__ void __launch_bounds__(FH_MECH_THREADS)
fh_k_mech_flag(const double* __restrict__ part, int nblk, double threshold, int* __restrict__ flag_out) {
  __shared__ double red[33];
  const size_t b = blockIdx.x;
  const double* p = part + b * nblk * 3;
  double n = 0.0, wsum = 0.0;
  for (int k = threadIdx.x; k < nblk; k += FH_MECH_THREADS) { n += p[3 * k]; wsum += p[3 * k] * p[3 * k + 1]; }
  n = fh_block_sum(n, red);
  wsum = fh_block_sum(wsum, red);
  const double mean = wsum / n;
  double m2 = 0.0;
  for (int k = threadIdx.x; k < nblk; k += FH_MECH_THREADS) { const double dlt = p[3 * k + 1] - mean; m2 += p[3 * k + 2] + p[3 * k] * dlt * dlt; }
  const double M2 = fh_block_sum(m2, red);
  if (threadIdx.x == 0) flag_out[b] = sqrt(M2 / (n - 1.0)) > threshold ? 1 : 0;
}
// score of the previous evaluation, conditioning_mechanisms.py:243: (mean_prev - x_prev) / sigma_prev^2, as float32 for the state
__global__ void __launch_bounds__(FH_MECH_THREADS)
fh_k_mech_score(const double* __restrict__ mean_prev, const double* __restrict__ x_prev, double sigma2, float* __restrict__ x32,
                float* __restrict__ score32, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { x32[i] = (float)x_prev[i]; score32[i] = (float)((mean_prev[i] - x_prev[i]) / sigma2); }
}
// torch.allclose(a, b) (rtol 1e-5, atol 1e-8), conditioning_mechanisms.py:250: flag[b] = 1 when image b is allclose
__global__ void __launch_bounds__(FH_MECH_THREADS)
fh_k_mech_allclose(const double* __restrict__ a, const double* __restrict__ bb, int* __restrict__ notclose, long d) {
  const size_t b = blockIdx.y;
  const long i0 = (long)blockIdx.x * FH_MECH_CHUNK;
  int bad = 0;
#pragma unroll
  for (int e = 0; e < FH_MECH_PER_THREAD; ++e) {
    const long i = i0 + threadIdx.x + (long)e * FH_MECH_THREADS;
    if (i < d) {
      const double x = a[b * d + i], y = bb[b * d + i];
      if (!(fabs(x - y) <= 1e-8 + 1e-5 * fabs(y))) bad = 1;
    }
  }
  if (bad) notclose[b] = 1;                                                  // benign race: every writer writes 1
}
// EDM Euler step (generate_conditional.py:146-149): d_cur = (x_hat - den) / t_cur ; x_prime = x_hat + h * d_cur
__global__ void __launch_bounds__(FH_MECH_THREADS)
fh_k_euler(const double* __restrict__ x_hat, const double* __restrict__ den, double t_cur, double h, double* __restrict__ d_cur,
           double* __restrict__ x_prime, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const double dc = (x_hat[i] - den[i]) / t_cur; d_cur[i] = dc; x_prime[i] = fh_add(x_hat[i], fh_mul(h, dc)); }
}
// Heun correction (:160-167): d_prime = x_prime / t_next - den2 / t_next ; x_next = x_hat + h * (0.5 * d_cur + 0.5 * d_prime)
__global__ void __launch_bounds__(FH_MECH_THREADS)
fh_k_heun(const double* __restrict__ x_hat, const double* __restrict__ d_cur, const double* __restrict__ x_prime, const double* __restrict__ den2,
          double t_next, double h, double* __restrict__ x_next, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double dp = x_prime[i] / t_next - den2[i] / t_next;
    x_next[i] = fh_add(x_hat[i], fh_mul(h, fh_add(fh_mul(0.5, d_cur[i]), fh_mul(0.5, dp))));
  }
}
