// Fused CG iteration for the DCT-basis covariance state (256 x 256 planes).
//
// One iteration of  (sigma_y^2 I + A Sigma A^T) u = b   (conditioning_utils/cg.py:244-274 with the operator of
// conditioning_mechanisms.py:384-419 / 489-527 / 641-675) is five launches when A is a mask (inpainting) or a
// separable blur (Gaussian), because A / A^T fold into the row and column passes of the 2-D DCT:
//
//   K1 fwd_rows   p = r + beta p (written back) -> mask -> row blur^T -> row DCT            -> T  (transposed)
//   K2 fwd_cols   column blur^T -> column DCT = w_hat (basis layout) ; per-strip partial Z (delta w_hat)
//   K3 inv_cols   coef = Gs sum(partials) ; delta (w_hat + Z^T coef) -> column IDCT -> column blur -> T' (transposed)
//   K4 inv_rows   row IDCT -> row blur -> mask = t ; q = sy2 p + t ; per-strip partial p.q
//   K5 xr         alpha = rr / sum(partials) (pAp <= 1e-16 break) ; x += alpha p ; r -= alpha q ; per-strip partial r.r
//   (K1 of the next iteration: rr' = sum(partials) ; stop test ; beta = rr'/rr)
//
// Non-separable operators (motion blur, x4 super-resolution) run K1..K4 without the fused prologue/epilogue and
// put their own convolution kernels and the element-wise p / q kernels around them.
// Reductions are two-level and fence-free: a producer kernel writes one fp64 partial per strip; every CTA of the
// CONSUMER kernel sums an image's partials itself, in the same fixed order, so all CTAs get bit-identical scalars
// (no atomics, no tickets, no grid barrier; results are reproducible and independent of the batch), while the
// reduction overlaps the consumer's own loads.  The CTA that owns an image's first strip records the scalars
// (FhCgScalars) for the host's convergence poll.  Z is streamed through shared memory by 1-D TMA bulk copies.
#pragma once
#include "fh_common.cuh"
#include "fh_dct.cuh"

#define FH_MAX_SEP_TAPS 64
#define FH_ZROUND_FLOATS 8192    // floats per TMA round: 32 / ROWS rows of Z, each a strip-wide slice of ROWS*256 floats (32 KB)
#define ZROUND_FLOATS 8192    // floats per TMA round: 32 / ROWS rows of Z, each a strip-wide slice of ROWS*256 floats
#define FH_ZSTAGES 2            // rounds in flight (double buffer): 2 * 32 KB = 64 KB of dynamic shared memory
#define FH_ZBUF_BYTES (FH_ZSTAGES * ZROUND_FLOATS * 4)

struct FhCgFused {
  int B, C, m;                    // images, channels, rows of Z
  int fuse;                       // 0: plain passes, 1: mask folded in (inpainting), 2: separable blur folded in
  int strips_per_image;           // C * 256 / ROWS   (image-space strips)
  int ystrips;                    // measurement-space strips per image (n / (ROWS*256))
  int zstages;                    // TMA rounds in flight (dynamic shared memory = zstages * 32 KB)
  const float* mask; int mask_batch;
  const float* kr; const float* kc; int nr, nc, cr, cc;     // separable taps (row kernel, column kernel)
  const float* Z; const float* delta; const double* Gs;     // covariance state (basis layout)
  float *x, *r, *p, *q;           // CG vectors (B, n); n == d in the fused modes
  float *T, *W;                   // (B, d) scratch: half-transformed planes, w_hat
  double* dots_partial;           // (B, strips_per_image, FH_MAXM)
  double *part_a, *part_b;        // (B, ystrips) per-strip partials of p.q and r.r
  FhCgScalars* sc;
  float sy2;
  int last_allowed;               // iteration limit of this solve (maxiter or the fixed iteration count)
  FhDctTables tab;
  long long* trace;               // optional (FH_TRACE=1): clock64 checkpoints of CTA 0, 16 slots per kernel
};

#ifdef FH_EMU
#define FH_TRACE(P, kern, slot) do {} while (0)
#else
#define FH_TRACE(P, kern, slot) do { if ((P).trace && blockIdx.x == 0 && threadIdx.x == 0) (P).trace[(kern) * 16 + (slot)] = clock64(); } while (0)
#endif

__device__ __forceinline__ double fh_ldcg(const double* p) {
#ifdef FH_EMU
  return *p;
#else
  return __ldcg(p);
#endif
}

// Fixed-order sum of `n` per-strip partial doubles by the whole CTA (bit-identical result in every thread of every
// CTA that calls it on the same data).  Contains __syncthreads.
__device__ __forceinline__ double fh_reduce_partials(const double* part, int n, double* s_red /* >= 33 */) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += fh_ldcg(part + i);
  return fh_block_sum(s, s_red);
}

// CTA-uniform read of the image's done flag (another CTA of the same launch may be setting it).  Contains __syncthreads.
__device__ __forceinline__ bool fh_done_uniform(const FhCgFused& P, int b) {
  __shared__ int s_done;
  __syncthreads();
  if (threadIdx.x == 0) s_done = P.sc ? *(volatile const int*)&P.sc[b].done : 0;
  __syncthreads();
  return s_done != 0;
}

// Consumer side of the r.r reduction of iteration `it - 1` (cg.py:258-263): returns true when image b has stopped.
// `record` (one CTA per image) stores the scalars for the host; every CTA gets beta.
__device__ __forceinline__ bool fh_cg_consume_rr(const FhCgFused& P, int b, int it, bool record, double* s_red, float* beta_out) {
  const double rr_new = fh_reduce_partials(P.part_b + (size_t)b * P.ystrips, P.ystrips, s_red);
  const FhCgScalars* sc = P.sc + b;
  const double rr_old = sc->rr[(it - 1) & 1];
  const double stop2 = sc->stop2;
  const bool converged = (stop2 >= 0.0) && (rr_new <= stop2);
  const bool stopped = converged || (it - 1 >= P.last_allowed);
  if (record && threadIdx.x == 0) {
    FhCgScalars* w = P.sc + b;
    w->resid2 = rr_new;
    w->rr[it & 1] = rr_new;
    w->niter = it - 1;
    if (converged) w->optimal = 1;
    if (stopped) w->done = 1;
  }
  *beta_out = (float)(rr_new / rr_old);
  return stopped;
}

// v *= mask for the thread's ROWS elements of the strip starting at image row `row0` (mask is (MB,256,256))
template <int ROWS>
__device__ __forceinline__ void fh_mask_mul(const FhCgFused& P, int b, int row0, float (&v)[ROWS]) {
  float mk[ROWS];
  fh_g_get<ROWS>(P.mask + (size_t)(P.mask_batch > 1 ? b : 0) * FH_N * FH_N + (size_t)row0 * FH_N, mk);
#pragma unroll
  for (int e = 0; e < ROWS; ++e) v[e] *= mk[e];
}

// Streams the strip-wide slices of the m rows of Z through shared memory with 1-D TMA bulk copies issued by
// thread 0: round j = rows [ZR j, ZR (j+1)) lands in buffer j % stages and completes on bar[j % stages].  `issued` /
// `waited` count rounds over the whole kernel so the barrier phase parity stays right across strips.
#define FH_ZMAX_STAGES 5
template <int ROWS>
struct FhZStream {
  static constexpr int ZR = 32 / ROWS;                 // rows per round
  static constexpr int SLICE = ROWS * FH_N;            // floats per row slice
  float* buf; fh_mbar_t* bar; const float* src; long row_stride; int m; int nrounds; int stages; unsigned issued, waited;
  __device__ __forceinline__ void begin(const float* src_, long row_stride_, int m_) {
    src = src_; row_stride = row_stride_; m = m_; nrounds = (m_ + ZR - 1) / ZR;
  }
  // thread 0 only
  __device__ __forceinline__ void issue(int j) {
    const int rows = min(ZR, m - j * ZR);
    const unsigned st = issued % (unsigned)stages;
    fh_mbar_expect_tx(bar + st, (unsigned)rows * SLICE * 4u);
    for (int u = 0; u < rows; ++u)
      fh_tma_load_1d(buf + (size_t)st * FH_ZROUND_FLOATS + (size_t)u * SLICE, src + (size_t)(j * ZR + u) * row_stride, SLICE * 4u, bar + st);
    ++issued;
  }
  // all threads; returns the buffer of the round just completed
  __device__ __forceinline__ const float* wait() {
    const unsigned st = waited % (unsigned)stages;
    fh_mbar_wait(bar + st, (waited / (unsigned)stages) & 1u);
    ++waited;
    return buf + (size_t)st * FH_ZROUND_FLOATS;
  }
};

// ---- K1 ------------------------------------------------------------------------------------------
// src != null: plain input (B,d) pixel layout.  src == null: fused CG prologue on P.p / P.r
//   it >= 1: consume the r.r partials of iteration it-1 (it == 1: p <- r), p <- r + beta p, written back;
//   it == 0: input is `vec` as it is (initial residual: x).
template <int ROWS>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_fwd_rows(FhCgFused P, const float* __restrict__ src, const float* __restrict__ vec, int it) {
  __shared__ __align__(16) FhStripSmem<ROWS> sm;
  __shared__ float s_taps[FH_MAX_SEP_TAPS];
  __shared__ double s_red[33];
  FH_TRACE(P, 0, 0);
  fh_strip_tables<false, ROWS>(sm, P.tab);
  if (P.fuse == 2 && threadIdx.x < P.nr) s_taps[threadIdx.x] = P.kr[threadIdx.x];
  const int nstrips = P.B * P.strips_per_image;
  int cur_b = -1;
  bool stopped = false;
  float beta = 0.f;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    const size_t base = (size_t)s * (ROWS * FH_N);
    const int row0 = (s % (FH_N / ROWS)) * ROWS;
    __syncthreads();
    FH_TRACE(P, 0, 1);
    if (b != cur_b) {
      cur_b = b;
      stopped = fh_done_uniform(P, b);
      FH_TRACE(P, 0, 2);
      if (!src && it >= 2 && !stopped) stopped = fh_cg_consume_rr(P, b, it, s == b * P.strips_per_image, s_red, &beta);
    }
    FH_TRACE(P, 0, 3);
    if (stopped) continue;
    float v[ROWS];
    if (src) {
      fh_g_get<ROWS>(src + base, v);
    } else if (it >= 1) {
      float rv[ROWS];
      fh_g_get<ROWS>(P.r + base, rv);
      if (it == 1) {
#pragma unroll
        for (int e = 0; e < ROWS; ++e) v[e] = rv[e];
      } else {
        fh_g_get<ROWS>(P.p + base, v);
#pragma unroll
        for (int e = 0; e < ROWS; ++e) v[e] = fmaf(beta, v[e], rv[e]);
      }
      fh_g_put<ROWS>(P.p + base, v);
    } else {
      fh_g_get<ROWS>(vec + base, v);
    }
    if (P.fuse == 1) fh_mask_mul<ROWS>(P, b, row0, v);
    fh_strip_put<ROWS>(sm.a, v);
    __syncthreads();
    FH_TRACE(P, 0, 4);
    if (P.fuse == 2) fh_strip_conv<ROWS>(sm.a, s_taps, P.nr, P.cr, 1);
    FH_TRACE(P, 0, 5);
    fh_dct_strip<false, ROWS>(sm.a, sm.b, sm);
    __syncthreads();
    FH_TRACE(P, 0, 6);
    fh_strip_store_t<ROWS>(sm.b, P.T + (size_t)(s / (FH_N / ROWS)) * FH_N * FH_N, row0);
    FH_TRACE(P, 0, 7);
  }
}

// ---- K2 ------------------------------------------------------------------------------------------
template <int ROWS>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_fwd_cols(FhCgFused P) {
  __shared__ __align__(16) FhStripSmem<ROWS> sm;
  __shared__ float s_taps[FH_MAX_SEP_TAPS];
  __shared__ double s_w[FH_PASS_THREADS / 32][FH_MAXM];
  __shared__ __align__(8) fh_mbar_t s_bar[FH_ZMAX_STAGES];
  FH_DYN_SMEM(float, zbuf);
  FH_TRACE(P, 1, 0);
  fh_strip_tables<false, ROWS>(sm, P.tab);
  if (P.fuse == 2 && threadIdx.x < P.nc) s_taps[threadIdx.x] = P.kc[threadIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < P.zstages; ++i) fh_mbar_init(s_bar + i, 1);
    fh_mbar_fence_init();
  }
  FhZStream<ROWS> zs;
  constexpr int ZR = FhZStream<ROWS>::ZR;
  zs.buf = zbuf; zs.bar = s_bar; zs.stages = P.zstages; zs.issued = 0; zs.waited = 0;
  const int nstrips = P.B * P.strips_per_image;
  const long d = (long)P.strips_per_image * (ROWS * FH_N);
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    if (P.sc && P.sc[b].done) continue;
    const size_t base = (size_t)s * (ROWS * FH_N);
    const size_t off = base - (size_t)b * d;              // offset inside the image
    __syncthreads();                                      // previous strip fully consumed (smem, Z buffers); barriers initialised
    zs.begin(P.Z + (size_t)b * FH_MAXM * d + off, d, P.m);
    if (tid == 0)                                         // Z does not depend on the transform: start it now
      for (int j = 0; j < min(zs.stages, zs.nrounds); ++j) zs.issue(j);
    FH_TRACE(P, 1, 1);
    fh_strip_load<ROWS>(sm.a, P.T + base);
    __syncthreads();
    FH_TRACE(P, 1, 2);
    if (P.fuse == 2) fh_strip_conv<ROWS>(sm.a, s_taps, P.nc, P.cc, 1);
    FH_TRACE(P, 1, 3);
    fh_dct_strip<false, ROWS>(sm.a, sm.b, sm);
    __syncthreads();
    FH_TRACE(P, 1, 4);
    float w[ROWS];
    fh_strip_get<ROWS>(sm.b, w);
    fh_g_put<ROWS>(P.W + base, w);
    if (P.m > 0) {
      float dl[ROWS];
      fh_g_get<ROWS>(P.delta + base, dl);
#pragma unroll
      for (int e = 0; e < ROWS; ++e) w[e] *= dl[e];
      FH_TRACE(P, 1, 5);
      for (int j = 0; j < zs.nrounds; ++j) {
        const float* zb = zs.wait();
        if (j < 8) FH_TRACE(P, 1, 6 + j);
        float acc[ZR];
#pragma unroll
        for (int u = 0; u < ZR; ++u) {
          float z[ROWS];
          fh_g_get<ROWS>(zb + u * (ROWS * FH_N), z);
          acc[u] = 0.f;
#pragma unroll
          for (int e = 0; e < ROWS; ++e) acc[u] = fmaf(z[e], w[e], acc[u]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int u = 0; u < ZR; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < ZR; ++u)
            if (j * ZR + u < P.m) s_w[warp][j * ZR + u] = (double)acc[u];
        }
        __syncthreads();                                  // everyone is done with this buffer
        if (tid == 0 && j + zs.stages < zs.nrounds) zs.issue(j + zs.stages);
      }
      const int sidx = s - b * P.strips_per_image;
      if (tid < P.m) {
        double t = 0.0;
#pragma unroll
        for (int wv = 0; wv < FH_PASS_THREADS / 32; ++wv) t += s_w[wv][tid];
        P.dots_partial[((size_t)b * P.strips_per_image + sidx) * FH_MAXM + tid] = t;
      }
      FH_TRACE(P, 1, 14);
    }
  }
  FH_TRACE(P, 1, 15);
}

// ---- K3 ------------------------------------------------------------------------------------------
template <int ROWS>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_inv_cols(FhCgFused P) {
  __shared__ __align__(16) FhStripSmem<ROWS> sm;
  __shared__ float s_taps[FH_MAX_SEP_TAPS];
  __shared__ float s_c[FH_MAXM];
  __shared__ double s_part[4][FH_MAXM];
  __shared__ double s_sum[FH_MAXM];
  __shared__ __align__(8) fh_mbar_t s_bar[FH_ZMAX_STAGES];
  FH_DYN_SMEM(float, zbuf);
  FH_TRACE(P, 2, 0);
  fh_strip_tables<true, ROWS>(sm, P.tab);
  if (P.fuse == 2 && threadIdx.x < P.nc) s_taps[threadIdx.x] = P.kc[threadIdx.x];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < P.zstages; ++i) fh_mbar_init(s_bar + i, 1);
    fh_mbar_fence_init();
  }
  FhZStream<ROWS> zs;
  constexpr int ZR = FhZStream<ROWS>::ZR;
  zs.buf = zbuf; zs.bar = s_bar; zs.stages = P.zstages; zs.issued = 0; zs.waited = 0;
  const int nstrips = P.B * P.strips_per_image;
  const long d = (long)P.strips_per_image * (ROWS * FH_N);
  int cur_b = -1;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    if (P.sc && P.sc[b].done) continue;
    const size_t base = (size_t)s * (ROWS * FH_N);
    const size_t off = base - (size_t)b * d;
    __syncthreads();
    zs.begin(P.Z + (size_t)b * FH_MAXM * d + off, d, P.m);
    if (tid == 0)
      for (int j = 0; j < min(zs.stages, zs.nrounds); ++j) zs.issue(j);
    FH_TRACE(P, 2, 1);
    float acc[ROWS], dl[ROWS];
    fh_g_get<ROWS>(P.W + base, acc);
    fh_g_get<ROWS>(P.delta + base, dl);
    if (b != cur_b && P.m > 0) {
      // coef = Gs * (sum over the strips of K2's partial dots), same order in every CTA; overlaps the Z copies
      cur_b = b;
      const int i = tid & (FH_MAXM - 1), part = tid >> 6;
      if (i < P.m) {
        double t = 0.0;
        for (int k = part; k < P.strips_per_image; k += 4) t += fh_ldcg(&P.dots_partial[((size_t)b * P.strips_per_image + k) * FH_MAXM + i]);
        s_part[part][i] = t;
      }
      __syncthreads();
      if (tid < P.m) s_sum[tid] = (s_part[0][tid] + s_part[1][tid]) + (s_part[2][tid] + s_part[3][tid]);
      __syncthreads();
      if (tid < P.m) {
        const double* g = P.Gs + ((size_t)b * FH_MAXM + tid) * FH_MAXM;
        double t = 0.0;
        for (int j = 0; j < P.m; ++j) t += g[j] * s_sum[j];
        s_c[tid] = (float)t;
      }
      __syncthreads();
    }
    FH_TRACE(P, 2, 2);
    for (int j = 0; j < zs.nrounds; ++j) {
      const float* zb = zs.wait();
      if (j < 8) FH_TRACE(P, 2, 3 + j);
#pragma unroll
      for (int u = 0; u < ZR; ++u) {
        float z[ROWS];
        fh_g_get<ROWS>(zb + u * (ROWS * FH_N), z);
        const bool live = j * ZR + u < P.m;
        const float c = live ? s_c[j * ZR + u] : 0.f;
#pragma unroll
        for (int e = 0; e < ROWS; ++e) acc[e] = fmaf(c, live ? z[e] : 0.f, acc[e]);
      }
      __syncthreads();
      if (tid == 0 && j + zs.stages < zs.nrounds) zs.issue(j + zs.stages);
    }
#pragma unroll
    for (int e = 0; e < ROWS; ++e) acc[e] *= dl[e];
    fh_strip_put<ROWS>(sm.a, acc);
    __syncthreads();
    FH_TRACE(P, 2, 11);
    fh_dct_strip<true, ROWS>(sm.a, sm.b, sm);
    __syncthreads();
    FH_TRACE(P, 2, 12);
    if (P.fuse == 2) fh_strip_conv<ROWS>(sm.b, s_taps, P.nc, P.cc, 0);
    FH_TRACE(P, 2, 13);
    fh_strip_store_t<ROWS>(sm.b, P.T + (size_t)(s / (FH_N / ROWS)) * FH_N * FH_N, (s % (FH_N / ROWS)) * ROWS);
    FH_TRACE(P, 2, 14);
  }
}

// ---- K4 ------------------------------------------------------------------------------------------
// dst != null: plain output (pixel layout).  dst == null: fused epilogue  q = sy2 * pv + t,  per-strip partial pv.q
// (it >= 1).  pv = P.p (iteration) or `vec` (initial residual: x).
template <int ROWS>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_inv_rows(FhCgFused P, float* __restrict__ dst, const float* __restrict__ vec, int it) {
  __shared__ __align__(16) FhStripSmem<ROWS> sm;
  __shared__ float s_taps[FH_MAX_SEP_TAPS];
  __shared__ double s_red[33];
  fh_strip_tables<true, ROWS>(sm, P.tab);
  if (P.fuse == 2 && threadIdx.x < P.nr) s_taps[threadIdx.x] = P.kr[threadIdx.x];
  const int tid = threadIdx.x;
  const int nstrips = P.B * P.strips_per_image;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    if (P.sc && P.sc[b].done) continue;
    const size_t base = (size_t)s * (ROWS * FH_N);
    const int row0 = (s % (FH_N / ROWS)) * ROWS;
    __syncthreads();
    fh_strip_load<ROWS>(sm.a, P.T + base);
    __syncthreads();
    fh_dct_strip<true, ROWS>(sm.a, sm.b, sm);
    __syncthreads();
    if (P.fuse == 2) fh_strip_conv<ROWS>(sm.b, s_taps, P.nr, P.cr, 0);
    float t[ROWS];
    fh_strip_get<ROWS>(sm.b, t);
    if (P.fuse == 1) fh_mask_mul<ROWS>(P, b, row0, t);
    if (dst) {
      fh_g_put<ROWS>(dst + base, t);
      continue;
    }
    const float* pv = vec ? vec : P.p;
    float p[ROWS];
    fh_g_get<ROWS>(pv + base, p);
    float a32 = 0.f;
#pragma unroll
    for (int e = 0; e < ROWS; ++e) { t[e] = fmaf(P.sy2, p[e], t[e]); a32 = fmaf(p[e], t[e], a32); }
    fh_g_put<ROWS>(P.q + base, t);
    if (it <= 0) continue;                                  // initial residual: no step length needed
    const double tot = fh_block_sum((double)a32, s_red);
    if (tid == 0) P.part_a[(size_t)b * P.ystrips + (s - b * P.strips_per_image)] = tot;
  }
}

// ---- element-wise CG kernels over measurement-space strips of ROWS*256 elements ----------------------
// q = sy2 p + t ; per-strip partial p.q        (non-fused operators; t = A Sigma A^T p)
template <int ROWS>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_q2(FhCgFused P, const float* __restrict__ t_in, const float* __restrict__ vec, int it) {
  __shared__ double s_red[33];
  const int tid = threadIdx.x;
  const int nstrips = P.B * P.ystrips;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.ystrips;
    if (P.sc[b].done) continue;
    const size_t base = (size_t)s * (ROWS * FH_N);
    const float* pv = vec ? vec : P.p;
    float p[ROWS], t[ROWS];
    fh_g_get<ROWS>(pv + base, p);
    fh_g_get<ROWS>(t_in + base, t);
    float a32 = 0.f;
#pragma unroll
    for (int e = 0; e < ROWS; ++e) { t[e] = fmaf(P.sy2, p[e], t[e]); a32 = fmaf(p[e], t[e], a32); }
    fh_g_put<ROWS>(P.q + base, t);
    if (it <= 0) continue;
    const double tot = fh_block_sum((double)a32, s_red);
    if (tid == 0) P.part_a[(size_t)b * P.ystrips + (s - b * P.ystrips)] = tot;
  }
}

// p <- r + beta p (it == 1: p <- r) with the r.r reduction of iteration it-1 consumed here   (non-fused operators)
template <int ROWS>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_p(FhCgFused P, int it) {
  __shared__ double s_red[33];
  const int nstrips = P.B * P.ystrips;
  int cur_b = -1;
  bool stopped = false;
  float beta = 0.f;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.ystrips;
    if (b != cur_b) {
      cur_b = b;
      stopped = fh_done_uniform(P, b);
      if (it >= 2 && !stopped) stopped = fh_cg_consume_rr(P, b, it, s == b * P.ystrips, s_red, &beta);
    }
    if (stopped) continue;
    const size_t base = (size_t)s * (ROWS * FH_N);
    float v[ROWS], rv[ROWS];
    fh_g_get<ROWS>(P.r + base, rv);
    if (it == 1) {
#pragma unroll
      for (int e = 0; e < ROWS; ++e) v[e] = rv[e];
    } else {
      fh_g_get<ROWS>(P.p + base, v);
#pragma unroll
      for (int e = 0; e < ROWS; ++e) v[e] = fmaf(beta, v[e], rv[e]);
    }
    fh_g_put<ROWS>(P.p + base, v);
  }
}

// K5:  alpha = rr / p.Ap (break when p.Ap <= 1e-16, cg.py:250) ; x += alpha p ; r -= alpha q ; per-strip partial r.r
template <int ROWS>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_xr(FhCgFused P, int it) {
  __shared__ double s_red[33];
  const int tid = threadIdx.x;
  const int nstrips = P.B * P.ystrips;
  int cur_b = -1;
  bool stopped = false;
  float alpha = 0.f;
  FH_TRACE(P, 4, 0);
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.ystrips;
    if (b != cur_b) {
      cur_b = b;
      stopped = fh_done_uniform(P, b);
      if (!stopped) {
        const double pAp = fh_reduce_partials(P.part_a + (size_t)b * P.ystrips, P.ystrips, s_red);
        if (pAp <= 1e-16) {
          stopped = true;
          if (s == b * P.ystrips && tid == 0) { P.sc[b].done = 1; P.sc[b].niter = it; }
        } else {
          alpha = (float)(P.sc[b].rr[it & 1] / pAp);
        }
      }
    }
    FH_TRACE(P, 4, 1);
    if (stopped) continue;
    const size_t base = (size_t)s * (ROWS * FH_N);
    float xv[ROWS], pv[ROWS], rv[ROWS], qv[ROWS];
    fh_g_get<ROWS>(P.x + base, xv);
    fh_g_get<ROWS>(P.p + base, pv);
    fh_g_get<ROWS>(P.r + base, rv);
    fh_g_get<ROWS>(P.q + base, qv);
    float rr = 0.f;
#pragma unroll
    for (int e = 0; e < ROWS; ++e) {
      xv[e] = fmaf(alpha, pv[e], xv[e]);
      rv[e] = fmaf(-alpha, qv[e], rv[e]);
      rr = fmaf(rv[e], rv[e], rr);
    }
    fh_g_put<ROWS>(P.x + base, xv);
    fh_g_put<ROWS>(P.r + base, rv);
    const double tot = fh_block_sum((double)rr, s_red);
    if (tid == 0) P.part_b[(size_t)b * P.ystrips + (s - b * P.ystrips)] = tot;
    FH_TRACE(P, 4, 2);
  }
}

// Host poll point after iteration it-1: records the scalars K1 / fh_k_cg_p of iteration `it` would record.  grid = B.
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_finish(FhCgFused P, int it) {
  __shared__ double s_red[33];
  const int b = blockIdx.x;
  if (P.sc[b].done) return;
  float beta;
  fh_cg_consume_rr(P, b, it, true, s_red, &beta);
}

// init:  r = b - q ; per-strip partials of r.r and b.b
template <int ROWS>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_r0(FhCgFused P, const float* __restrict__ bvec) {
  __shared__ double s_red[33];
  const int tid = threadIdx.x;
  const int nstrips = P.B * P.ystrips;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.ystrips;
    const size_t base = (size_t)s * (ROWS * FH_N);
    float bv[ROWS], qv[ROWS];
    fh_g_get<ROWS>(bvec + base, bv);
    fh_g_get<ROWS>(P.q + base, qv);
    float rr = 0.f, bb = 0.f;
#pragma unroll
    for (int e = 0; e < ROWS; ++e) {
      qv[e] = bv[e] - qv[e];
      rr = fmaf(qv[e], qv[e], rr);
      bb = fmaf(bv[e], bv[e], bb);
    }
    fh_g_put<ROWS>(P.r + base, qv);
    const double t1 = fh_block_sum((double)rr, s_red);
    const double t2 = fh_block_sum((double)bb, s_red);
    if (tid == 0) { P.part_a[(size_t)b * P.ystrips + (s - b * P.ystrips)] = t1; P.part_b[(size_t)b * P.ystrips + (s - b * P.ystrips)] = t2; }
  }
}
// scalars of a fresh solve from the partials of fh_k_cg_r0 (stop threshold cg.py:237).  grid = B.
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_r0_scalars(FhCgFused P, double rtol, double atol, int use_stop) {
  __shared__ double s_red[33];
  const int b = blockIdx.x;
  const double rr0 = fh_reduce_partials(P.part_a + (size_t)b * P.ystrips, P.ystrips, s_red);
  const double bb0 = fh_reduce_partials(P.part_b + (size_t)b * P.ystrips, P.ystrips, s_red);
  if (threadIdx.x == 0) {
    FhCgScalars sc;
    sc.rr[0] = rr0; sc.rr[1] = rr0;
    sc.bnorm2 = bb0;
    const double stop = fmax(rtol * sqrt(bb0), atol);
    sc.stop2 = use_stop ? stop * stop : -1.0;
    sc.resid2 = rr0;
    sc.alpha = 0.0; sc.beta = 0.0;
    sc.done = 0; sc.optimal = 0; sc.niter = 0; sc.pad = 0;
    P.sc[b] = sc;
  }
}
