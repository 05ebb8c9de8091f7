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

#define FH_MAX_SEP_TAPS 64       // taps are staged zero-padded to a multiple of 4
#define FH_ZROUND_FLOATS 8192    // floats per TMA round: 32 / ROWS rows of Z, each a strip-wide slice of ROWS*256 floats (32 KB)

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
  volatile int* mirror;           // pinned, host-mapped (B, 2) ints: [progress = last iteration whose x/r update ran, done]
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

// The (done, tag) word of an image.  A kernel whose own CTAs may set `done` while others are starting (K1 / the p
// kernel on convergence, xr on the p.Ap break) must reach the same decision in every thread: "stopped by an EARLIER
// launch" = done set with a tag other than this launch's; if this very launch is setting it, every CTA re-derives the
// stop from the same partials instead of trusting the racing flag.
__device__ __forceinline__ bool fh_stopped_earlier(const FhCgFused& P, int b, int my_tag) {
  if (!P.sc) return false;
  const long long w = *(volatile const long long*)&P.sc[b].done;
  return (int)w != 0 && (int)(w >> 32) != my_tag;
}
__device__ __forceinline__ void fh_set_done(FhCgScalars* sc, int tag) {
  *(volatile long long*)&sc->done = ((long long)tag << 32) | 1LL;
}
// Zero-copy progress report to the host (pinned, mapped): the host throttles its launches on it instead of
// synchronising the stream, so the GPU queue never drains and no iteration is launched far past convergence.
__device__ __forceinline__ void fh_mirror_done(const FhCgFused& P, int b) { if (P.mirror) P.mirror[2 * b + 1] = 1; }
__device__ __forceinline__ void fh_mirror_progress(const FhCgFused& P, int b, int it) { if (P.mirror) P.mirror[2 * b] = it; }

// Consumer side of the r.r reduction of iteration `it - 1` (cg.py:258-263): returns true when image b has stopped.
// `record` (one CTA per image) stores the scalars for the host; every CTA gets beta.  Contains __syncthreads.
__device__ __forceinline__ bool fh_cg_consume_rr(const FhCgFused& P, int b, int it, bool record, double* s_red, float* beta_out) {
  const FhCgScalars* sc = P.sc + b;
  const double rr_old = sc->rr[(it - 1) & 1];          // loads issued before the reduction's barriers
  const double stop2 = sc->stop2;
  const double rr_new = fh_reduce_partials(P.part_b + (size_t)b * P.ystrips, P.ystrips, s_red);
  const bool converged = (stop2 >= 0.0) && (rr_new <= stop2);
  const bool stopped = converged || (it - 1 >= P.last_allowed);
  if (record && threadIdx.x == 0) {
    FhCgScalars* w = P.sc + b;
    w->resid2 = rr_new;
    w->rr[it & 1] = rr_new;
    w->niter = it - 1;
    if (converged) w->optimal = 1;
    if (stopped) { fh_set_done(w, it); fh_mirror_done(P, b); }
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
  // all 32 lanes of warp 0: rounds [j0, j0 + count) -- one lane per row slice, so the copies are issued in parallel
  __device__ __forceinline__ void issue(int j0, int count) {
    const int lane = threadIdx.x & 31;
    const int jr = lane / ZR, u = lane % ZR;
    if (jr < count) {
      const int j = j0 + jr;
      const int rows = min(ZR, m - j * ZR);
      const unsigned st = (issued + (unsigned)jr) % (unsigned)stages;
      if (u == 0) fh_mbar_expect_tx(bar + st, (unsigned)rows * SLICE * 4u);
      if (u < rows)
        fh_tma_load_1d(buf + (size_t)st * FH_ZROUND_FLOATS + (size_t)u * SLICE, src + (size_t)(j * ZR + u) * row_stride, SLICE * 4u, bar + st);
    }
    issued += (unsigned)count;
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
  __shared__ __align__(16) float s_taps[FH_MAX_SEP_TAPS];
  __shared__ double s_red[33];
  FH_TRACE(P, 0, 0);
  fh_pdl_trigger();
  fh_strip_tables<false, ROWS>(sm, P.tab);
  const int tc = (P.fuse == 2) ? fh_stage_taps(s_taps, P.kr, P.nr, P.cr, 0) : 0;
  fh_pdl_wait();                                        // everything below reads what the previous kernel wrote
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
      stopped = fh_stopped_earlier(P, b, it);
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
    if (P.fuse == 2) fh_strip_conv<ROWS>(sm.a, s_taps, P.nr, tc);
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
  __shared__ __align__(16) float s_taps[FH_MAX_SEP_TAPS];
  __shared__ float s_w[FH_PASS_THREADS / 32][FH_MAXM];      // per-warp fp32 sums; added up in fp64 below
  __shared__ __align__(8) fh_mbar_t s_bar[FH_ZMAX_STAGES];
  FH_DYN_SMEM(float, zbuf);
  FH_TRACE(P, 1, 0);
  fh_pdl_trigger();
  fh_strip_tables<false, ROWS>(sm, P.tab);
  const int tc = (P.fuse == 2) ? fh_stage_taps(s_taps, P.kc, P.nc, P.cc, 0) : 0;
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
  // prologue (may overlap the previous kernel): Z of this CTA's first strip starts streaming now
  __syncthreads();                                        // barriers initialised
  if ((int)blockIdx.x < nstrips) {
    const int b0 = blockIdx.x / P.strips_per_image;
    zs.begin(P.Z + (size_t)b0 * FH_MAXM * d + ((size_t)blockIdx.x * (ROWS * FH_N) - (size_t)b0 * d), d, P.m);
    if (tid < 32) zs.issue(0, min(zs.stages, zs.nrounds));
  }
  fh_pdl_wait();
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    const bool prefetched = (s == (int)blockIdx.x);
    const size_t base = (size_t)s * (ROWS * FH_N);
    const size_t off = base - (size_t)b * d;              // offset inside the image
    if (P.sc && P.sc[b].done) {
      if (prefetched) for (int j = 0; j < min(zs.stages, zs.nrounds); ++j) zs.wait();   // drain the prologue's copies
      continue;
    }
    __syncthreads();                                      // previous strip fully consumed (smem, Z buffers)
    if (!prefetched) {
      zs.begin(P.Z + (size_t)b * FH_MAXM * d + off, d, P.m);
      if (tid < 32) zs.issue(0, min(zs.stages, zs.nrounds));
    }
    FH_TRACE(P, 1, 1);
    fh_strip_load<ROWS>(sm.a, P.T + base);
    __syncthreads();
    FH_TRACE(P, 1, 2);
    if (P.fuse == 2) fh_strip_conv<ROWS>(sm.a, s_taps, P.nc, tc);
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
        if constexpr (ZR == 4) {
          // 4 warp sums with 6 shuffles: halve the value set while halving the lane set, then finish one value per lane
          const bool hi = lane & 16, mid = lane & 8;
          float ka = hi ? acc[2] : acc[0], kb = hi ? acc[3] : acc[1];
          ka += __shfl_xor_sync(0xffffffffu, hi ? acc[0] : acc[2], 16);
          kb += __shfl_xor_sync(0xffffffffu, hi ? acc[1] : acc[3], 16);
          float kp = mid ? kb : ka;
          kp += __shfl_xor_sync(0xffffffffu, mid ? ka : kb, 8);
          kp += __shfl_xor_sync(0xffffffffu, kp, 4);
          kp += __shfl_xor_sync(0xffffffffu, kp, 2);
          kp += __shfl_xor_sync(0xffffffffu, kp, 1);
          if ((lane & 7) == 0 && j * ZR + (lane >> 3) < P.m) s_w[warp][j * ZR + (lane >> 3)] = kp;   // lane 8r holds row r
        } else {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < ZR; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
          if (lane == 0) {
#pragma unroll
            for (int u = 0; u < ZR; ++u)
              if (j * ZR + u < P.m) s_w[warp][j * ZR + u] = acc[u];
          }
        }
        if (j + zs.stages < zs.nrounds) {                   // recycle this buffer for a later round
          __syncthreads();
          if (tid < 32) zs.issue(j + zs.stages, 1);
        }
      }
      __syncthreads();
      const int sidx = s - b * P.strips_per_image;
      if (tid < P.m) {
        double t = 0.0;
#pragma unroll
        for (int wv = 0; wv < FH_PASS_THREADS / 32; ++wv) t += (double)s_w[wv][tid];
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
  __shared__ __align__(16) float s_taps[FH_MAX_SEP_TAPS];
  __shared__ float s_c[FH_MAXM];
  __shared__ double s_sum[FH_MAXM];
  __shared__ __align__(8) fh_mbar_t s_bar[FH_ZMAX_STAGES];
  FH_DYN_SMEM(float, zbuf);
  FH_TRACE(P, 2, 0);
  fh_pdl_trigger();
  fh_strip_tables<true, ROWS>(sm, P.tab);
  const int tc = (P.fuse == 2) ? fh_stage_taps(s_taps, P.kc, P.nc, P.cc, 1) : 0;
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
  __syncthreads();                                        // barriers initialised
  if ((int)blockIdx.x < nstrips) {                        // prologue: Z of the first strip (independent of K2)
    const int b0 = blockIdx.x / P.strips_per_image;
    zs.begin(P.Z + (size_t)b0 * FH_MAXM * d + ((size_t)blockIdx.x * (ROWS * FH_N) - (size_t)b0 * d), d, P.m);
    if (tid < 32) zs.issue(0, min(zs.stages, zs.nrounds));
  }
  fh_pdl_wait();
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    const bool prefetched = (s == (int)blockIdx.x);
    const size_t base = (size_t)s * (ROWS * FH_N);
    const size_t off = base - (size_t)b * d;
    if (P.sc && P.sc[b].done) {
      if (prefetched) for (int j = 0; j < min(zs.stages, zs.nrounds); ++j) zs.wait();
      continue;
    }
    __syncthreads();
    if (!prefetched) {
      zs.begin(P.Z + (size_t)b * FH_MAXM * d + off, d, P.m);
      if (tid < 32) zs.issue(0, min(zs.stages, zs.nrounds));
    }
    FH_TRACE(P, 2, 1);
    float acc[ROWS], dl[ROWS];
    fh_g_get<ROWS>(P.W + base, acc);
    fh_g_get<ROWS>(P.delta + base, dl);
    if (b != cur_b && P.m > 0) {
      // coef = Gs * (sum over the strips of K2's partial dots), same order in every CTA; overlaps the Z copies
      cur_b = b;
      // thread (i, part): row i = tid / 4 of the partial matrix, strips k = part (mod 4); loads are batched 8 deep
      const int i = tid >> 2, part = tid & 3;
      double t = 0.0;
      if (i < P.m) {
        const double* col = P.dots_partial + (size_t)b * P.strips_per_image * FH_MAXM + i;
        for (int k0 = part; k0 < P.strips_per_image; k0 += 32) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = (k0 + 4 * u < P.strips_per_image) ? fh_ldcg(col + (size_t)(k0 + 4 * u) * FH_MAXM) : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) t += v[u];
        }
      }
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      if (part == 0 && i < P.m) s_sum[i] = t;
      __syncthreads();
      {
        // coef_i = sum_j Gs[i][j] s_sum[j]: 4 threads per row, 16 columns each (m <= 64), loads batched
        double a = 0.0;
        if (i < P.m) {
          const double* g = P.Gs + ((size_t)b * FH_MAXM + i) * FH_MAXM;
          double gv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) gv[u] = (part + 4 * u < P.m) ? g[part + 4 * u] : 0.0;
#pragma unroll
          for (int u = 0; u < 16; ++u) a += gv[u] * ((part + 4 * u < P.m) ? s_sum[part + 4 * u] : 0.0);
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (part == 0 && i < P.m) s_c[i] = (float)a;
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
      if (j + zs.stages < zs.nrounds) {
        __syncthreads();
        if (tid < 32) zs.issue(j + zs.stages, 1);
      }
    }
#pragma unroll
    for (int e = 0; e < ROWS; ++e) acc[e] *= dl[e];
    fh_strip_put<ROWS>(sm.a, acc);
    __syncthreads();
    FH_TRACE(P, 2, 11);
    fh_dct_strip<true, ROWS>(sm.a, sm.b, sm);
    __syncthreads();
    FH_TRACE(P, 2, 12);
    if (P.fuse == 2) fh_strip_conv<ROWS>(sm.b, s_taps, P.nc, tc);
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
  __shared__ __align__(16) float s_taps[FH_MAX_SEP_TAPS];
  __shared__ double s_red[33];
  fh_pdl_trigger();
  fh_strip_tables<true, ROWS>(sm, P.tab);
  const int tc = (P.fuse == 2) ? fh_stage_taps(s_taps, P.kr, P.nr, P.cr, 1) : 0;
  fh_pdl_wait();
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
    if (P.fuse == 2) fh_strip_conv<ROWS>(sm.b, s_taps, P.nr, tc);
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
  fh_pdl_trigger();
  fh_pdl_wait();
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
  fh_pdl_trigger();
  fh_pdl_wait();
  const int nstrips = P.B * P.ystrips;
  int cur_b = -1;
  bool stopped = false;
  float beta = 0.f;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.ystrips;
    if (b != cur_b) {
      cur_b = b;
      stopped = fh_stopped_earlier(P, b, it);
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
  fh_pdl_trigger();
  fh_pdl_wait();
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.ystrips;
    if (b != cur_b) {
      cur_b = b;
      stopped = fh_stopped_earlier(P, b, -it);
      if (!stopped) {
        const double rr_cur = P.sc[b].rr[it & 1];
        const double pAp = fh_reduce_partials(P.part_a + (size_t)b * P.ystrips, P.ystrips, s_red);
        if (pAp <= 1e-16) {
          stopped = true;
          if (s == b * P.ystrips && tid == 0) { P.sc[b].niter = it; fh_set_done(P.sc + b, -it); fh_mirror_done(P, b); }
        } else {
          alpha = (float)(rr_cur / pAp);
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
    if (tid == 0) {
      P.part_b[(size_t)b * P.ystrips + (s - b * P.ystrips)] = tot;
      if (s == b * P.ystrips) fh_mirror_progress(P, b, it);
    }
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
    sc.done = 0; sc.done_tag = 0; sc.optimal = 0; sc.niter = 0;
    P.sc[b] = sc;
  }
}
