// Fused CG iteration for the DCT-basis covariance state (256 x 256 planes).
//
// One iteration of  (sigma_y^2 I + A Sigma A^T) u = b   (conditioning_utils/cg.py:244-274 with the operator of
// conditioning_mechanisms.py:384-419 / 489-527 / 641-675) is five launches when A is a mask (inpainting) or a
// separable blur (Gaussian), because A / A^T fold into the row and column passes of the 2-D DCT:
//
//   K1 fwd_rows   p = r + beta p (written back) -> mask -> row blur^T -> row DCT            -> T  (transposed)
//   K2 fwd_cols   column blur^T -> column DCT = w_hat (basis layout) ; partial Z (delta w_hat)  [ticket: coef = Gs Z(delta w_hat)]
//   K3 inv_cols   delta (w_hat + Z^T coef) -> column IDCT -> column blur                    -> T' (transposed)
//   K4 inv_rows   row IDCT -> row blur -> mask = t ; q = sy2 p + t ; partial p.q               [ticket: alpha, pAp test]
//   K5 xr         x += alpha p ; r -= alpha q ; partial r.r                                     [ticket: stop test, beta]
//
// Non-separable operators (motion blur, x4 super-resolution) run K1..K4 without the fused prologue/epilogue and
// put their own convolution kernels and the element-wise p / q kernels around them.
// "ticket" = the last CTA to finish an image's strips reduces the per-strip partials in a fixed order (results are
// bit-reproducible and independent of the batch) and writes the per-image scalars the next kernel needs.
#pragma once
#include "fh_common.cuh"
#include "fh_dct.cuh"

#define FH_MAX_SEP_TAPS 64


struct FhCgFused {
  int B, C, m;                    // images, channels, rows of Z
  int fuse;                       // 0: plain passes, 1: mask folded in (inpainting), 2: separable blur folded in
  int strips_per_image;           // C * 32
  const float* mask; int mask_batch;
  const float* kr; const float* kc; int nr, nc, cr, cc;     // separable taps (row kernel, column kernel)
  const float* Z; const float* delta; const double* Gs;     // covariance state (basis layout)
  float *x, *r, *p, *q;           // CG vectors (B, n); n == d in the fused modes
  float *T, *W;                   // (B, d) scratch: half-transformed planes, w_hat
  double* dots_partial;           // (B, strips_per_image, FH_MAXM)
  double* coef;                   // (B, FH_MAXM)
  double *part_a, *part_b;        // (B, strips) scalar-dot partials
  unsigned* tickets;              // (B, 4)
  FhCgScalars* sc;
  float sy2;
  FhDctTables tab;
};

// All threads learn whether this CTA is the last of `total` arrivals at `ctr`; the counter is re-armed by the last.
__device__ __forceinline__ bool fh_last_arrival(unsigned* ctr, unsigned total) {
  __shared__ unsigned s_old;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_old = atomicAdd(ctr, 1u);
  __syncthreads();
  const bool last = (s_old == total - 1);
  if (last) {
    __threadfence();
    if (threadIdx.x == 0) *ctr = 0;
  }
  return last;
}

__device__ __forceinline__ double fh_ldcg(const double* p) {
#ifdef FH_EMU
  return *p;
#else
  return __ldcg(p);
#endif
}

// Fixed-order sum of `n` per-strip partial doubles by the whole CTA (result in every thread).
__device__ __forceinline__ double fh_reduce_partials(const double* part, int n, double* s_red /* >= 33 */) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += fh_ldcg(part + i);
  return fh_block_sum(s, s_red);
}

__device__ __forceinline__ float fh_mask_at(const FhCgFused& P, int b, int row, int col) {
  return P.mask[(size_t)(P.mask_batch > 1 ? b : 0) * FH_N * FH_N + row * FH_N + col];
}

// ---- K1 ------------------------------------------------------------------------------------------
// src != null: plain input (B,d) pixel layout.  src == null: fused CG prologue on P.p / P.r
//   update_p = 1: p <- r + beta p (first = 1: p <- r), written back;  update_p = 0: input is `vec` as it is (init: x).
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_fwd_rows(FhCgFused P, const float* __restrict__ src, const float* __restrict__ vec, int update_p, int first) {
  __shared__ __align__(16) FhStripSmem sm;
  __shared__ float s_taps[FH_MAX_SEP_TAPS];
  fh_strip_tables<false>(sm, P.tab);
  if (P.fuse == 2 && threadIdx.x < P.nr) s_taps[threadIdx.x] = P.kr[threadIdx.x];
  const int nstrips = P.B * P.strips_per_image;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    if (P.sc && P.sc[b].done) continue;
    const size_t base = (size_t)s * FH_STRIP_ELEMS;
    const int row0 = (s & 31) * FH_STRIP;
    __syncthreads();
    float v[8];
    if (src) {
      fh_g_get(src + base, v);
    } else if (update_p) {
      float rv[8];
      fh_g_get(P.r + base, rv);
      if (first) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = rv[e];
      } else {
        const float beta = (float)P.sc[b].beta;
        fh_g_get(P.p + base, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaf(beta, v[e], rv[e]);
      }
      fh_g_put(P.p + base, v);
    } else {
      fh_g_get(vec + base, v);
    }
    if (P.fuse == 1) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int f = threadIdx.x + FH_PASS_THREADS * h;
        const float4 mk = *reinterpret_cast<const float4*>(&P.mask[(size_t)(P.mask_batch > 1 ? b : 0) * FH_N * FH_N + (size_t)(row0 + (f >> 6)) * FH_N + 4 * (f & 63)]);
        v[4 * h] *= mk.x; v[4 * h + 1] *= mk.y; v[4 * h + 2] *= mk.z; v[4 * h + 3] *= mk.w;
      }
    }
    fh_strip_put(sm.a, v);
    __syncthreads();
    if (P.fuse == 2) fh_strip_conv(sm.a, s_taps, P.nr, P.cr, 1);
    if (threadIdx.x < 64) fh_dct8_core<false>(sm.a, sm.b, sm);
    __syncthreads();
    fh_strip_store_t(sm.b, P.T + (size_t)(s >> 5) * FH_N * FH_N, row0);
  }
}

// ---- K2 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_fwd_cols(FhCgFused P) {
  __shared__ __align__(16) FhStripSmem sm;
  __shared__ float s_taps[FH_MAX_SEP_TAPS];
  __shared__ double s_w[FH_PASS_THREADS / 32][FH_MAXM];
  __shared__ double s_part[4][FH_MAXM];
  __shared__ double s_sum[FH_MAXM];
  fh_strip_tables<false>(sm, P.tab);
  if (P.fuse == 2 && threadIdx.x < P.nc) s_taps[threadIdx.x] = P.kc[threadIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nstrips = P.B * P.strips_per_image;
  const long d = (long)P.strips_per_image * FH_STRIP_ELEMS;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    if (P.sc && P.sc[b].done) continue;
    const size_t base = (size_t)s * FH_STRIP_ELEMS;
    const size_t off = base - (size_t)b * d;              // offset inside the image
    __syncthreads();
    fh_strip_load(sm.a, P.T + base);
    __syncthreads();
    if (P.fuse == 2) fh_strip_conv(sm.a, s_taps, P.nc, P.cc, 1);
    if (tid < 64) fh_dct8_core<false>(sm.a, sm.b, sm);
    __syncthreads();
    float w[8];
    fh_strip_get(sm.b, w);
    fh_g_put(P.W + base, w);
    if (P.m > 0) {
      float dl[8];
      fh_g_get(P.delta + base, dl);
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] *= dl[e];
      const float* Zb = P.Z + (size_t)b * FH_MAXM * d + off;
      for (int i = 0; i < P.m; ++i) {
        float z[8];
        fh_g_get(Zb + (size_t)i * d, z);
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(z[e], w[e], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_w[warp][i] = (double)acc;
      }
      __syncthreads();
      const int sidx = s - b * P.strips_per_image;
      if (tid < P.m) {
        double t = 0.0;
#pragma unroll
        for (int wv = 0; wv < FH_PASS_THREADS / 32; ++wv) t += s_w[wv][tid];
        P.dots_partial[((size_t)b * P.strips_per_image + sidx) * FH_MAXM + tid] = t;
      }
      if (fh_last_arrival(P.tickets + 4 * b, (unsigned)P.strips_per_image)) {
        // coef = Gs * sum over strips (fixed order: 4 interleaved partial sums, then across)
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
          P.coef[(size_t)b * FH_MAXM + tid] = t;
        }
      }
    }
  }
}

// ---- K3 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_inv_cols(FhCgFused P) {
  __shared__ __align__(16) FhStripSmem sm;
  __shared__ float s_taps[FH_MAX_SEP_TAPS];
  __shared__ float s_c[FH_MAXM];
  fh_strip_tables<true>(sm, P.tab);
  if (P.fuse == 2 && threadIdx.x < P.nc) s_taps[threadIdx.x] = P.kc[threadIdx.x];
  const int tid = threadIdx.x;
  const int nstrips = P.B * P.strips_per_image;
  const long d = (long)P.strips_per_image * FH_STRIP_ELEMS;
  int cur_b = -1;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    if (P.sc && P.sc[b].done) continue;
    const size_t base = (size_t)s * FH_STRIP_ELEMS;
    const size_t off = base - (size_t)b * d;
    __syncthreads();
    if (b != cur_b && P.m > 0) {
      if (tid < P.m) s_c[tid] = (float)fh_ldcg(&P.coef[(size_t)b * FH_MAXM + tid]);
      cur_b = b;
      __syncthreads();
    }
    float acc[8], dl[8];
    fh_g_get(P.W + base, acc);
    const float* Zb = P.Z + (size_t)b * FH_MAXM * d + off;
    for (int i = 0; i < P.m; ++i) {
      float z[8];
      fh_g_get(Zb + (size_t)i * d, z);
      const float c = s_c[i];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(c, z[e], acc[e]);
    }
    fh_g_get(P.delta + base, dl);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= dl[e];
    fh_strip_put(sm.a, acc);
    __syncthreads();
    if (tid < 64) fh_dct8_core<true>(sm.a, sm.b, sm);
    __syncthreads();
    if (P.fuse == 2) fh_strip_conv(sm.b, s_taps, P.nc, P.cc, 0);
    fh_strip_store_t(sm.b, P.T + (size_t)(s >> 5) * FH_N * FH_N, (s & 31) * FH_STRIP);
  }
}

// ---- K4 ------------------------------------------------------------------------------------------
// dst != null: plain output (pixel layout).  dst == null: fused epilogue  q = sy2 * pv + t,  partial pv.q,
// ticket -> alpha = rr/pAp (or the pAp <= 1e-16 break of cg.py:250).  pv = P.p (iteration) or `vec` (init: x).
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_inv_rows(FhCgFused P, float* __restrict__ dst, const float* __restrict__ vec, int it) {
  __shared__ __align__(16) FhStripSmem sm;
  __shared__ float s_taps[FH_MAX_SEP_TAPS];
  __shared__ double s_red[33];
  fh_strip_tables<true>(sm, P.tab);
  if (P.fuse == 2 && threadIdx.x < P.nr) s_taps[threadIdx.x] = P.kr[threadIdx.x];
  const int tid = threadIdx.x;
  const int nstrips = P.B * P.strips_per_image;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / P.strips_per_image;
    if (P.sc && P.sc[b].done) continue;
    const size_t base = (size_t)s * FH_STRIP_ELEMS;
    const int row0 = (s & 31) * FH_STRIP;
    __syncthreads();
    fh_strip_load(sm.a, P.T + base);
    __syncthreads();
    if (tid < 64) fh_dct8_core<true>(sm.a, sm.b, sm);
    __syncthreads();
    if (P.fuse == 2) fh_strip_conv(sm.b, s_taps, P.nr, P.cr, 0);
    float t[8];
    fh_strip_get(sm.b, t);
    if (P.fuse == 1) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int f = tid + FH_PASS_THREADS * h;
        const float4 mk = *reinterpret_cast<const float4*>(&P.mask[(size_t)(P.mask_batch > 1 ? b : 0) * FH_N * FH_N + (size_t)(row0 + (f >> 6)) * FH_N + 4 * (f & 63)]);
        t[4 * h] *= mk.x; t[4 * h + 1] *= mk.y; t[4 * h + 2] *= mk.z; t[4 * h + 3] *= mk.w;
      }
    }
    if (dst) {
      fh_g_put(dst + base, t);
      continue;
    }
    const float* pv = vec ? vec : P.p;
    float p[8];
    fh_g_get(pv + base, p);
    float a32 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { t[e] = fmaf(P.sy2, p[e], t[e]); a32 = fmaf(p[e], t[e], a32); }
    fh_g_put(P.q + base, t);
    if (it <= 0) continue;                                  // init: no step length needed
    const double tot = fh_block_sum((double)a32, s_red);
    const int sidx = s - b * P.strips_per_image;
    if (tid == 0) P.part_a[(size_t)b * P.strips_per_image + sidx] = tot;
    if (fh_last_arrival(P.tickets + 4 * b + 1, (unsigned)P.strips_per_image)) {
      const double pAp = fh_reduce_partials(P.part_a + (size_t)b * P.strips_per_image, P.strips_per_image, s_red);
      if (tid == 0) {
        if (pAp <= 1e-16) { P.sc[b].done = 1; P.sc[b].niter = it; }
        else P.sc[b].alpha = P.sc[b].rr[it & 1] / pAp;
      }
    }
  }
}

// ---- element-wise CG kernels over strips of 2048 elements (n % 2048 == 0) ---------------------------
// q = sy2 p + t ; partial p.q ; ticket -> alpha        (non-fused operators; t = A Sigma A^T p)
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_q2(FhCgFused P, const float* __restrict__ t_in, const float* __restrict__ vec, int strips_per_image, int it) {
  __shared__ double s_red[33];
  const int tid = threadIdx.x;
  const int nstrips = P.B * strips_per_image;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / strips_per_image;
    if (P.sc[b].done) continue;
    const size_t base = (size_t)s * FH_STRIP_ELEMS;
    const float* pv = vec ? vec : P.p;
    float p[8], t[8];
    fh_g_get(pv + base, p);
    fh_g_get(t_in + base, t);
    float a32 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { t[e] = fmaf(P.sy2, p[e], t[e]); a32 = fmaf(p[e], t[e], a32); }
    fh_g_put(P.q + base, t);
    if (it <= 0) continue;
    const double tot = fh_block_sum((double)a32, s_red);
    const int sidx = s - b * strips_per_image;
    if (tid == 0) P.part_a[(size_t)b * strips_per_image + sidx] = tot;
    if (fh_last_arrival(P.tickets + 4 * b + 1, (unsigned)strips_per_image)) {
      const double pAp = fh_reduce_partials(P.part_a + (size_t)b * strips_per_image, strips_per_image, s_red);
      if (tid == 0) {
        if (pAp <= 1e-16) { P.sc[b].done = 1; P.sc[b].niter = it; }
        else P.sc[b].alpha = P.sc[b].rr[it & 1] / pAp;
      }
    }
  }
}

// p <- r + beta p (first: p <- r)                         (non-fused operators)
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_p(FhCgFused P, int strips_per_image, int first) {
  const int nstrips = P.B * strips_per_image;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / strips_per_image;
    if (P.sc[b].done) continue;
    const size_t base = (size_t)s * FH_STRIP_ELEMS;
    float v[8], rv[8];
    fh_g_get(P.r + base, rv);
    if (first) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = rv[e];
    } else {
      const float beta = (float)P.sc[b].beta;
      fh_g_get(P.p + base, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(beta, v[e], rv[e]);
    }
    fh_g_put(P.p + base, v);
  }
}

// K5:  x += alpha p ; r -= alpha q ; partial r.r ; ticket -> stop test (cg.py:259), beta, bookkeeping.
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_xr(FhCgFused P, int strips_per_image, int it, int last_allowed) {
  __shared__ double s_red[33];
  const int tid = threadIdx.x;
  const int nstrips = P.B * strips_per_image;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / strips_per_image;
    if (P.sc[b].done) continue;
    const size_t base = (size_t)s * FH_STRIP_ELEMS;
    const float alpha = (float)P.sc[b].alpha;
    float xv[8], pv[8], rv[8], qv[8];
    fh_g_get(P.x + base, xv);
    fh_g_get(P.p + base, pv);
    fh_g_get(P.r + base, rv);
    fh_g_get(P.q + base, qv);
    float rr = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      xv[e] = fmaf(alpha, pv[e], xv[e]);
      rv[e] = fmaf(-alpha, qv[e], rv[e]);
      rr = fmaf(rv[e], rv[e], rr);
    }
    fh_g_put(P.x + base, xv);
    fh_g_put(P.r + base, rv);
    const double tot = fh_block_sum((double)rr, s_red);
    const int sidx = s - b * strips_per_image;
    if (tid == 0) P.part_b[(size_t)b * strips_per_image + sidx] = tot;
    if (fh_last_arrival(P.tickets + 4 * b + 2, (unsigned)strips_per_image)) {
      const double rr_new = fh_reduce_partials(P.part_b + (size_t)b * strips_per_image, strips_per_image, s_red);
      if (tid == 0) {
        FhCgScalars* sc = P.sc + b;
        const double rr_old = sc->rr[it & 1];
        const bool converged = (sc->stop2 >= 0.0) && (rr_new <= sc->stop2);
        sc->resid2 = rr_new;
        sc->rr[(it + 1) & 1] = rr_new;
        sc->niter = it;
        sc->beta = rr_new / rr_old;
        if (converged) { sc->done = 1; sc->optimal = 1; }
        else if (it >= last_allowed) sc->done = 1;
      }
    }
  }
}

// init:  r = b - q ; partial r.r and b.b ; ticket -> scalars (stop threshold cg.py:237)
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_cg_r0(FhCgFused P, const float* __restrict__ bvec, int strips_per_image, double rtol, double atol, int use_stop) {
  __shared__ double s_red[33];
  const int tid = threadIdx.x;
  const int nstrips = P.B * strips_per_image;
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const int b = s / strips_per_image;
    const size_t base = (size_t)s * FH_STRIP_ELEMS;
    float bv[8], qv[8];
    fh_g_get(bvec + base, bv);
    fh_g_get(P.q + base, qv);
    float rr = 0.f, bb = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      qv[e] = bv[e] - qv[e];
      rr = fmaf(qv[e], qv[e], rr);
      bb = fmaf(bv[e], bv[e], bb);
    }
    fh_g_put(P.r + base, qv);
    const double t1 = fh_block_sum((double)rr, s_red);
    const double t2 = fh_block_sum((double)bb, s_red);
    const int sidx = s - b * strips_per_image;
    if (tid == 0) { P.part_a[(size_t)b * strips_per_image + sidx] = t1; P.part_b[(size_t)b * strips_per_image + sidx] = t2; }
    if (fh_last_arrival(P.tickets + 4 * b + 3, (unsigned)strips_per_image)) {
      const double rr0 = fh_reduce_partials(P.part_a + (size_t)b * strips_per_image, strips_per_image, s_red);
      const double bb0 = fh_reduce_partials(P.part_b + (size_t)b * strips_per_image, strips_per_image, s_red);
      if (tid == 0) {
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
  }
}
