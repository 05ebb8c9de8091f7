// The covariance updates as ONE cooperative launch each, float64 transfer: update_time_step (online_update_bfgs.py:153-192
// + DCT wrapper :357-360) and update_space_step (:247-312 + :362-368) of the DCT-basis state.
//
// Round 1 ran a time update as 16 launches (8 stand-alone DCT passes, delta, three-row helper, Gram, Gram reduction, core,
// apply, ...; 186 us at rank 32) in float32, and a space update as 14 (158 us).  Here the strips of fh_k_cg_res64 are
// reused: CTA (plane, c) owns 6 rows of the plane in the pixel and in the basis layout, its slices of Z sit in shared
// memory (1-D TMA), the 2-D DCTs are the float64 strip transforms with the transposed exchange through L2, and two extra
// "core" CTAs (the grid is 129 + 2 of the 148 SMs) do the m x m algebra while the strip CTAs wait at a grid barrier.
//
//   time update     P0  delta' (stored), s = (mean - x) / sigma^2, row DCT -> T0, Gram partials of Z     | barrier 1
//                   P1  column DCT = s_hat, Gram reduction (every CTA a share of the entries) -> K,
//                       partials of Z (delta s_hat), Z (delta pi s_hat), Z s_hat                          | barrier 2
//                   P2  core CTA: A = I + K Gz, Gauss-Jordan, G = -Gz A^-1, coef (second core CTA: the
//                       Sigma-side core when the Hessian-side variance has drifted away from it)          | barrier 3
//                   P3  s_hat' = delta / var2 (s_hat (1 + var2_old pi) + Z^T coef), column IDCT -> T1     | barrier 4
//                   P4  row IDCT = s', mean' = x + sigma'^2 s'                      (rank 0: barriers 1 and 4 only)
//   space update    P0  dx = x' - x, de = sigma^2 (mu' - mu_hat) (float64), row DCTs -> T0, T1, Gram of Z | barrier 1
//                   P1  column DCTs, rounded to the float32 the state stores; Gram reduction; Z de, dx.de | barrier 2
//                   P2  g = de / delta + Z^T Gz (Z de) -> rows m, m+1 of Z; their Gram rows; de.g         | barrier 3
//                   P3  core CTA: BFGS block of Gz, K extended, Gs = -Gz (I + K Gz)^-1
//
// Arithmetic: everything a vector passes through is float64 (the transfer of the score, the differences, the DCTs, the
// low-rank products, all reductions).  The Gram Z diag(w) Z^T keeps round 1's arithmetic -- float32 products summed in
// float32 over 48 columns per lane, float64 from there on -- because B200's FP64 pipe would make it the longest phase
// (1.5 M FMAs per CTA at rank 32: 13 us as DFMA against 5 us as FFMA) and its operands are the float32 Z anyway.
#pragma once
#include "fh_cg_res64.cuh"
#include "fh_kernels.cuh"

#define FH_U64_CP 4                                        // column parts of a strip in the Gram: unit of work = (8 x 8 pair block, column part)
#define FH_U64_NB (FH_MAXM / 8)
#define FH_U64_NBLK2 (FH_U64_NB * (FH_U64_NB + 1) / 2)      // 36 upper-triangular 8 x 8 blocks at rank 64
#define FH_U64_DP (4 * FH_MAXM + 8)                        // doubles of dot partials per strip CTA
#define FH_U64_LD FH_GRAM_LD                               // leading dimension of K (68)

struct FhUpd64 {
  int mode;                     // 1: time update, 2: space update
  int m;                        // rank before the update
  int C;                        // channel planes
  int consistent;               // time update: the Hessian-side variance equals the Sigma-side one (one core serves both)
  int mr_cap;                   // rows of Z that fit in shared memory
  // the image's state (pointers already offset to the image)
  const float* pi; float* delta; float* Z; double* Gz; double* Gs;
  long d;
  // time update: s = score_in, or (mean_prev - x_prev) / sig2_t when mean_prev != null
  const double* x_prev; const double* mean_prev; const double* score_in;
  double sig2_t, var2_old, var2_new, sig2_new;
  double* mean_out; double* score_out;
  // space update
  const double *x, *xnext, *mean_at_x, *mean_at_xnext;
  double sig2;                  // sigma_t^2 (scale of de); delta is read from the state as it stands
  // workspace
  double* T[2];                 // exchange planes (d each)
  double* gpart;                // Gram partials  [2][strips * FH_U64_CP][nblk2 * 64]
  double* K;                    // [2][FH_U64_LD * FH_U64_LD]
  double* dpart;                // [strips][FH_U64_DP]
  double* coef;                 // [FH_MAXM]
  unsigned* grid_counter;
  int* status;
  const double2* tw; const double2* e512;
};

// float32 strip store in the thread layout (float2 #t, #t+256, #t+512)
__device__ __forceinline__ void fh_u64_storef(float* dst, const float (&r)[FH_RES_ROWS], int npair) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h)
    if (h < npair) reinterpret_cast<float2*>(dst)[threadIdx.x + FH_PASS_THREADS * h] = make_float2(r[2 * h], r[2 * h + 1]);
}
__device__ __forceinline__ void fh_u64_zrow(const float* Zs, const float* zsrc, long d, int mr, int i, float (&z)[FH_RES_ROWS], int npair) {
  if (i < mr) fh_r64_getf(Zs + (size_t)i * FH_RES_SLICE, z, npair);
  else fh_r64_get_z_l2(zsrc + (size_t)i * d, z, npair);
}
__device__ __forceinline__ float2 fh_u64_z2(const float* Zs, const float* zsrc, long d, int mr, int m, int i, int idx) {
  if (i >= m) return make_float2(0.f, 0.f);
  if (i < mr) return reinterpret_cast<const float2*>(Zs + (size_t)i * FH_RES_SLICE)[idx];
  return fh_ld_cg(reinterpret_cast<const float2*>(zsrc + (size_t)i * d) + idx);
}

// Block totals of  sum_e Z_i[e] vec_v[e]  for i < m, v < NV  ->  out[v * FH_MAXM + i]  (this CTA's partial, global).
// float64 products of the widened rows, 32 dot products per warp transpose.  s_w: FH_PASS_THREADS / 32 x NV x FH_MAXM doubles.
template <int NV>
__device__ __forceinline__ void fh_u64_dots(const float* Zs, const float* zsrc, long d, int m, int mr, int npair,
                                            const double (&vec)[NV][FH_RES_ROWS], double* s_w, double* out) {
  constexpr int RB = 32 / NV;                              // rows per batch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < m; c0 += RB) {
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.0;
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
      if (c0 + rr < m) {
        float z[FH_RES_ROWS];
        fh_u64_zrow(Zs, zsrc, d, mr, c0 + rr, z, npair);
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) {
          const double zw = FH_WIDEN(z[e], e);
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[rr * NV + v] = fma(zw, vec[v][e], acc[rr * NV + v]);
        }
      }
    }
    const double tot = fh_warp_transpose_sum32_d(acc);
    const int rr = lane / NV, v = lane - rr * NV;
    if (rr < RB && c0 + rr < m) s_w[(warp * NV + v) * FH_MAXM + c0 + rr] = tot;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < NV * FH_MAXM; idx += FH_PASS_THREADS) {
    const int v = idx / FH_MAXM, i = idx - v * FH_MAXM;
    if (i < m) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < FH_PASS_THREADS / 32; ++w) t += s_w[(w * NV + v) * FH_MAXM + i];
      out[idx] = t;
    }
  }
  __syncthreads();
}

// Gram partials of the strip: for every upper-triangular 8 x 8 block of row pairs and every column part,
//   sum over the part's columns of  Z_i w Z_j        (w = wbuf[0], and wbuf[1] as a second set when TWO)
// One warp per unit; a lane owns 2 columns of every 64-column chunk; float32 products and per-lane sums (nrows x 2 terms
// x 4 parts = what round 1's kernel summed in float32), float64 across lanes.  gpart: [set][strip * CP + part][nblk2][64].
template <bool TWO>
__device__ __forceinline__ void fh_u64_gram(const float* Zs, const float* zsrc, long d, int m, int mr, int nrows, const float* wbuf,
                                            double* gpart, int strip, int nstrip) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = (m + 7) >> 3, nblk2 = nb * (nb + 1) / 2;
  const size_t set_stride = (size_t)nstrip * FH_U64_CP * nblk2 * 64;
  for (int unit = warp; unit < nblk2 * FH_U64_CP; unit += FH_PASS_THREADS / 32) {
    int blk = unit / FH_U64_CP;
    const int cp = unit - blk * FH_U64_CP;
    int bi = 0, rem = blk;
    while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
    const int bj = bi + rem;
#ifdef FH_U64_GRAM_F64
    typedef double gacc_t;
#else
    typedef float gacc_t;
#endif
    gacc_t acc[64], acc2[TWO ? 64 : 1];
#pragma unroll
    for (int k = 0; k < 64; ++k) acc[k] = 0.f;
    if (TWO) {
#pragma unroll
      for (int k = 0; k < 64; ++k) acc2[TWO ? k : 0] = 0.f;
    }
    for (int it = 0; it < nrows; ++it) {
      const int idx = (cp * nrows + it) * 32 + lane;               // float2 index inside the strip slice
      const float2 w0 = reinterpret_cast<const float2*>(wbuf)[idx];
      float2 w1 = make_float2(0.f, 0.f);
      if (TWO) w1 = reinterpret_cast<const float2*>(wbuf + FH_RES_SLICE)[idx];
      float2 a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = fh_u64_z2(Zs, zsrc, d, mr, m, 8 * bi + u, idx);
      if (bi == bj) {
#pragma unroll
        for (int u = 0; u < 8; ++u) b[u] = a[u];
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) b[u] = fh_u64_z2(Zs, zsrc, d, mr, m, 8 * bj + u, idx);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const gacc_t ax = (gacc_t)a[u].x * (gacc_t)w0.x, ay = (gacc_t)a[u].y * (gacc_t)w0.y;
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[8 * u + v] = ay * (gacc_t)b[v].y + (ax * (gacc_t)b[v].x + acc[8 * u + v]);
        if (TWO) {
          const gacc_t cx = (gacc_t)a[u].x * (gacc_t)w1.x, cy = (gacc_t)a[u].y * (gacc_t)w1.y;
#pragma unroll
          for (int v = 0; v < 8; ++v) acc2[TWO ? 8 * u + v : 0] = cy * (gacc_t)b[v].y + (cx * (gacc_t)b[v].x + acc2[TWO ? 8 * u + v : 0]);
        }
      }
    }
    double* o = gpart + (((size_t)strip * FH_U64_CP + cp) * nblk2 + blk) * 64;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = (double)acc[32 * half + k];
      o[32 * half + lane] = fh_warp_transpose_sum32_d(v);
      if (TWO) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = (double)acc2[TWO ? 32 * half + k : 0];
        o[set_stride + 32 * half + lane] = fh_warp_transpose_sum32_d(v);
      }
    }
  }
}

// Every strip CTA sums a share of the Gram entries over all (strip, part) partials, in a fixed order, into K (both
// triangles).  Diagonal blocks: only u <= v is used, so K is exactly symmetric.
__device__ __forceinline__ void fh_u64_gram_reduce(const double* gpart, double* K, int m, int nsets, int strip, int nstrip) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = (m + 7) >> 3, nblk2 = nb * (nb + 1) / 2;
  const int nparts = nstrip * FH_U64_CP;
  const size_t set_stride = (size_t)nparts * nblk2 * 64;
  for (int ent = strip * (FH_PASS_THREADS / 32) + warp; ent < nsets * nblk2 * 64; ent += nstrip * (FH_PASS_THREADS / 32)) {
    const int set = ent / (nblk2 * 64);
    int rem = ent - set * (nblk2 * 64);
    const int blk = rem >> 6, e = rem & 63, u = e >> 3, v = e & 7;
    int bi = 0;
    rem = blk;
    while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
    const int bj = bi + rem, i = 8 * bi + u, j = 8 * bj + v;
    if (i >= m || j >= m || (bi == bj && u > v)) continue;
    const double* src = gpart + set * set_stride + (size_t)blk * 64 + e;
    double t = 0.0;
    for (int p = lane; p < nparts; p += 32) t += fh_ldcg(src + (size_t)p * nblk2 * 64);
    t = fh_warp_sum(t);
    if (lane == 0) {
      double* Kd = K + (size_t)set * FH_U64_LD * FH_U64_LD;
      Kd[i * FH_U64_LD + j] = t;
      Kd[j * FH_U64_LD + i] = t;
    }
  }
}

// ---- the m x m algebra of a core CTA (256 threads, everything in shared memory) --------------------------------------
// G = -Gz (I + K Gz)^-1, symmetrised, left in `A` (ld FH_MAXM).  Ks: ld FH_U64_LD, Gzs: ld FH_MAXM.
__device__ __forceinline__ int fh_u64_core(const double* Ks, const double* Gzs, int m, double* A, double* Ai, double* red, int* piv) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int j = tid & (FH_MAXM - 1), g = tid / FH_MAXM, ng = nt / FH_MAXM;
  for (int i0 = 4 * g; i0 < m; i0 += 4 * ng) {
    double s4[4] = {0.0, 0.0, 0.0, 0.0};
    if (j < m) {
      for (int k = 0; k < m; ++k) {
        const double gz = Gzs[k * FH_MAXM + j];
#pragma unroll
        for (int q = 0; q < 4; ++q) s4[q] = fma(Ks[min(i0 + q, m - 1) * FH_U64_LD + k], gz, s4[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (i0 + q < m) A[(i0 + q) * FH_MAXM + j] = s4[q] + ((i0 + q == j) ? 1.0 : 0.0);
    }
  }
  __syncthreads();
#ifdef FH_EMU_DEBUG_UPD
  if (tid < m) printf("A[%d] = %.17g %.17g %.17g %.17g %.17g %.17g | Gz %.17g %.17g %.17g %.17g %.17g %.17g\n", tid, A[tid * FH_MAXM], A[tid * FH_MAXM + 1], A[tid * FH_MAXM + 2], A[tid * FH_MAXM + 3], A[tid * FH_MAXM + 4], A[tid * FH_MAXM + 5],
      Gzs[tid * FH_MAXM], Gzs[tid * FH_MAXM + 1], Gzs[tid * FH_MAXM + 2], Gzs[tid * FH_MAXM + 3], Gzs[tid * FH_MAXM + 4], Gzs[tid * FH_MAXM + 5]);
  __syncthreads();
#endif
  const int bad = fh_invert(A, Ai, m, red, piv);
  __syncthreads();
  for (int i0 = 4 * g; i0 < m; i0 += 4 * ng) {
    double s4[4] = {0.0, 0.0, 0.0, 0.0};
    if (j < m) {
      for (int k = 0; k < m; ++k) {
        const double ai = Ai[k * FH_MAXM + j];
#pragma unroll
        for (int q = 0; q < 4; ++q) s4[q] = fma(Gzs[min(i0 + q, m - 1) * FH_MAXM + k], ai, s4[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (i0 + q < m) A[(i0 + q) * FH_MAXM + j] = -s4[q];
    }
  }
  __syncthreads();
  // symmetrise in place: the pair (i, j), i < j, is owned by one thread
  for (int idx = tid; idx < m * m; idx += nt) {
    const int i = idx / m, jj = idx - i * m;
    if (i < jj) {
      const double t = 0.5 * (A[i * FH_MAXM + jj] + A[jj * FH_MAXM + i]);
      A[i * FH_MAXM + jj] = t;
      A[jj * FH_MAXM + i] = t;
    }
  }
  __syncthreads();
  return bad;
}

struct FhUpd64Core {                                        // shared memory of a core CTA
  double Ks[FH_U64_LD * FH_U64_LD];
  double Gzs[FH_MAXM * FH_MAXM];
  double A[FH_MAXM * FH_MAXM];
  double Ai[FH_MAXM * FH_MAXM];
  double red[4 * FH_MAXM];
  double dsum[FH_U64_DP];
  int piv[4];
};

// sum of the strips' dot partials, entry by entry, fixed order (core CTA; needs a barrier before dsum is read)
__device__ __forceinline__ void fh_u64_sum_dpart(const double* dpart, int nstrip, int n, double* dsum) {
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    int p = 0;
    for (; p + 4 <= nstrip; p += 4) {
      t0 += fh_ldcg(dpart + (size_t)p * FH_U64_DP + idx); t1 += fh_ldcg(dpart + (size_t)(p + 1) * FH_U64_DP + idx);
      t2 += fh_ldcg(dpart + (size_t)(p + 2) * FH_U64_DP + idx); t3 += fh_ldcg(dpart + (size_t)(p + 3) * FH_U64_DP + idx);
    }
    for (; p < nstrip; ++p) t0 += fh_ldcg(dpart + (size_t)p * FH_U64_DP + idx);
    dsum[idx] = (t0 + t1) + (t2 + t3);
  }
}

__global__ void __launch_bounds__(FH_PASS_THREADS, 1)
fh_k_upd64(FhUpd64 U) {
  FH_DYN_SMEM(unsigned char, smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nstrip = U.C * FH_R64_CPP;
  const unsigned nblk = gridDim.x;                           // nstrip + 2
  const int s = blockIdx.x;
  const int m = U.m;
  const long d = U.d;
  unsigned bar_target = 0;

  // =========================================================================================== core CTAs
  if (s >= nstrip) {
    FhUpd64Core& cs = *reinterpret_cast<FhUpd64Core*>(smem_raw);
    const int which = s - nstrip;                            // 0: the core the update itself needs; 1: Sigma-side core of an inconsistent time update
    if (U.mode == 1) {
      fh_grid_sync(U.grid_counter, bar_target, nblk);        // barrier 1
      if (m == 0) { fh_grid_sync(U.grid_counter, bar_target, nblk); return; }      // barrier 4
      fh_grid_sync(U.grid_counter, bar_target, nblk);        // barrier 2
      const bool work = (which == 0) || !U.consistent;
      if (work) {
        const double* Kg = U.K + (size_t)which * FH_U64_LD * FH_U64_LD;
        for (int idx = tid; idx < m * m; idx += FH_PASS_THREADS) {
          const int i = idx / m, j = idx - i * m;
          cs.Ks[i * FH_U64_LD + j] = fh_ldcg(Kg + i * FH_U64_LD + j);
          cs.Gzs[i * FH_MAXM + j] = U.Gz[i * FH_MAXM + j];
        }
        if (which == 0) fh_u64_sum_dpart(U.dpart, nstrip, 3 * FH_MAXM, cs.dsum);
        __syncthreads();
        const int bad = fh_u64_core(cs.Ks, cs.Gzs, m, cs.A, cs.Ai, cs.red, cs.piv);
        if (bad && tid == 0) atomicAdd(U.status, 1);
        if (which == 1 || U.consistent) {
          for (int idx = tid; idx < m * m; idx += FH_PASS_THREADS) {
            const int i = idx / m, j = idx - i * m;
            U.Gs[i * FH_MAXM + j] = cs.A[i * FH_MAXM + j];
          }
        }
        if (which == 0) {
          // coef = var2_old c1 + G zt,  c1 = Gz zs,  zt = zds + var2_old zdps + var2_old K c1
          const double* zds = cs.dsum; const double* zdps = cs.dsum + FH_MAXM; const double* zs = cs.dsum + 2 * FH_MAXM;
          double* c1 = cs.red; double* zt = cs.red + FH_MAXM;
          if (tid < m) {
            double t = 0.0;
            for (int k = 0; k < m; ++k) t = fma(cs.Gzs[tid * FH_MAXM + k], zs[k], t);
            c1[tid] = t;
          }
          __syncthreads();
          if (tid < m) {
            double kc = 0.0;
            for (int k = 0; k < m; ++k) kc = fma(cs.Ks[tid * FH_U64_LD + k], c1[k], kc);
            zt[tid] = zds[tid] + U.var2_old * zdps[tid] + U.var2_old * kc;
          }
          __syncthreads();
          if (tid < m) {
            double t = 0.0;
            for (int k = 0; k < m; ++k) t = fma(cs.A[tid * FH_MAXM + k], zt[k], t);
            U.coef[tid] = U.var2_old * c1[tid] + t;
#ifdef FH_EMU_DEBUG_UPD
            printf("K[%d] = %.17g %.17g %.17g %.17g %.17g %.17g\n", tid, cs.Ks[tid * FH_U64_LD], cs.Ks[tid * FH_U64_LD + 1], cs.Ks[tid * FH_U64_LD + 2], cs.Ks[tid * FH_U64_LD + 3], cs.Ks[tid * FH_U64_LD + 4], cs.Ks[tid * FH_U64_LD + 5]);
            printf("G[%d] = %.17g %.17g %.17g %.17g %.17g %.17g\n", tid, cs.A[tid * FH_MAXM], cs.A[tid * FH_MAXM + 1], cs.A[tid * FH_MAXM + 2], cs.A[tid * FH_MAXM + 3], cs.A[tid * FH_MAXM + 4], cs.A[tid * FH_MAXM + 5]);
            printf("coef[%d] = %.17g  c1 %.17g zt %.17g zds %.17g zdps %.17g zs %.17g\n", tid, U.coef[tid], c1[tid], zt[tid], zds[tid], zdps[tid], zs[tid]);
#endif
          }
        }
      }
      fh_grid_sync(U.grid_counter, bar_target, nblk);        // barrier 3
      fh_grid_sync(U.grid_counter, bar_target, nblk);        // barrier 4
      return;
    }
    // space update
    fh_grid_sync(U.grid_counter, bar_target, nblk);          // barrier 1
    fh_grid_sync(U.grid_counter, bar_target, nblk);          // barrier 2
    fh_grid_sync(U.grid_counter, bar_target, nblk);          // barrier 3
    if (which != 0) return;
    const int mn = m + 2;
    // dpart layout (space): [0, m): Z de | FH_MAXM: dx.de | FH_MAXM+1: de.g | 2 FH_MAXM + [0, m+2): K row m | 3 FH_MAXM + [0, m+2): K row m+1
    fh_u64_sum_dpart(U.dpart, nstrip, 4 * FH_MAXM, cs.dsum);
    for (int idx = tid; idx < m * m; idx += FH_PASS_THREADS) {
      const int i = idx / m, j = idx - i * m;
      cs.Ks[i * FH_U64_LD + j] = fh_ldcg(U.K + i * FH_U64_LD + j);
      cs.Gzs[i * FH_MAXM + j] = U.Gz[i * FH_MAXM + j];
    }
    __syncthreads();
    if (tid < mn) {
      const double k0 = cs.dsum[2 * FH_MAXM + tid], k1 = cs.dsum[3 * FH_MAXM + tid];
      cs.Ks[m * FH_U64_LD + tid] = k0; cs.Ks[tid * FH_U64_LD + m] = k0;
      cs.Ks[(m + 1) * FH_U64_LD + tid] = k1; cs.Ks[tid * FH_U64_LD + m + 1] = k1;
      cs.Gzs[m * FH_MAXM + tid] = 0.0; cs.Gzs[(m + 1) * FH_MAXM + tid] = 0.0;
      cs.Gzs[tid * FH_MAXM + m] = 0.0; cs.Gzs[tid * FH_MAXM + m + 1] = 0.0;
    }
    __syncthreads();
    if (tid == 0) {
      // K[m][m+1] was summed twice (as row m, column m+1 and as row m+1, column m): both are the same fixed-order sum
      const double gamma = 1.0 / cs.dsum[FH_MAXM];          // 1 / dx.de   (no guard on its sign, like the reference)
      const double deg = cs.dsum[FH_MAXM + 1];
      cs.Gzs[m * FH_MAXM + m] = gamma + gamma * gamma * deg;
      cs.Gzs[m * FH_MAXM + m + 1] = -gamma;
      cs.Gzs[(m + 1) * FH_MAXM + m] = -gamma;
    }
    __syncthreads();
    for (int idx = tid; idx < mn * mn; idx += FH_PASS_THREADS) {
      const int i = idx / mn, j = idx - i * mn;
      if (i >= m || j >= m) U.Gz[i * FH_MAXM + j] = cs.Gzs[i * FH_MAXM + j];
    }
    const int bad = fh_u64_core(cs.Ks, cs.Gzs, mn, cs.A, cs.Ai, cs.red, cs.piv);
    if (bad && tid == 0) atomicAdd(U.status, 1);
    for (int idx = tid; idx < mn * mn; idx += FH_PASS_THREADS) {
      const int i = idx / mn, j = idx - i * mn;
      U.Gs[i * FH_MAXM + j] = cs.A[i * FH_MAXM + j];
    }
    return;
  }

  // =========================================================================================== strip CTAs
  FhRes64Smem& sm = *reinterpret_cast<FhRes64Smem*>(smem_raw);
  float* Zs = reinterpret_cast<float*>(smem_raw + sizeof(FhRes64Smem));
  double* rows_a = reinterpret_cast<double*>(sm.X);
  double* rows_b = reinterpret_cast<double*>(sm.Y);
  float* wbuf = reinterpret_cast<float*>(sm.Y);              // Gram weights (2 x 1536 floats) while no transform is in flight
  double* s_w = reinterpret_cast<double*>(sm.Y);             // scratch of the dot products (8 x 3 x 64 doubles = 12 KB of Y's 15 KB)
  double* s_bs = s_w + (FH_PASS_THREADS / 32) * 3 * FH_MAXM; // 40 doubles behind it: scratch of fh_block_sum
  const int mr = min(m, U.mr_cap);
  const int pl = s / FH_R64_CPP, cc = s - pl * FH_R64_CPP;
  const int prow0 = cc * FH_RES_ROWS;
  const int nrows = min(FH_RES_ROWS, FH_N - prow0), npair = nrows >> 1;
  const int grow0 = pl * FH_N + prow0;
  const size_t soff = (size_t)grow0 * FH_N;
  const float* zsrc = U.Z + soff;
  double* dp = U.dpart + (size_t)s * FH_U64_DP;

  if (tid == 0) { fh_mbar_init(&sm.bar, 1); fh_mbar_fence_init(); }
  __syncthreads();
  if (mr > 0 && warp == 0) {
    if (lane == 0) fh_mbar_expect_tx(&sm.bar, (unsigned)mr * (unsigned)nrows * FH_N * 4u);
    __syncwarp();
    for (int i = lane; i < mr; i += 32) fh_tma_load_1d(Zs + (size_t)i * FH_RES_SLICE, zsrc + (size_t)i * d, (unsigned)nrows * FH_N * 4u, &sm.bar);
  }
  float pif[FH_RES_ROWS];
  fh_r64_getf(U.pi + soff, pif, npair);

  if (U.mode == 1) {
    // ------------------------------------------------------------------ P0
    float dS[FH_RES_ROWS], dH[FH_RES_ROWS];
    double dHd[FH_RES_ROWS];                                 // the transfer's own diagonal in float64 (pi is exact data; |x| >> |mu'| amplifies its rounding)
    {
      const float tauS = (float)(1.0 / U.sig2_new), tauH = (float)(1.0 / U.var2_new);
      const double tauHd = 1.0 / U.var2_new;
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) {
        dS[e] = 1.0f / (pif[e] + tauS);                      // fh_k_delta's expression: the delta the CG reads
        dH[e] = U.consistent ? dS[e] : 1.0f / (pif[e] + tauH);
        dHd[e] = 1.0 / ((double)pif[e] + tauHd);
      }
      fh_u64_storef(U.delta + soff, dS, npair);
    }
    double xs[FH_RES_ROWS], sv[FH_RES_ROWS];
    fh_r64_load(U.x_prev + soff, xs, npair);
    if (U.mean_prev) {
      double mp[FH_RES_ROWS];
      fh_r64_load(U.mean_prev + soff, mp, npair);
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) sv[e] = (mp[e] - xs[e]) / U.sig2_t;          // conditioning_mechanisms.py:243-244
    } else {
      fh_r64_load(U.score_in + soff, sv, npair);
    }
    fh_r64_put(rows_a, sv);
    __syncthreads();
    fh_dct64_core<false, false>(rows_a, rows_b, sm.X, sm.Y, U.tw, U.e512, npair);
    __syncthreads();
    fh_r64_store_t(rows_b, U.T[0], grow0, npair);
    if (m > 0) {
      __syncthreads();
      fh_u64_storef(wbuf, dH, npair);
      if (!U.consistent) fh_u64_storef(wbuf + FH_RES_SLICE, dS, npair);
      if (mr > 0) fh_mbar_wait(&sm.bar, 0);
      __syncthreads();
      if (U.consistent) fh_u64_gram<false>(Zs, zsrc, d, m, mr, nrows, wbuf, U.gpart, s, nstrip);
      else fh_u64_gram<true>(Zs, zsrc, d, m, mr, nrows, wbuf, U.gpart, s, nstrip);
    }
    fh_grid_sync(U.grid_counter, bar_target, nblk);          // barrier 1
    // ------------------------------------------------------------------ P1
    double sh[FH_RES_ROWS];
    {
      double t[FH_RES_ROWS];
      fh_r64_load_plane(U.T[0] + soff, t, npair);
      fh_r64_put(rows_a, t);
      __syncthreads();
      fh_dct64_core<false, false>(rows_a, rows_b, sm.X, sm.Y, U.tw, U.e512, npair);
      __syncthreads();
      fh_r64_get(rows_b, sh, npair);
      __syncthreads();
    }
    double out[FH_RES_ROWS];
    if (m > 0) {
      fh_u64_gram_reduce(U.gpart, U.K, m, U.consistent ? 1 : 2, s, nstrip);
      double vec[3][FH_RES_ROWS];
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) {
        vec[0][e] = dHd[e] * sh[e];                          // -> Z (delta s)
        vec[1][e] = vec[0][e] * (double)pif[e];              // -> Z (delta pi s)
        vec[2][e] = sh[e];                                   // -> Z s
      }
      fh_u64_dots<3>(Zs, zsrc, d, m, mr, npair, vec, s_w, dp);
      fh_grid_sync(U.grid_counter, bar_target, nblk);        // barrier 2
      fh_grid_sync(U.grid_counter, bar_target, nblk);        // barrier 3
      // ---------------------------------------------------------------- P3
      if (tid < m) s_w[tid] = fh_ldcg(U.coef + tid);
      __syncthreads();
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) out[e] = sh[e] * fma(U.var2_old, (double)pif[e], 1.0);
      for (int i = 0; i < m; ++i) {
        float z[FH_RES_ROWS];
        fh_u64_zrow(Zs, zsrc, d, mr, i, z, npair);
        const double c = s_w[i];
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) out[e] = fma(c, FH_WIDEN(z[e], e), out[e]);
      }
      __syncthreads();
    } else {
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) out[e] = sh[e] * fma(U.var2_old, (double)pif[e], 1.0);
    }
    {
      const double inv_var2 = 1.0 / U.var2_new;
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) out[e] *= dHd[e] * inv_var2;
    }
    fh_r64_put(rows_a, out);
    __syncthreads();
    fh_dct64_core<true, false>(rows_a, rows_b, sm.X, sm.Y, U.tw, U.e512, npair);
    __syncthreads();
    fh_r64_store_t(rows_b, U.T[1], grow0, npair);
    fh_grid_sync(U.grid_counter, bar_target, nblk);          // barrier 4
    // ------------------------------------------------------------------ P4
    {
      double t[FH_RES_ROWS];
      fh_r64_load_plane(U.T[1] + soff, t, npair);
      fh_r64_put(rows_a, t);
      __syncthreads();
      fh_dct64_core<true, false>(rows_a, rows_b, sm.X, sm.Y, U.tw, U.e512, npair);
      __syncthreads();
      fh_r64_get(rows_b, t, npair);
      double mu[FH_RES_ROWS];
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) mu[e] = fma(U.sig2_new, t[e], xs[e]);          // mu' = x + sigma'^2 s'
      fh_r64_store(U.score_out + soff, t, npair);
      fh_r64_store(U.mean_out + soff, mu, npair);
    }
    return;
  }

  // ============================================================================================ space update
  float dl[FH_RES_ROWS];
  fh_r64_getf(U.delta + soff, dl, npair);
  {
    double a[FH_RES_ROWS], b[FH_RES_ROWS], v[FH_RES_ROWS];
    fh_r64_load(U.xnext + soff, a, npair);
    fh_r64_load(U.x + soff, b, npair);
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) v[e] = a[e] - b[e];
    fh_r64_put(rows_a, v);
    __syncthreads();
    fh_dct64_core<false, false>(rows_a, rows_b, sm.X, sm.Y, U.tw, U.e512, npair);
    __syncthreads();
    fh_r64_store_t(rows_b, U.T[0], grow0, npair);
    fh_r64_load(U.mean_at_xnext + soff, a, npair);
    fh_r64_load(U.mean_at_x + soff, b, npair);
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) v[e] = U.sig2 * (a[e] - b[e]);
    __syncthreads();
    fh_r64_put(rows_a, v);
    __syncthreads();
    fh_dct64_core<false, false>(rows_a, rows_b, sm.X, sm.Y, U.tw, U.e512, npair);
    __syncthreads();
    fh_r64_store_t(rows_b, U.T[1], grow0, npair);
  }
  if (m > 0) {
    __syncthreads();
    fh_u64_storef(wbuf, dl, npair);
    if (mr > 0) fh_mbar_wait(&sm.bar, 0);
    __syncthreads();
    fh_u64_gram<false>(Zs, zsrc, d, m, mr, nrows, wbuf, U.gpart, s, nstrip);
  }
  fh_grid_sync(U.grid_counter, bar_target, nblk);            // barrier 1
  // -------------------------------------------------------------------- P1
  float dxf[FH_RES_ROWS], def[FH_RES_ROWS];
  {
    double t[FH_RES_ROWS];
    fh_r64_load_plane(U.T[0] + soff, t, npair);
    fh_r64_put(rows_a, t);
    __syncthreads();
    fh_dct64_core<false, false>(rows_a, rows_b, sm.X, sm.Y, U.tw, U.e512, npair);
    __syncthreads();
    fh_r64_get(rows_b, t, npair);
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) dxf[e] = (float)t[e];
    fh_r64_load_plane(U.T[1] + soff, t, npair);
    __syncthreads();
    fh_r64_put(rows_a, t);
    __syncthreads();
    fh_dct64_core<false, false>(rows_a, rows_b, sm.X, sm.Y, U.tw, U.e512, npair);
    __syncthreads();
    fh_r64_get(rows_b, t, npair);
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) def[e] = (float)t[e];
    __syncthreads();
  }
  if (m > 0) fh_u64_gram_reduce(U.gpart, U.K, m, 1, s, nstrip);
  double vec1[1][FH_RES_ROWS];
  {
    double dd = 0.0;
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) { vec1[0][e] = (double)def[e]; dd = fma((double)dxf[e], (double)def[e], dd); }
    if (m > 0) fh_u64_dots<1>(Zs, zsrc, d, m, mr, npair, vec1, s_w, dp);
    const double tot = fh_block_sum(dd, s_bs);
    if (tid == 0) dp[FH_MAXM] = tot;
  }
  fh_grid_sync(U.grid_counter, bar_target, nblk);            // barrier 2
  // -------------------------------------------------------------------- P2: g = de / delta + Z^T Gz (Z de)
  float gf[FH_RES_ROWS];
  {
    double g[FH_RES_ROWS];
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) g[e] = (dl[e] != 0.f) ? (double)def[e] / (double)dl[e] : 0.0;
    if (m > 0) {
      double* zde = s_w;                                     // [FH_MAXM] totals, then [FH_MAXM] coef
      double* cf = s_w + FH_MAXM;
      for (int i = warp; i < m; i += FH_PASS_THREADS / 32) { // one warp per row: sum over the strips, fixed order
        double t = 0.0;
        for (int p = lane; p < nstrip; p += 32) t += fh_ldcg(U.dpart + (size_t)p * FH_U64_DP + i);
        t = fh_warp_sum(t);
        if (lane == 0) zde[i] = t;
      }
      __syncthreads();
      if (tid < m) {
        double t = 0.0;
        for (int k = 0; k < m; ++k) t = fma(U.Gz[tid * FH_MAXM + k], zde[k], t);
        cf[tid] = t;
      }
      __syncthreads();
      for (int i = 0; i < m; ++i) {
        float z[FH_RES_ROWS];
        fh_u64_zrow(Zs, zsrc, d, mr, i, z, npair);
        const double c = cf[i];
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) g[e] = fma(c, FH_WIDEN(z[e], e), g[e]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) gf[e] = (float)g[e];
  }
  fh_u64_storef(U.Z + (size_t)m * d + soff, dxf, npair);
  fh_u64_storef(U.Z + (size_t)(m + 1) * d + soff, gf, npair);
  {
    // Gram rows of the two new rows (weights delta), float64: K[m][i], K[m+1][i] for i < m, the 2 x 2 corner, and de.g
    double vec2[2][FH_RES_ROWS];
    double k00 = 0.0, k01 = 0.0, k11 = 0.0, deg = 0.0;
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) {
      vec2[0][e] = (double)dl[e] * (double)dxf[e];
      vec2[1][e] = (double)dl[e] * (double)gf[e];
      k00 = fma(vec2[0][e], (double)dxf[e], k00);
      k01 = fma(vec2[0][e], (double)gf[e], k01);
      k11 = fma(vec2[1][e], (double)gf[e], k11);
      deg = fma((double)def[e], (double)gf[e], deg);
    }
    if (m > 0) fh_u64_dots<2>(Zs, zsrc, d, m, mr, npair, vec2, s_w, dp + 2 * FH_MAXM);
    k00 = fh_block_sum(k00, s_bs);
    k01 = fh_block_sum(k01, s_bs);
    k11 = fh_block_sum(k11, s_bs);
    deg = fh_block_sum(deg, s_bs);
    if (tid == 0) {
      dp[2 * FH_MAXM + m] = k00; dp[2 * FH_MAXM + m + 1] = k01;
      dp[3 * FH_MAXM + m] = k01; dp[3 * FH_MAXM + m + 1] = k11;
      dp[FH_MAXM + 1] = deg;
    }
  }
  fh_grid_sync(U.grid_counter, bar_target, nblk);            // barrier 3
}
