// Device kernels of the Free Hunch hot path other than the DCT: low-rank Sigma.v pieces, weighted Gram,
// small fp64 algebra, measurement operators, CG vector updates with fused dots.
// All image vectors are (B, d) float32; blockIdx.y indexes the image.  Reductions are deterministic:
// per-block partial sums in fp64, summed in a fixed order by whoever consumes them.
#pragma once
#include "fh_common.cuh"

#define FH_EW_THREADS 256
#define FH_EW_PER_THREAD 8
#define FH_CHUNK (FH_EW_THREADS * FH_EW_PER_THREAD)   // elements of d per CTA

__host__ __device__ inline int fh_nchunks(long d) { return (int)((d + FH_CHUNK - 1) / FH_CHUNK); }

// Load/store 8 consecutive-per-thread-pair elements: thread owns float4 #tid and #tid+256 of the chunk.
// Vector path needs 16-B aligned pointers (true for cudaMalloc'd planes with d % 4 == 0).
__device__ __forceinline__ void fh_load8(const float* __restrict__ base, long off, long d, bool vec, float (&r)[8]) {
  const int tid = threadIdx.x;
  if (vec && off + FH_CHUNK <= d) {
    const float4* p = reinterpret_cast<const float4*>(base + off);
    float4 a = p[tid], b = p[tid + FH_EW_THREADS];
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        long i = off + 4 * (tid + h * FH_EW_THREADS) + e;
        r[4 * h + e] = (i < d) ? base[i] : 0.f;
      }
  }
}
__device__ __forceinline__ void fh_store8(float* __restrict__ base, long off, long d, bool vec, const float (&r)[8]) {
  const int tid = threadIdx.x;
  if (vec && off + FH_CHUNK <= d) {
    float4* p = reinterpret_cast<float4*>(base + off);
    p[tid] = make_float4(r[0], r[1], r[2], r[3]);
    p[tid + FH_EW_THREADS] = make_float4(r[4], r[5], r[6], r[7]);
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        long i = off + 4 * (tid + h * FH_EW_THREADS) + e;
        if (i < d) base[i] = r[4 * h + e];
      }
  }
}

// ------------------------------------------------------------------------------------------------
// delta = 1 / (pi + tau)      (diag of Sigma at the state's sigma)
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_delta(const float* __restrict__ pi, float* __restrict__ delta, float tau, long d, int vec) {
  const long off = (long)blockIdx.x * FH_CHUNK;
  const size_t b = (size_t)blockIdx.y * d;
  float p[8], o[8];
  fh_load8(pi + b, off, d, vec, p);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 1.0f / (p[e] + tau);
  fh_store8(delta + b, off, d, vec, o);
}

// pi = 1/v - tau0   (fp64 evaluation, fp32 storage);  v shared across the batch
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_init_pi(const float* __restrict__ var, float* __restrict__ pi, double tau0, long d) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d) pi[(size_t)blockIdx.y * d + i] = (float)(1.0 / (double)var[i] - tau0);
}

// out = a*x + b*y (elementwise; y may be null -> a*x)
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_axpby(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, float a, float bcoef,
           long d, int vec) {
  const long off = (long)blockIdx.x * FH_CHUNK;
  const size_t b = (size_t)blockIdx.y * d;
  float xv[8], yv[8], o[8];
  fh_load8(x + b, off, d, vec, xv);
  if (y) fh_load8(y + b, off, d, vec, yv);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = a * xv[e] + (y ? bcoef * yv[e] : 0.f);
  fh_store8(out + b, off, d, vec, o);
}

// ------------------------------------------------------------------------------------------------
// Partial dots of the m rows of Z with (w .* v):  partial[b][chunk][i] = sum_{j in chunk} Z[b][i][j] w[j] v[j]
// (w may be null).  Extra "rows" xr0/xr1 (may be null) are appended as rows m, m+1 (used for scalar dots).
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_dots(const float* __restrict__ Z, int m, const float* __restrict__ v, const float* __restrict__ w,
          const float* __restrict__ xr0, const float* __restrict__ xr1, double* __restrict__ partial, long d,
          int vec) {
  __shared__ double s_w[FH_EW_THREADS / 32][FH_MAXM + 2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long off = (long)blockIdx.x * FH_CHUNK;
  const size_t b = blockIdx.y;
  float vv[8];
  fh_load8(v + b * d, off, d, vec, vv);
  if (w) {
    float ww[8];
    fh_load8(w + b * d, off, d, vec, ww);
#pragma unroll
    for (int e = 0; e < 8; ++e) vv[e] *= ww[e];
  }
  const int mt = m + (xr0 ? 1 : 0) + (xr1 ? 1 : 0);
  for (int i = 0; i < mt; ++i) {
    const float* row = (i < m) ? Z + (b * FH_MAXM + i) * (size_t)d : ((i == m && xr0) ? xr0 + b * d : xr1 + b * d);
    float z[8];
    fh_load8(row, off, d, vec, z);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc = fmaf(z[e], vv[e], acc);
    double s = fh_warp_sum((double)acc);
    if (lane == 0) s_w[warp][i] = s;
  }
  __syncthreads();
  if (tid < mt) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < FH_EW_THREADS / 32; ++wv) s += s_w[wv][tid];
    partial[(b * gridDim.x + blockIdx.x) * (FH_MAXM + 2) + tid] = s;
  }
}

// coef[b][i] = sum_j G[b][i][j] * c[j],  c[j] = sum_chunks partial[b][chunk][j];  raw sums also kept in
// csum[b][0..m+2).  One CTA per image.
__global__ void __launch_bounds__(128)
fh_k_dots_finalize(const double* __restrict__ partial, int nchunk, int m, int mextra, const double* __restrict__ G,
                   double* __restrict__ csum, double* __restrict__ coef) {
  __shared__ double c[FH_MAXM + 2];
  const size_t b = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid < m + mextra) {
    double s = 0.0;
    for (int k = 0; k < nchunk; ++k) s += partial[(b * nchunk + k) * (FH_MAXM + 2) + tid];
    c[tid] = s;
    csum[b * (FH_MAXM + 2) + tid] = s;
  }
  __syncthreads();
  if (tid < m && G && coef) {
    const double* g = G + (b * FH_MAXM + tid) * FH_MAXM;
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += g[j] * c[j];
    coef[b * FH_MAXM + tid] = s;
  }
}

// out = a .* (cv * v + sum_i coef[i] Z_i)    (a may be null -> 1; v may be null)
// Sigma.v :  a = delta, v = delta.*w_hat/delta... see host code: out = delta .* (w_hat' + Z^T coef) with
// w_hat' = w_hat (the dots were taken against delta.*w_hat).
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_lowrank_apply(const float* __restrict__ Z, int m, const double* __restrict__ coef, const float* __restrict__ a,
                   const float* __restrict__ v, float cv, float* __restrict__ out, long d, int vec) {
  __shared__ float s_c[FH_MAXM];
  const int tid = threadIdx.x;
  const size_t b = blockIdx.y;
  const long off = (long)blockIdx.x * FH_CHUNK;
  if (tid < m) s_c[tid] = (float)coef[b * FH_MAXM + tid];
  __syncthreads();
  float acc[8];
  if (v) {
    fh_load8(v + b * d, off, d, vec, acc);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= cv;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  }
  for (int i = 0; i < m; ++i) {
    float z[8];
    fh_load8(Z + (b * FH_MAXM + i) * (size_t)d, off, d, vec, z);
    const float c = s_c[i];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(c, z[e], acc[e]);
  }
  if (a) {
    float av[8];
    fh_load8(a + b * d, off, d, vec, av);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= av[e];
  }
  fh_store8(out + b * d, off, d, vec, acc);
}

// ------------------------------------------------------------------------------------------------
// Weighted Gram of up to FH_MAXM+4 rows:  K[i][j] = sum_c R_i[c] R_j[c] w[c]   (w null -> 1).
// rows 0..m-1 are Z rows, rows m..m+nx-1 come from `xrows` pointers (each (B,d)).
// grid = (FH_GRAM_CTAS, B); each CTA strides over 128-column tiles and writes one partial matrix.
#define FH_GRAM_CTAS 148        // one column slab per SM (batch 1); partial matrices are summed by fh_k_gram_reduce
#define FH_GRAM_TILE 128
#define FH_GRAM_MAXR (FH_MAXM + 4)          // 68 rows = 17 x 4
#define FH_GRAM_LD (FH_GRAM_MAXR)           // leading dim of the partial / result matrices

struct FhRowPtrs { const float* p[4]; };

__global__ void __launch_bounds__(256)
fh_k_gram(const float* __restrict__ Z, int m, FhRowPtrs xrows, int nx, const float* __restrict__ w,
          double* __restrict__ partial, long d) {
  FH_DYN_SMEM(float, sm);                    // tile[FH_GRAM_MAXR][FH_GRAM_TILE+1] , wt[FH_GRAM_TILE]
  float* tile = sm;
  float* wt = sm + FH_GRAM_MAXR * (FH_GRAM_TILE + 1);
  const int tid = threadIdx.x;
  const size_t b = blockIdx.y;
  const int R = m + nx;
  const int nb = (R + 3) / 4;                // 4x4 register blocks per dimension
  // thread = (column part, block slot): the 128 columns of a tile are split over 4 groups of 64 threads, each thread
  // owns up to 3 of the nb (nb + 1) / 2 <= 153 upper-triangular 4x4 blocks (the first version gave one block to one
  // thread over all 128 columns: at m = 32 only 45 threads of the CTA computed and the kernel took 100 us)
  const int part = tid >> 6, slot = tid & 63;
  const int nblk2 = nb * (nb + 1) / 2;
  int bpi[3], bpj[3];
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    int blk = slot + 64 * h, pi_ = 0;
    bpi[h] = -1; bpj[h] = 0;
    if (blk < nblk2) {
      while (blk >= nb - pi_) { blk -= nb - pi_; ++pi_; }       // row pi_ of the triangle holds nb - pi_ blocks
      bpi[h] = pi_; bpj[h] = pi_ + blk;
    }
  }
  double acc[3][16];
#pragma unroll
  for (int h = 0; h < 3; ++h)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[h][e] = 0.0;
  const long ntile = (d + FH_GRAM_TILE - 1) / FH_GRAM_TILE;
  for (long tIdx = blockIdx.x; tIdx < ntile; tIdx += gridDim.x) {
    const long c0 = tIdx * FH_GRAM_TILE;
    __syncthreads();
    for (int idx = tid; idx < R * FH_GRAM_TILE; idx += 256) {
      const int r = idx / FH_GRAM_TILE, c = idx % FH_GRAM_TILE;
      const float* row = (r < m) ? Z + (b * FH_MAXM + r) * (size_t)d : xrows.p[r - m] + b * d;
      tile[r * (FH_GRAM_TILE + 1) + c] = (c0 + c < d) ? row[c0 + c] : 0.f;
    }
    for (int idx = tid; idx < (4 * nb - R) * FH_GRAM_TILE; idx += 256) {   // zero the padding rows
      const int r = R + idx / FH_GRAM_TILE, c = idx % FH_GRAM_TILE;
      tile[r * (FH_GRAM_TILE + 1) + c] = 0.f;
    }
    if (tid < FH_GRAM_TILE) wt[tid] = (c0 + tid < d) ? (w ? w[b * d + c0 + tid] : 1.f) : 0.f;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 3; ++h) {
      if (bpi[h] >= 0) {
        float a32[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) a32[e] = 0.f;
        const float* ri = tile + (4 * bpi[h]) * (FH_GRAM_TILE + 1);
        const float* rj = tile + (4 * bpj[h]) * (FH_GRAM_TILE + 1);
        const int cb = part * (FH_GRAM_TILE / 4);
#pragma unroll 4
        for (int c = cb; c < cb + FH_GRAM_TILE / 4; ++c) {
          const float wv = wt[c];
          float x[4], y[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) { x[q] = ri[q * (FH_GRAM_TILE + 1) + c] * wv; y[q] = rj[q * (FH_GRAM_TILE + 1) + c]; }
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int s = 0; s < 4; ++s) a32[4 * q + s] = fmaf(x[q], y[s], a32[4 * q + s]);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[h][e] += (double)a32[e];
      }
    }
  }
  // sum the 4 column parts (fixed order) through shared memory, 64 slots x 16 doubles per part and round
  double* red = reinterpret_cast<double*>(sm);
  double* P = partial + (b * gridDim.x + blockIdx.x) * (size_t)(FH_GRAM_LD * FH_GRAM_LD);
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    __syncthreads();
    if (part > 0) {
#pragma unroll
      for (int e = 0; e < 16; ++e) red[((part - 1) * 64 + slot) * 16 + e] = acc[h][e];
    }
    __syncthreads();
    if (part == 0 && bpi[h] >= 0) {
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[h][e] = ((acc[h][e] + red[slot * 16 + e]) + red[(64 + slot) * 16 + e]) + red[(128 + slot) * 16 + e];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int i = 4 * bpi[h] + q, j = 4 * bpj[h] + s;
          if (i < R && j < R) { P[i * FH_GRAM_LD + j] = acc[h][4 * q + s]; P[j * FH_GRAM_LD + i] = acc[h][4 * q + s]; }
        }
    }
  }
}

// K[b] = sum over CTAs of partial (R x R, leading dim FH_GRAM_LD), fixed order.  grid = (B, FH_GRAM_RED_CTAS), block 256:
// each CTA owns a slice of the matrix entries; each thread keeps 4 independent loads in flight per entry.
#define FH_GRAM_RED_CTAS 8
__global__ void __launch_bounds__(256)
fh_k_gram_reduce(const double* __restrict__ partial, int nparts, int R, double* __restrict__ K) {
  const size_t b = blockIdx.x;
  const size_t mat = (size_t)(FH_GRAM_LD * FH_GRAM_LD);
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < R * R; idx += gridDim.y * blockDim.x) {
    const int i = idx / R, j = idx % R;
    const double* src = partial + b * nparts * mat + i * FH_GRAM_LD + j;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int p = 0;
    for (; p + 4 <= nparts; p += 4) {
      s0 += src[(size_t)p * mat]; s1 += src[(size_t)(p + 1) * mat]; s2 += src[(size_t)(p + 2) * mat]; s3 += src[(size_t)(p + 3) * mat];
    }
    for (; p < nparts; ++p) s0 += src[(size_t)p * mat];
    K[b * mat + i * FH_GRAM_LD + j] = (s0 + s1) + (s2 + s3);
  }
}

// Change of the low-rank basis, in place:  Z'[i][e] = sum_j T[j][i] Z[j][e]   (T: (B, m, m_new) row-major, float64).
// Every thread owns whole columns e, so reading all m rows of a column before writing its m_new rows is race-free;
// float64 accumulation, float32 storage like every row of Z.  grid = (chunks, B), block 256.  Rare (fh_state_recombine).
__global__ void __launch_bounds__(256)
fh_k_recombine(float* __restrict__ Z, int m, int m_new, const double* __restrict__ T, long d) {
  __shared__ double sT[FH_MAXM * FH_MAXM];
  const size_t b = blockIdx.y;
  for (int i = threadIdx.x; i < m * m_new; i += blockDim.x) sT[i] = T[b * (size_t)m * m_new + i];
  __syncthreads();
  float* Zb = Z + b * FH_MAXM * d;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < d; e += (long)gridDim.x * blockDim.x) {
    float z[FH_MAXM];
    for (int j = 0; j < m; ++j) z[j] = Zb[(size_t)j * d + e];
    for (int i = 0; i < m_new; ++i) {
      double a = 0.0;
      for (int j = 0; j < m; ++j) a = fma(sT[j * m_new + i], (double)z[j], a);
      Zb[(size_t)i * d + e] = (float)a;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp64 small algebra, one CTA (256 threads) per image, all matrices <= 64 x 64 in shared memory.
// Gauss-Jordan inverse with partial pivoting of the n x n matrix A (row-major, ld = FH_MAXM) into Ainv.
// Returns (in every thread) 0 on success, 1 if a pivot underflowed.
__device__ inline int fh_invert(double* A, double* Ainv, int n, double* s_red, int* s_piv) {
  // Three barriers per column (the first version had six and let thread 0 scan the column for the pivot alone, which was
  // half of fh_k_core's time): warp 0 finds the pivot with a shuffle arg-max (first row among equals, like a serial
  // scan), then the old pivot row / row c and the elimination factors are staged, then every element is updated from
  // the staged copies.  Same arithmetic, same order of operations per element as the plain algorithm.
  __shared__ double s_rowp[2 * FH_MAXM], s_rowc[2 * FH_MAXM];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int idx = tid; idx < n * n; idx += nt) Ainv[(idx / n) * FH_MAXM + (idx % n)] = ((idx / n) == (idx % n)) ? 1.0 : 0.0;
  __syncthreads();
  int bad = 0;
  for (int c = 0; c < n; ++c) {
    if (tid < 32) {
      double best = -1.0; int p = c;
      for (int r = c + tid; r < n; r += 32) { const double v = fabs(A[r * FH_MAXM + c]); if (v > best) { best = v; p = r; } }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int op = __shfl_xor_sync(0xffffffffu, p, o);
        if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
      }
      if (tid == 0) { s_piv[0] = p; s_piv[1] = (best < 1e-300) ? 1 : 0; }
    }
    __syncthreads();
    const int p = s_piv[0];
    bad |= s_piv[1];
    const double inv = 1.0 / A[p * FH_MAXM + c];
    for (int j = tid; j < 2 * n; j += nt) {
      const double* M = (j < n) ? A : Ainv; const int jj = (j < n) ? j : j - n;
      s_rowp[j] = M[p * FH_MAXM + jj] * inv;               // the new row c
      s_rowc[j] = M[c * FH_MAXM + jj];                     // what row p receives (p != c)
    }
    // elimination factors of the matrix AFTER the row swap: row p holds the old row c
    for (int r = tid; r < n; r += nt) s_red[r] = (r == c) ? 0.0 : A[((r == p) ? c : r) * FH_MAXM + c];
    __syncthreads();
    {
      // thread = (column j of [A | Ainv], row group): the staged pivot-row entry stays in a register while the thread
      // walks down its rows -- no division per element, conflict-free shared-memory columns
      const int per = (2 * n <= 64) ? 64 : 128, j = tid & (per - 1), g = tid / per, ng = nt / per;
      if (j < 2 * n) {
        double* M = (j < n) ? A : Ainv; const int jj = (j < n) ? j : j - n;
        const double rp = s_rowp[j], rc = s_rowc[j];
        for (int r = g; r < n; r += ng) {
          const double f = s_red[r];
          const double cur = (r == p) ? rc : M[r * FH_MAXM + jj];
          M[r * FH_MAXM + jj] = (r == c) ? rp : ((f != 0.0) ? cur - f * rp : cur);
        }
      }
    }
    __syncthreads();
  }
  return bad;
}

// Core of Sigma at the current sigma:  Gs = -Gz (I + K Gz)^-1   (symmetrised).  K has ld FH_GRAM_LD.
// If tu != null also evaluates the time-update coefficients (see tests/algorithm_model.py::time_update):
//   c1 = Gz zs;  zt = zds + var2_old*zdps + var2_old*K c1;  c2 = Gs zt;  coef = var2_old*c1 + c2
// where zs = Z s, zds = Z(delta s), zdps = Z(delta pi s) are columns m+2, m, m+1 of the weighted Gram whose
// extra rows were [s, pi.*s, s./delta].
#define FH_CORE_THREADS 1024          // one CTA per image: latency-bound column steps, so as many row groups as a CTA can have
__global__ void __launch_bounds__(FH_CORE_THREADS)
fh_k_core(const double* __restrict__ K, const double* __restrict__ Gz, int m, double* __restrict__ Gs,
          int time_update, double var2_old, double* __restrict__ coef, int* __restrict__ status) {
  FH_DYN_SMEM(double, sm);
  double* A = sm;                              // FH_MAXM*FH_MAXM
  double* Ai = sm + FH_MAXM * FH_MAXM;         // FH_MAXM*FH_MAXM
  double* red = Ai + FH_MAXM * FH_MAXM;        // FH_MAXM*3
  __shared__ int piv[2];
  const size_t b = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double* Kb = K + b * (size_t)(FH_GRAM_LD * FH_GRAM_LD);
  const double* Gzb = Gz + b * (size_t)(FH_MAXM * FH_MAXM);
  double* Gsb = Gs + b * (size_t)(FH_MAXM * FH_MAXM);
  if (m == 0) return;
  // A = I + K Gz.  Thread = (column j, row group of 4 rows at a time): the Gz entry of a k is loaded once for four
  // rows, the four K entries are warp-uniform; no division per element (the first version: one element per thread and
  // iteration with two dependent global loads per FMA, ~12 us per product at m = 56).
  {
    const int j = tid & (FH_MAXM - 1), g = tid / FH_MAXM, ng = nt / FH_MAXM;
    for (int i0 = 4 * g; i0 < m; i0 += 4 * ng) {
      double s4[4] = {0.0, 0.0, 0.0, 0.0};
      if (j < m) {
        for (int k = 0; k < m; ++k) {
          const double gz = Gzb[k * FH_MAXM + j];
#pragma unroll
          for (int q = 0; q < 4; ++q) s4[q] += Kb[min(i0 + q, m - 1) * FH_GRAM_LD + k] * gz;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i0 + q < m) A[(i0 + q) * FH_MAXM + j] = s4[q] + ((i0 + q == j) ? 1.0 : 0.0);
      }
    }
  }
  __syncthreads();
  const int bad = fh_invert(A, Ai, m, red, piv);
  if (bad && tid == 0) atomicAdd(status, 1);
  // Gs = -Gz * Ai   (into A, then symmetrise into global)
  {
    const int j = tid & (FH_MAXM - 1), g = tid / FH_MAXM, ng = nt / FH_MAXM;
    for (int i0 = 4 * g; i0 < m; i0 += 4 * ng) {
      double s4[4] = {0.0, 0.0, 0.0, 0.0};
      if (j < m) {
        for (int k = 0; k < m; ++k) {
          const double ai = Ai[k * FH_MAXM + j];
#pragma unroll
          for (int q = 0; q < 4; ++q) s4[q] += Gzb[min(i0 + q, m - 1) * FH_MAXM + k] * ai;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i0 + q < m) A[(i0 + q) * FH_MAXM + j] = -s4[q];
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < m * m; idx += nt) {
    const int i = idx / m, j = idx % m;
    Gsb[i * FH_MAXM + j] = 0.5 * (A[i * FH_MAXM + j] + A[j * FH_MAXM + i]);
  }
  if (time_update) {
    __syncthreads();
    double* c1 = red;
    double* zt = red + FH_MAXM;
    if (tid < m) {
      double s = 0.0;
      for (int k = 0; k < m; ++k) s += Gzb[tid * FH_MAXM + k] * Kb[k * FH_GRAM_LD + (m + 2)];
      c1[tid] = s;
    }
    __syncthreads();
    if (tid < m) {
      double s = Kb[tid * FH_GRAM_LD + m] + var2_old * Kb[tid * FH_GRAM_LD + (m + 1)];
      double kc = 0.0;
      for (int k = 0; k < m; ++k) kc += Kb[tid * FH_GRAM_LD + k] * c1[k];
      zt[tid] = s + var2_old * kc;
    }
    __syncthreads();
    if (tid < m) {
      double s = 0.0;
      for (int k = 0; k < m; ++k) s += 0.5 * (A[tid * FH_MAXM + k] + A[k * FH_MAXM + tid]) * zt[k];
      coef[b * FH_MAXM + tid] = var2_old * c1[tid] + s;
    }
  }
}

// Precision core from a covariance-form low rank:  Gz = -J (I + Kp J)^-1  (J = diag(+1 x ku, -1 x kv)),
// used by fh_state_set_cov_vectors.  Kp = W P W^T with ld FH_GRAM_LD.
__global__ void __launch_bounds__(256)
fh_k_core_from_cov(const double* __restrict__ Kp, int ku, int kv, double* __restrict__ Gz, int* __restrict__ status) {
  FH_DYN_SMEM(double, sm);
  double* A = sm;
  double* Ai = sm + FH_MAXM * FH_MAXM;
  double* red = Ai + FH_MAXM * FH_MAXM;
  __shared__ int piv[2];
  const size_t b = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x, m = ku + kv;
  const double* Kb = Kp + b * (size_t)(FH_GRAM_LD * FH_GRAM_LD);
  double* Gzb = Gz + b * (size_t)(FH_MAXM * FH_MAXM);
  if (m == 0) return;
  for (int idx = tid; idx < m * m; idx += nt) {
    const int i = idx / m, j = idx % m;
    const double Jj = (j < ku) ? 1.0 : -1.0;
    A[i * FH_MAXM + j] = ((i == j) ? 1.0 : 0.0) + Kb[i * FH_GRAM_LD + j] * Jj;
  }
  __syncthreads();
  const int bad = fh_invert(A, Ai, m, red, piv);
  if (bad && tid == 0) atomicAdd(status, 1);
  for (int idx = tid; idx < m * m; idx += nt) {
    const int i = idx / m, j = idx % m;
    const double Ji = (i < ku) ? 1.0 : -1.0;
    A[i * FH_MAXM + j] = -Ji * Ai[i * FH_MAXM + j];
  }
  __syncthreads();
  for (int idx = tid; idx < FH_MAXM * FH_MAXM; idx += nt) {
    const int i = idx / FH_MAXM, j = idx % FH_MAXM;
    Gzb[idx] = (i < m && j < m) ? 0.5 * (A[i * FH_MAXM + j] + A[j * FH_MAXM + i]) : 0.0;
  }
}

// BFGS block of the precision core (space update): rows m, m+1 are [dx, g = Sigma^-1 de];
// Gz[m][m] = gamma + gamma^2 de.g ; Gz[m][m+1] = Gz[m+1][m] = -gamma ; gamma = 1/(dx.de)
// (dense twin: online_update_bfgs.py:458).  dx.de = csum[b][m] (extra row of the dots pass),
// de.g = sum of part_deg[b][*].  One thread per image.
__global__ void fh_k_bfgs_block(double* __restrict__ Gz, int m, const double* __restrict__ csum,
                                const double* __restrict__ part_deg, int nchunk, int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  double* G = Gz + (size_t)b * FH_MAXM * FH_MAXM;
  const double gamma = 1.0 / csum[(size_t)b * (FH_MAXM + 2) + m];
  const double deg = fh_sum_partials(part_deg + (size_t)b * nchunk, nchunk);
  for (int j = 0; j < m + 2; ++j) { G[m * FH_MAXM + j] = 0.0; G[(m + 1) * FH_MAXM + j] = 0.0; G[j * FH_MAXM + m] = 0.0; G[j * FH_MAXM + m + 1] = 0.0; }
  G[m * FH_MAXM + m] = gamma + gamma * gamma * deg;
  G[m * FH_MAXM + m + 1] = -gamma;
  G[(m + 1) * FH_MAXM + m] = -gamma;
}

// ------------------------------------------------------------------------------------------------
// Time update elementwise pieces (DCT domain).
// rows for the Gram: r0 = s, r1 = pi.*s, r2 = s./delta      (s = score in the state's basis)
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_tu_rows(const float* __restrict__ s, const float* __restrict__ pi, const float* __restrict__ delta,
             float* __restrict__ r1, float* __restrict__ r2, long d, int vec) {
  const long off = (long)blockIdx.x * FH_CHUNK;
  const size_t b = (size_t)blockIdx.y * d;
  float sv[8], pv[8], dv[8], o1[8], o2[8];
  fh_load8(s + b, off, d, vec, sv);
  fh_load8(pi + b, off, d, vec, pv);
  fh_load8(delta + b, off, d, vec, dv);
#pragma unroll
  for (int e = 0; e < 8; ++e) { o1[e] = pv[e] * sv[e]; o2[e] = (dv[e] != 0.f) ? sv[e] / dv[e] : 0.f; }
  fh_store8(r1 + b, off, d, vec, o1);
  fh_store8(r2 + b, off, d, vec, o2);
}
// s' = delta/var2 .* ( s.*(1 + var2_old*pi) + Z^T coef ) ;  mu' = x + sig2n * s'
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_tu_apply(const float* __restrict__ Z, int m, const double* __restrict__ coef, const float* __restrict__ s,
              const float* __restrict__ pi, const float* __restrict__ delta, const float* __restrict__ x,
              float var2_old, float inv_var2, float sig2n, float* __restrict__ s_out, float* __restrict__ mu_out,
              long d, int vec) {
  __shared__ float s_c[FH_MAXM];
  const int tid = threadIdx.x;
  const size_t b = blockIdx.y;
  const long off = (long)blockIdx.x * FH_CHUNK;
  if (tid < m) s_c[tid] = (float)coef[b * FH_MAXM + tid];
  __syncthreads();
  float acc[8], sv[8], pv[8], dv[8], xv[8], mu[8];
  fh_load8(s + b * d, off, d, vec, sv);
  fh_load8(pi + b * d, off, d, vec, pv);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = sv[e] * fmaf(var2_old, pv[e], 1.0f);
  for (int i = 0; i < m; ++i) {
    float z[8];
    fh_load8(Z + (b * FH_MAXM + i) * (size_t)d, off, d, vec, z);
    const float c = s_c[i];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(c, z[e], acc[e]);
  }
  fh_load8(delta + b * d, off, d, vec, dv);
  fh_load8(x + b * d, off, d, vec, xv);
#pragma unroll
  for (int e = 0; e < 8; ++e) { acc[e] *= dv[e] * inv_var2; mu[e] = fmaf(sig2n, acc[e], xv[e]); }
  fh_store8(s_out + b * d, off, d, vec, acc);
  fh_store8(mu_out + b * d, off, d, vec, mu);
}

// Space update elementwise:  dxp = xnext - x ;  dmp = sig2 * (mean_next - mean)
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_su_diffs(const float* __restrict__ x, const float* __restrict__ xn, const float* __restrict__ mu,
              const float* __restrict__ mun, float sig2, float* __restrict__ dx, float* __restrict__ de, long d, int vec) {
  const long off = (long)blockIdx.x * FH_CHUNK;
  const size_t b = (size_t)blockIdx.y * d;
  float a[8], c[8], o[8];
  fh_load8(xn + b, off, d, vec, a);
  fh_load8(x + b, off, d, vec, c);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = a[e] - c[e];
  fh_store8(dx + b, off, d, vec, o);
  fh_load8(mun + b, off, d, vec, a);
  fh_load8(mu + b, off, d, vec, c);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = sig2 * (a[e] - c[e]);
  fh_store8(de + b, off, d, vec, o);
}
// g = de./delta + Z^T coef   (Sigma^-1 de), written straight into row m+1 of Z; dx copied into row m;
// partial sums of de.g go to part_deg[b][chunk].
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_su_rows(float* __restrict__ Z, int m, const double* __restrict__ coef, const float* __restrict__ de,
             const float* __restrict__ delta, const float* __restrict__ dx, double* __restrict__ part_deg, long d, int vec) {
  __shared__ float s_c[FH_MAXM];
  __shared__ double red[33];
  const int tid = threadIdx.x;
  const size_t b = blockIdx.y;
  const long off = (long)blockIdx.x * FH_CHUNK;
  if (tid < m) s_c[tid] = (float)coef[b * FH_MAXM + tid];
  __syncthreads();
  float acc[8], dv[8], ev[8];
  fh_load8(de + b * d, off, d, vec, ev);
  fh_load8(delta + b * d, off, d, vec, dv);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = (dv[e] != 0.f) ? ev[e] / dv[e] : 0.f;   // padded lanes load delta = 0
  for (int i = 0; i < m; ++i) {
    float z[8];
    fh_load8(Z + (b * FH_MAXM + i) * (size_t)d, off, d, vec, z);
    const float c = s_c[i];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(c, z[e], acc[e]);
  }
  fh_store8(Z + (b * FH_MAXM + m + 1) * (size_t)d, off, d, vec, acc);
  float dg = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) dg = fmaf(ev[e], acc[e], dg);
  const double tot = fh_block_sum((double)dg, red);
  if (tid == 0) part_deg[b * gridDim.x + blockIdx.x] = tot;
  fh_load8(dx + b * d, off, d, vec, ev);
  fh_store8(Z + (b * FH_MAXM + m) * (size_t)d, off, d, vec, ev);
}

// project_to_diagonal (online_update_bfgs.py:274-277): D' = delta + gamma de^2 - (Sigma dx)^2 / (dx.Sigma dx) with
// gamma = 1 / (dx.de); csum[b][0] = dx.Sigma dx, csum[b][1] = dx.de.  Writes pi' = 1/D' - tau, the Gram weight
// delta^2 / D' and the row scale delta / D'.
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_su_project(const float* __restrict__ delta, const float* __restrict__ de, const float* __restrict__ sdx, const double* __restrict__ csum,
                float tau, float* __restrict__ pi, float* __restrict__ wgt, float* __restrict__ scale, long d, int vec) {
  const long off = (long)blockIdx.x * FH_CHUNK;
  const size_t b = (size_t)blockIdx.y;
  const double dsd = csum[b * (FH_MAXM + 2)], gamma = 1.0 / csum[b * (FH_MAXM + 2) + 1];
  float dl[8], ev[8], sv[8], p[8], w[8], sc[8];
  fh_load8(delta + b * d, off, d, vec, dl);
  fh_load8(de + b * d, off, d, vec, ev);
  fh_load8(sdx + b * d, off, d, vec, sv);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const double D = (double)dl[e] + gamma * (double)ev[e] * (double)ev[e] - (double)sv[e] * (double)sv[e] / dsd;
    p[e] = (float)(1.0 / D - (double)tau);
    w[e] = (float)((double)dl[e] * (double)dl[e] / D);
    sc[e] = (float)((double)dl[e] / D);
  }
  fh_store8(pi + b * d, off, d, vec, p);
  fh_store8(wgt + b * d, off, d, vec, w);
  fh_store8(scale + b * d, off, d, vec, sc);
}
// Z[b][i][:] *= scale[b][:]   (grid.z = row i)
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_scale_rows(float* __restrict__ Z, const float* __restrict__ scale, long d, int vec) {
  const long off = (long)blockIdx.x * FH_CHUNK;
  const size_t b = blockIdx.y;
  float* row = Z + (b * FH_MAXM + blockIdx.z) * (size_t)d;
  float z[8], sc[8];
  fh_load8(row, off, d, vec, z);
  fh_load8(scale + b * d, off, d, vec, sc);
#pragma unroll
  for (int e = 0; e < 8; ++e) z[e] *= sc[e];
  fh_store8(row, off, d, vec, z);
}

// ------------------------------------------------------------------------------------------------
// Measurement operators.
// Inpainting: out = mask .* in  (mask (MB,H*W) broadcast over channels; MB = 1 or B)
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_mask_mul(const float* __restrict__ in, const float* __restrict__ mask, int mask_batch, float* __restrict__ out,
              int C, long hw) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // over C*hw
  const size_t b = blockIdx.y;
  if (i < C * hw) {
    const float mk = mask[(mask_batch > 1 ? b * hw : 0) + (i % hw)];
    out[b * C * hw + i] = in[b * C * hw + i] * mk;
  }
}

// General circular convolution with a sparse tap list (dy, dx, w):
//   transpose = 0:  out[i][j] = sum_t w_t in[(i - dy_t) mod H][(j - dx_t) mod W]
//   transpose = 1:  out[i][j] = sum_t w_t in[(i + dy_t) mod H][(j + dx_t) mod W]
// 32x32 output tile per CTA with a halo staged in shared memory.  grid = (W/32, H/32, B*C), block (32,8).
struct FhTaps { const int* dy; const int* dx; const float* w; int n; int ymin, ymax, xmin, xmax; };

#define FH_CONV_MAX_TAPS 1024
__global__ void __launch_bounds__(256)
fh_k_conv_taps(const float* __restrict__ in, float* __restrict__ out, FhTaps taps, int H, int W, int transpose) {
  FH_DYN_SMEM(float, sm);
  __shared__ float2 s_tap[FH_CONV_MAX_TAPS];               // (tap offset inside the staged tile as int bits, weight)
  // halo extents in "input offset" space: offsets o = -dy (fwd) or +dy (transpose)
  const int oy0 = transpose ? taps.ymin : -taps.ymax, oy1 = transpose ? taps.ymax : -taps.ymin;
  const int ox0 = transpose ? taps.xmin : -taps.xmax, ox1 = transpose ? taps.xmax : -taps.xmin;
  const int th = 32 + (oy1 - oy0), tw = 32 + (ox1 - ox0);
  const int tws = tw | 1;                                   // odd stride
  float* tile = sm;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const size_t plane = (size_t)blockIdx.z * H * W;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const bool staged = taps.n <= FH_CONV_MAX_TAPS;
  if (staged)
    for (int t = tid; t < taps.n; t += 256) {
      const int oy = transpose ? taps.dy[t] : -taps.dy[t];
      const int ox = transpose ? taps.dx[t] : -taps.dx[t];
      s_tap[t] = make_float2(__int_as_float((oy - oy0) * tws + (ox - ox0)), taps.w[t]);
    }
  const bool pow2 = ((H & (H - 1)) | (W & (W - 1))) == 0;
  for (int r = threadIdx.y; r < th; r += 8) {               // one warp per tile row: no per-element division
    int gi = i0 + oy0 + r;
    if (pow2) gi &= H - 1; else { gi %= H; if (gi < 0) gi += H; }
    const float* srow = in + plane + (size_t)gi * W;
    for (int c = threadIdx.x; c < tw; c += 32) {
      int gj = j0 + ox0 + c;
      if (pow2) gj &= W - 1; else { gj %= W; if (gj < 0) gj += W; }
      tile[r * tws + c] = srow[gj];
    }
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* base = tile + threadIdx.y * tws + threadIdx.x;
  if (staged) {
#pragma unroll 4
    for (int t = 0; t < taps.n; ++t) {
      const float2 tp = s_tap[t];
      const float wv = tp.y;
      const float* p = base + __float_as_int(tp.x);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(wv, p[8 * q * tws], acc[q]);
    }
  } else {
    for (int t = 0; t < taps.n; ++t) {
      const int oy = transpose ? taps.dy[t] : -taps.dy[t];
      const int ox = transpose ? taps.dx[t] : -taps.dx[t];
      const float wv = taps.w[t];
      const float* p = base + (oy - oy0) * tws + (ox - ox0);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(wv, p[8 * q * tws], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) out[plane + (size_t)(i0 + threadIdx.y + 8 * q) * W + j0 + threadIdx.x] = acc[q];
}

// The same convolution organised by COLUMNS of the PSF: column dx of the PSF is one dense vertical run (dy0 .. dy0+L-1,
// zeros inside kept), and a thread produces 8 consecutive output rows of one image column from a rolling 8-register
// window over the input column: a run costs 8 + L shared-memory loads for 8 L FMAs, where the tap list
// (fh_k_conv_taps) spends one load and ~4 instructions per FMA and is issue-bound (ncu: 85 % issue-active, 22 % of the
// instructions FMAs).  The motion-blur PSF of the reference (217 taps in a 56 x 16 box, measurements.py:126-161) becomes
// 16 runs with 302 taps (332 after padding every run to a multiple of 4 with zero weights, so that the inner loop has
// no per-tap predicate).  64 x 32 output tile per CTA (8 row groups x 32 columns), halo staged in shared memory.
//   run r: weights w[woff + k], k < Lpad, already in the order the window meets them:
//   transpose = 0:  out[i][j] = sum_l psf[dy0 + l][dx] in[i - dy0 - l][j - dx]   (window starts at row i - dy0 - (L-1))
//   transpose = 1:  out[i][j] = sum_l psf[dy0 + l][dx] in[i + dy0 + l][j + dx]   (window starts at row i + dy0)
struct FhRuns { const int4* run; const float* w_fwd; const float* w_tr; int nruns, ntaps; int ymin, ymax, xmin, xmax; };   // run = (dx, dy0, L, woff)
#define FH_CONV_MAX_RUNS 128
#define FH_CONV_PAD_ROWS 4            // the window runs up to 3 zero-weight taps past the run, plus one look-ahead load
// TILE_H = 64 (256 threads) for many planes, 32 (128 threads, twice the CTAs) when a launch has only a few planes.
template <int TILE_H>
__global__ void __launch_bounds__(TILE_H * 4)
fh_k_conv_runs(const float* __restrict__ in, float* __restrict__ out, FhRuns R, int H, int W, int transpose) {
  constexpr int NT = TILE_H * 4, NY = TILE_H / 8;
  FH_DYN_SMEM(float, sm);
  __shared__ int4 s_run[FH_CONV_MAX_RUNS];                 // (first tile row of the window, tile column offset, Lpad, woff)
  __shared__ __align__(16) float s_w[FH_CONV_MAX_TAPS];
  const int oy0 = transpose ? R.ymin : -R.ymax, oy1 = transpose ? R.ymax : -R.ymin;
  const int ox0 = transpose ? R.xmin : -R.xmax, ox1 = transpose ? R.xmax : -R.xmin;
  const int th = TILE_H + (oy1 - oy0) + FH_CONV_PAD_ROWS, tw = 32 + (ox1 - ox0);
  const int tws = tw | 1;
  float* tile = sm;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const size_t plane = (size_t)blockIdx.z * H * W;
  const int i0 = blockIdx.y * TILE_H, j0 = blockIdx.x * 32;
  for (int t = tid; t < R.nruns; t += NT) {
    const int4 r = R.run[t];                               // (dx, dy0, L, woff)
    const int row0 = transpose ? (r.y - oy0) : (-r.y - (r.z - 1) - oy0);
    const int col = (transpose ? r.x : -r.x) - ox0;
    s_run[t] = make_int4(row0, col, (r.z + 3) & ~3, r.w);
  }
  const float* wsrc = transpose ? R.w_tr : R.w_fwd;
  for (int t = tid; t < R.ntaps; t += NT) s_w[t] = wsrc[t];
  const bool pow2 = ((H & (H - 1)) | (W & (W - 1))) == 0;
  for (int r = threadIdx.y; r < th; r += NY) {             // the pad rows are real image rows too (finite: 0 * x = 0)
    int gi = i0 + oy0 + r;
    if (pow2) gi &= H - 1; else { gi %= H; if (gi < 0) gi += H; }
    const float* srow = in + plane + (size_t)gi * W;
    for (int c = threadIdx.x; c < tw; c += 32) {
      int gj = j0 + ox0 + c;
      if (pow2) gj &= W - 1; else { gj %= W; if (gj < 0) gj += W; }
      tile[r * tws + c] = srow[gj];
    }
  }
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const float* base = tile + (threadIdx.y * 8) * tws + threadIdx.x;
#define FH_CONV_TAP(wv, kk)                                                                   \
  do {                                                                                        \
    _Pragma("unroll") for (int e = 0; e < 8; ++e) acc[e] = fmaf((wv), win[(e + (kk)) & 7], acc[e]);  \
    win[(kk)] = q[(kk) * tws];                           /* the window moves down one input row */     \
  } while (0)
  for (int t = 0; t < R.nruns; ++t) {
    const int4 r = s_run[t];
    const float* q = base + r.x * tws + r.y;
    const float4* wq = reinterpret_cast<const float4*>(s_w + r.w);
    float win[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) win[e] = q[e * tws];
    q += 8 * tws;
    for (int k0 = 0; k0 < r.z; k0 += 8) {
      const float4 wa = wq[0];
      FH_CONV_TAP(wa.x, 0); FH_CONV_TAP(wa.y, 1); FH_CONV_TAP(wa.z, 2); FH_CONV_TAP(wa.w, 3);
      if (k0 + 4 >= r.z) break;                            // uniform: runs are padded to multiples of 4
      const float4 wb = wq[1];
      FH_CONV_TAP(wb.x, 4); FH_CONV_TAP(wb.y, 5); FH_CONV_TAP(wb.z, 6); FH_CONV_TAP(wb.w, 7);
      q += 8 * tws; wq += 2;
    }
  }
#undef FH_CONV_TAP
#pragma unroll
  for (int e = 0; e < 8; ++e) out[plane + (size_t)(i0 + threadIdx.y * 8 + e) * W + j0 + threadIdx.x] = acc[e];
}

// Separable circular convolution (rank-1 PSF = col (x) row):  out = conv_cols(kc) conv_rows(kr) in.
// Taps: kr[a] at offset (a - cr), kc[a] at offset (a - cc); transpose flips the sign of the offsets.
struct FhSepTaps { const float* kr; const float* kc; int nr, nc, cr, cc; };

__global__ void __launch_bounds__(256)
fh_k_conv_sep(const float* __restrict__ in, float* __restrict__ out, FhSepTaps k, int H, int W, int transpose) {
  FH_DYN_SMEM(float, sm);
  // forward: out[i][j] = sum_a sum_b kc[a] kr[b] in[i-(a-cc)][j-(b-cr)]
  const int oy0 = transpose ? -k.cc : -(k.nc - 1 - k.cc), oy1 = transpose ? (k.nc - 1 - k.cc) : k.cc;
  const int ox0 = transpose ? -k.cr : -(k.nr - 1 - k.cr), ox1 = transpose ? (k.nr - 1 - k.cr) : k.cr;
  const int th = 32 + (oy1 - oy0), tw = 32 + (ox1 - ox0);
  const int tws = tw | 1;
  float* tile = sm;                       // th x tws
  float* tmp = sm + th * tws;             // th x 33   (row-convolved)
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const size_t plane = (size_t)blockIdx.z * H * W;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int idx = tid; idx < th * tw; idx += 256) {
    const int r = idx / tw, c = idx % tw;
    int gi = (i0 + oy0 + r) % H; if (gi < 0) gi += H;
    int gj = (j0 + ox0 + c) % W; if (gj < 0) gj += W;
    tile[r * tws + c] = in[plane + (size_t)gi * W + gj];
  }
  __syncthreads();
  for (int idx = tid; idx < th * 32; idx += 256) {
    const int r = idx / 32, c = idx % 32;
    float acc = 0.f;
    for (int bb = 0; bb < k.nr; ++bb) {
      const int ox = transpose ? (bb - k.cr) : -(bb - k.cr);
      acc = fmaf(k.kr[bb], tile[r * tws + c + ox - ox0], acc);
    }
    tmp[r * 33 + c] = acc;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = threadIdx.y + 8 * q;
    float acc = 0.f;
    for (int a = 0; a < k.nc; ++a) {
      const int oy = transpose ? (a - k.cc) : -(a - k.cc);
      acc = fmaf(k.kc[a], tmp[(r + oy - oy0) * 33 + threadIdx.x], acc);
    }
    out[plane + (size_t)(i0 + r) * W + j0 + threadIdx.x] = acc;
  }
}

// Super-resolution forward: y[I][J] = sum_t w_t x[(sf*I - dy_t) mod H][(sf*J - dx_t) mod W]  (conv, then [0::sf,0::sf])
__global__ void __launch_bounds__(256)
fh_k_sr_forward(const float* __restrict__ in, float* __restrict__ out, FhTaps taps, int H, int W, int sf) {
  FH_DYN_SMEM(float, sm);                 // taps staged: dy, dx (as int), w
  int* sdy = reinterpret_cast<int*>(sm);
  int* sdx = sdy + taps.n;
  float* sw = sm + 2 * taps.n;
  for (int t = threadIdx.x; t < taps.n; t += blockDim.x) { sdy[t] = taps.dy[t]; sdx[t] = taps.dx[t]; sw[t] = taps.w[t]; }
  __syncthreads();
  const int h = H / sf, w = W / sf;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t p = blockIdx.y;
  if (idx >= h * w) return;
  const int I = idx / w, J = idx % w;
  const float* src = in + p * (size_t)H * W;
  float acc = 0.f;
  for (int t = 0; t < taps.n; ++t) {
    const int gi = (sf * I - sdy[t] + H) & (H - 1), gj = (sf * J - sdx[t] + W) & (W - 1);
    acc = fmaf(sw[t], src[(size_t)gi * W + gj], acc);
  }
  out[p * (size_t)h * w + idx] = acc;
}
// Super-resolution transpose: x[i][j] = sum_{t: (i+dy_t) % sf == 0 and (j+dx_t) % sf == 0} w_t y[(i+dy_t)/sf][(j+dx_t)/sf]
__global__ void __launch_bounds__(256)
fh_k_sr_transpose(const float* __restrict__ in, float* __restrict__ out, FhTaps taps, int H, int W, int sf) {
  FH_DYN_SMEM(float, sm);
  int* sdy = reinterpret_cast<int*>(sm);
  int* sdx = sdy + taps.n;
  float* sw = sm + 2 * taps.n;
  for (int t = threadIdx.x; t < taps.n; t += blockDim.x) { sdy[t] = taps.dy[t]; sdx[t] = taps.dx[t]; sw[t] = taps.w[t]; }
  __syncthreads();
  const int h = H / sf, w = W / sf;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t p = blockIdx.y;
  if (idx >= H * W) return;
  const int i = idx / W, j = idx % W;
  const float* src = in + p * (size_t)h * w;
  float acc = 0.f;
  for (int t = 0; t < taps.n; ++t) {
    const int gi = (i + sdy[t] + H) & (H - 1), gj = (j + sdx[t] + W) & (W - 1);
    if ((gi % sf) == 0 && (gj % sf) == 0) acc = fmaf(sw[t], src[(size_t)(gi / sf) * w + gj / sf], acc);
  }
  out[p * (size_t)H * W + idx] = acc;
}

// ------------------------------------------------------------------------------------------------
// CG vector kernels (y-space vectors of length n per image).  Per-image scalars live in FhCgScalars.

// b = (mask? mask.*)(y - Amu) ;  x = b   (Amu = A mu, or mu itself for inpainting)
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_cg_rhs(const float* __restrict__ y, const float* __restrict__ amu, const float* __restrict__ mask, int mask_batch,
            long hw, float* __restrict__ bvec, float* __restrict__ x, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t b = blockIdx.y;
  if (i < n) {
    float mk = 1.f;
    if (mask) mk = mask[(mask_batch > 1 ? b * hw : 0) + (i % hw)];
    // inpainting: mask*y - mask*mu (conditioning_mechanisms.py:407); otherwise y - A mu (:515, :663)
    const float v = mask ? (mk * y[b * n + i] - mk * amu[b * n + i]) : (y[b * n + i] - amu[b * n + i]);
    bvec[b * n + i] = v;
    x[b * n + i] = v;
  }
}

// q = sy2 * p + (mask? mask.*)t  and partial p.q         (t = A Sigma A^T p computed by the operator chain)
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_cg_q(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ mask, int mask_batch,
          long hw, float sy2, float* __restrict__ q, double* __restrict__ partial, const FhCgScalars* __restrict__ sc,
          long n) {
  __shared__ double red[33];
  const size_t b = blockIdx.y;
  if (sc[b].done) return;
  const long i0 = (long)blockIdx.x * FH_CHUNK;
  double acc = 0.0;
  float a32 = 0.f;
  for (int e = 0; e < FH_EW_PER_THREAD; ++e) {
    const long i = i0 + threadIdx.x + (long)e * FH_EW_THREADS;
    if (i < n) {
      float tv = t[b * n + i];
      if (mask) tv *= mask[(mask_batch > 1 ? b * hw : 0) + (i % hw)];
      const float pv = p[b * n + i];
      const float qv = fmaf(sy2, pv, tv);
      q[b * n + i] = qv;
      a32 = fmaf(pv, qv, a32);
    }
  }
  acc = fh_block_sum((double)a32, red);
  if (threadIdx.x == 0) partial[b * gridDim.x + blockIdx.x] = acc;
}

// init:  r = b - q ; p = r ; partials of r.r and b.b
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_cg_init(const float* __restrict__ bvec, const float* __restrict__ q, float* __restrict__ r, float* __restrict__ p,
             double* __restrict__ part_rr, double* __restrict__ part_bb, long n) {
  __shared__ double red[33];
  const size_t b = blockIdx.y;
  const long i0 = (long)blockIdx.x * FH_CHUNK;
  float rr = 0.f, bb = 0.f;
  for (int e = 0; e < FH_EW_PER_THREAD; ++e) {
    const long i = i0 + threadIdx.x + (long)e * FH_EW_THREADS;
    if (i < n) {
      const float bv = bvec[b * n + i];
      const float rv = bv - q[b * n + i];
      r[b * n + i] = rv;
      p[b * n + i] = rv;
      rr = fmaf(rv, rv, rr);
      bb = fmaf(bv, bv, bb);
    }
  }
  const double s1 = fh_block_sum((double)rr, red);
  const double s2 = fh_block_sum((double)bb, red);
  if (threadIdx.x == 0) { part_rr[b * gridDim.x + blockIdx.x] = s1; part_bb[b * gridDim.x + blockIdx.x] = s2; }
}
// one thread per image: set up the scalars after fh_k_cg_init
__global__ void fh_k_cg_init_scalars(FhCgScalars* sc, const double* part_rr, const double* part_bb, int nblk,
                                     double rtol, double atol, int use_stop, int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const double rr = fh_sum_partials(part_rr + (size_t)b * nblk, nblk);
  const double bb = fh_sum_partials(part_bb + (size_t)b * nblk, nblk);
  FhCgScalars s;
  s.rr[0] = rr; s.rr[1] = rr;
  s.bnorm2 = bb;
  const double stop = fmax(rtol * sqrt(bb), atol);      // cg.py:237
  s.stop2 = use_stop ? stop * stop : -1.0;
  s.resid2 = rr;
  s.alpha = 0.0; s.beta = 0.0;
  s.done = 0; s.done_tag = 0; s.optimal = 0; s.niter = 0;
  sc[b] = s;
}

// update 1 (iteration `it`, 1-based):  pAp = sum partial;  if pAp <= 1e-16: break (cg.py:250)
//   alpha = rr/pAp ; x += alpha p ; r -= alpha q ; partial r.r
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_cg_update1(float* __restrict__ x, float* __restrict__ r, const float* __restrict__ p, const float* __restrict__ q,
                const double* __restrict__ part_pq, double* __restrict__ part_rr, FhCgScalars* __restrict__ sc, int it,
                long n) {
  __shared__ double red[33];
  const size_t b = blockIdx.y;
  if (sc[b].done) return;
  const int nblk = gridDim.x;
  const double pAp = fh_sum_partials(part_pq + b * nblk, nblk);
  if (pAp <= 1e-16) {
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { sc[b].done = 1; sc[b].niter = it; }
    return;
  }
  const float alpha = (float)(sc[b].rr[it & 1] / pAp);
  const long i0 = (long)blockIdx.x * FH_CHUNK;
  float rr = 0.f;
  for (int e = 0; e < FH_EW_PER_THREAD; ++e) {
    const long i = i0 + threadIdx.x + (long)e * FH_EW_THREADS;
    if (i < n) {
      x[b * n + i] = fmaf(alpha, p[b * n + i], x[b * n + i]);
      const float rv = fmaf(-alpha, q[b * n + i], r[b * n + i]);
      r[b * n + i] = rv;
      rr = fmaf(rv, rv, rr);
    }
  }
  const double s = fh_block_sum((double)rr, red);
  if (threadIdx.x == 0) part_rr[b * nblk + blockIdx.x] = s;
}
// update 2:  rr' = sum partial ; if sqrt(rr') <= stop: optimal, break (cg.py:259) ; beta = rr'/rr ; p = r + beta p
// A pAp-break detected by update1 of the same iteration is visible here through sc.done.
__global__ void __launch_bounds__(FH_EW_THREADS)
fh_k_cg_update2(const float* __restrict__ r, float* __restrict__ p, const double* __restrict__ part_rr,
                FhCgScalars* __restrict__ sc, int it, int last_allowed, long n) {
  const size_t b = blockIdx.y;
  if (sc[b].done) return;
  const int nblk = gridDim.x;
  const double rr_new = fh_sum_partials(part_rr + b * nblk, nblk);
  const double rr_old = sc[b].rr[it & 1];
  const bool converged = (sc[b].stop2 >= 0.0) && (rr_new <= sc[b].stop2);
  __syncthreads();     // every thread of this CTA has read sc before thread 0 of CTA 0 may write it
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc[b].resid2 = rr_new;
    sc[b].rr[(it + 1) & 1] = rr_new;
    sc[b].niter = it;
    if (converged) { sc[b].done = 1; sc[b].optimal = 1; }
    else if (it >= last_allowed) { sc[b].done = 1; }
  }
  if (converged) return;
  const float beta = (float)(rr_new / rr_old);
  const long i0 = (long)blockIdx.x * FH_CHUNK;
  for (int e = 0; e < FH_EW_PER_THREAD; ++e) {
    const long i = i0 + threadIdx.x + (long)e * FH_EW_THREADS;
    if (i < n) p[b * n + i] = fmaf(beta, p[b * n + i], r[b * n + i]);
  }
}

// ---- x4 super-resolution, fast path: 256x256 planes, sf = 4, dense ks x ks PSF with ks <= 25, centre c = ks/2 = 12 ----
// (utils_sisr.py:44-61 + the FFT operator of conditioning_mechanisms.py:641-675 evaluated in pixel space)
#define FH_SR_KS 25
// forward: y[I][J] = sum_{a,b} k[a][b] x[(4I - a + 12) & 255][(4J - b + 12) & 255].  One CTA = 4 output rows of a plane:
// the 37 image rows it needs are staged in shared memory de-interleaved by column phase (col = 4q + ph) so that the 64
// threads of an output row read consecutive words.  grid = (16, planes), block 256.
__global__ void __launch_bounds__(256)
fh_k_sr4_forward(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ psf, int ks) {
  __shared__ float xs[37][4 * 65];
  __shared__ float sk[FH_SR_KS * FH_SR_KS];
  const int tid = threadIdx.x;
  const size_t plane = blockIdx.y;
  const int I0 = blockIdx.x * 4;
  const int c = ks / 2;
  for (int i = tid; i < FH_SR_KS * FH_SR_KS; i += 256) {
    const int a = i / FH_SR_KS, b = i % FH_SR_KS;
    sk[i] = (a < ks && b < ks) ? psf[a * ks + b] : 0.f;    // zero-padded to 25 x 25: the tap loops below are fixed-size
  }
  const float* src = in + plane * 65536;
  const int nrows = 4 * 3 + 2 * c + 1;                     // rows 4 I0 - c ... 4 (I0 + 3) + c
  for (int idx = tid; idx < nrows * 64; idx += 256) {
    const int r = idx >> 6, q4 = idx & 63;
    const int gi = (4 * I0 - c + r) & 255;
    const float4 v = *reinterpret_cast<const float4*>(src + gi * 256 + 4 * q4);
    xs[r][q4] = v.x; xs[r][65 + q4] = v.y; xs[r][130 + q4] = v.z; xs[r][195 + q4] = v.w;
  }
  __syncthreads();
  const int Il = tid >> 6, J = tid & 63;
  // column offsets of the 25 taps for this thread's J: identical for every tap row, so computed once
  int offb[FH_SR_KS];
#pragma unroll
  for (int b = 0; b < FH_SR_KS; ++b) {
    const int col = 4 * J - b + c;                         // may be negative or >= 256: wrap
    offb[b] = (col & 3) * 65 + ((col >> 2) & 63);
  }
  float acc = 0.f;
  for (int a = 0; a < ks; ++a) {
    const float* xr = xs[4 * Il + 2 * c - a];              // image row 4 (I0 + Il) - a + c
    const float* kr = sk + a * FH_SR_KS;
#pragma unroll
    for (int b = 0; b < FH_SR_KS; ++b) acc = fmaf(kr[b], xr[offb[b]], acc);
  }
  out[plane * 4096 + (size_t)(I0 + Il) * 64 + J] = acc;
}
// transpose: x[i][j] = sum over the taps with (i + a - 12) % 4 == 0 and (j + b - 12) % 4 == 0 of
// k[a][b] y[((i + a - 12) & 255) / 4][((j + b - 12) & 255) / 4].  One CTA = 16 image rows of a plane; thread = one column.
// grid = (16, planes), block 256.
__global__ void __launch_bounds__(256)
fh_k_sr4_transpose(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ psf, int ks) {
  __shared__ float ys[12][64];
  __shared__ float sk[FH_SR_KS * (FH_SR_KS + 3)];          // rows padded to 28 so b0 + 4 t (t < 7) never leaves the row
  const int tid = threadIdx.x;
  const size_t plane = blockIdx.y;
  const int i0 = blockIdx.x * 16;
  const int c = ks / 2;                                    // multiple of 4 (checked by the host)
  for (int i = tid; i < FH_SR_KS * (FH_SR_KS + 3); i += 256) {
    const int a = i / (FH_SR_KS + 3), b = i % (FH_SR_KS + 3);
    sk[i] = (a < ks && b < ks) ? psf[a * ks + b] : 0.f;
  }
  const float* src = in + plane * 4096;
  const int Ilo = (i0 - c) >> 2;                           // floor; y rows Ilo ... Ilo + 11 cover rows i0-c ... i0+15+c
  for (int idx = tid; idx < 12 * 64; idx += 256) {
    const int r = idx >> 6, q = idx & 63;
    ys[r][q] = src[((Ilo + r) & 63) * 64 + q];
  }
  __syncthreads();
  const int j = tid;
  const int b0 = (4 - (j & 3)) & 3;                        // first tap column with (j + b - c) % 4 == 0
  int qt[7];                                               // y columns of the 7 candidate taps: the same for every row
#pragma unroll
  for (int t = 0; t < 7; ++t) qt[t] = ((j + b0 + 4 * t - c) >> 2) & 63;
  float* dst = out + plane * 65536;
  for (int il = 0; il < 16; ++il) {
    const int i = i0 + il;
    const int a0 = (4 - (i & 3)) & 3;
    float acc = 0.f;
    for (int a = a0; a < ks; a += 4) {
      const float* yr = ys[((i + a - c) >> 2) - Ilo];      // 0 ... 11
      const float* kr = sk + a * (FH_SR_KS + 3) + b0;
#pragma unroll
      for (int t = 0; t < 7; ++t) acc = fmaf(kr[4 * t], yr[qt[t]], acc);   // taps beyond ks are zero
    }
    dst[(size_t)i * 256 + j] = acc;
  }
}
