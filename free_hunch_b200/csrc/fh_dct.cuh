// Orthonormal 256-point DCT-II / DCT-III along rows of 256x256 planes.
//
// Core: two real rows are packed into one 256-point complex FFT (Makhoul reordering).  The FFT is a radix-4
// Stockham autosort transform: 64 threads per FFT, 4 stages, one 4-point butterfly per thread and stage, data
// ping-ponging between two shared-memory buffers.  A CTA of 256 threads owns a strip of ROWS in {2,4,8}
// consecutive rows (ROWS*256 contiguous floats): all threads move data with coalesced vector accesses and do the
// fused element-wise / convolution / dot-product work of the CG kernels in fh_cg.cuh; ROWS/2 groups of 64 threads
// run the FFTs.  Small batches use short strips so that every SM gets work (B=1: 384 strips of 2 rows).
//
// A 2-D transform is two row passes: the first stores its result TRANSPOSED, the second stores it as is.  The
// covariance state therefore lives in the "basis layout"  W[c][kj][ki]  (DCT coefficient (ki,kj) of channel c),
// which no one outside the library sees: fh_dct2 (public) transposes twice and returns the natural layout.
//
// Replaces torch_dct.dct_2d / idct_2d(norm='ortho') (reference call sites
// conditioning_utils/online_update_bfgs.py:352-374; torch-dct 0.1.6 computes the same transform through an FFT
// of the even/odd reordered sequence).
#pragma once
#include "fh_common.cuh"

struct FhDctTables {
  const float2* tw256;   // Stockham stage twiddles in thread order: [stage s = 0..2 (Ns = 4, 16, 64)][r = 1..3][j = 0..63]
                         //   = exp(-2 pi i (j mod Ns) r / (4 Ns))   (576 entries; conjugated for the inverse)
  const float2* post;    // f_k * exp(-i pi k / 512): forward post-twiddle incl. orthonormal scale
  const float2* cw;      // exp(+i pi k / 512): inverse pre-twiddle
};

#define FH_RS 264             // shared-memory row stride in floats (16-B aligned rows)
#define FH_PAD(i) ((i) + (((i) >> 4) << 2))      // 4 float2 of padding after every 16: stage writes 16g+k land on 20g+k
#define FH_FFT_LD 320
#define FH_PASS_THREADS 256

template <int ROWS>
struct FhStripSmem {
  float a[ROWS * FH_RS];
  float b[ROWS * FH_RS];
  float2 X[ROWS / 2][FH_FFT_LD];      // FFT ping-pong buffers, padded (FH_PAD) so that every Stockham stage's
  float2 Y[ROWS / 2][FH_FFT_LD];      // strided writes fall on distinct banks
  float2 tw[3 * 3 * 64];              // per-stage twiddles, laid out so that thread j reads entry j: conflict-free
  float2 tab[FH_N];
};

__device__ __forceinline__ float2 fh_cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 fh_cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 fh_csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by exp(SIGN * i pi/2):  SIGN=-1 -> -i*a = (a.y, -a.x);  SIGN=+1 -> i*a = (-a.y, a.x)
template <int SIGN>
__device__ __forceinline__ float2 fh_rot90(float2 a) {
  return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

// 4-point DFT, W = exp(SIGN*2 pi i/4); natural order in and out.
template <int SIGN>
__device__ __forceinline__ void fh_dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 s02 = fh_cadd(a0, a2), d02 = fh_csub(a0, a2);
  float2 s13 = fh_cadd(a1, a3), d13 = fh_rot90<SIGN>(fh_csub(a1, a3));
  a0 = fh_cadd(s02, s13);
  a2 = fh_csub(s02, s13);
  a1 = fh_cadd(d02, d13);
  a3 = fh_csub(d02, d13);
}

// Load the twiddle tables of one direction into the strip's shared memory (all threads; caller syncs).
template <bool INV, int ROWS>
__device__ __forceinline__ void fh_strip_tables(FhStripSmem<ROWS>& sm, const FhDctTables& tab) {
  for (int i = threadIdx.x; i < 3 * 3 * 64; i += blockDim.x) {
    const float2 w = tab.tw256[i];
    sm.tw[i] = INV ? make_float2(w.x, -w.y) : w;
  }
  for (int i = threadIdx.x; i < FH_N; i += blockDim.x) sm.tab[i] = INV ? tab.cw[i] : tab.post[i];
}

// One Stockham radix-4 stage of a 256-point FFT for thread j in [0, 64): sub-transform length Ns -> 4 Ns.
// CONJ: the table holds the forward twiddles and the inverse transform conjugates them on the fly.
template <int SIGN, int NS, bool CONJ = false>
__device__ __forceinline__ void fh_fft256_stage(const float2* __restrict__ in, float2* __restrict__ out, int j, const float2* __restrict__ tw) {
  const int k = j & (NS - 1);
  float2 v0 = in[FH_PAD(j)], v1 = in[FH_PAD(j + 64)], v2 = in[FH_PAD(j + 128)], v3 = in[FH_PAD(j + 192)];
  if (NS > 1) {
    constexpr int S = (NS == 4) ? 0 : (NS == 16) ? 1 : 2;   // angle = SIGN * 2 pi k r / (4 Ns), tabulated per thread
    float2 w1 = tw[(S * 3 + 0) * 64 + j], w2 = tw[(S * 3 + 1) * 64 + j], w3 = tw[(S * 3 + 2) * 64 + j];
    if (CONJ) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
    v1 = fh_cmul(v1, w1);
    v2 = fh_cmul(v2, w2);
    v3 = fh_cmul(v3, w3);
  }
  fh_dft4<SIGN>(v0, v1, v2, v3);
  const int j0 = ((j - k) << 2) + k;
  out[FH_PAD(j0)] = v0; out[FH_PAD(j0 + NS)] = v1; out[FH_PAD(j0 + 2 * NS)] = v2; out[FH_PAD(j0 + 3 * NS)] = v3;
}

// Transform the ROWS rows in `in` (shared, stride FH_RS) into `out` (shared, stride FH_RS; must not alias `in`).
// COLLECTIVE: every thread of the 256-thread CTA calls it (it contains __syncthreads); `in` must be complete
// (caller synchronised) and `out` is complete only after the caller's next __syncthreads.
// Core with explicit buffers: Xb / Yb are the [ROWS/2][FH_FFT_LD] ping-pong buffers, tw the stage twiddles (forward
// table conjugated on the fly when CONJ), tab the post- (forward) or pre- (inverse) twiddle table.  `in` may alias Xb
// and `out` may alias Yb: `in` is dead after the first barrier (before any write to Xb) and Yb is dead after the last
// one (before any write to `out`) -- the resident CG kernel (fh_cg_resident.cuh) relies on this to save 12 KB.
template <bool INV, int ROWS, bool CONJ>
__device__ __forceinline__ void fh_dct_strip_core(const float* in, float* out, float2 (*Xb)[FH_FFT_LD], float2 (*Yb)[FH_FFT_LD],
                                                  const float2* __restrict__ tw, const float2* __restrict__ tab) {
  constexpr int SIGN = INV ? +1 : -1;
  const int f = threadIdx.x >> 6, j = threadIdx.x & 63;
  const bool act = f < ROWS / 2;
  const int fc = act ? f : 0;
  const float* ra = in + (2 * fc) * FH_RS;
  const float* rb = ra + FH_RS;
  float* oa = out + (2 * fc) * FH_RS;
  float* ob = oa + FH_RS;
  float2* X = Xb[fc];
  float2* Y = Yb[fc];
  if (act) {
    float2 v[4];
    if (!INV) {
      // z[n] = va[n] + i vb[n],  v[n] = a[2n] (n<128) | a[511-2n] (n>=128);  this thread: n = j + 64 r
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = j + 64 * r;
        const int src = (n < 128) ? 2 * n : 511 - 2 * n;
        v[r] = make_float2(ra[src], rb[src]);
      }
    } else {
      // Z[k] = cw[k] * ((Ca[k] + Cb[N-k]) + i (Cb[k] - Ca[N-k])),  C[k] = X[k]*g_k/N, C[N] = 0
      const float sc0 = 0.0625f, sc1 = 0.044194173824159216f;     // 1/sqrt(N), 1/sqrt(2N)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = j + 64 * r;
        const float sk = (k == 0) ? sc0 : sc1;
        const int kn = (FH_N - k) & (FH_N - 1);
        const float sn = (k == 0) ? 0.f : sc1;
        const float A = ra[k] * sk + rb[kn] * sn;
        const float B = rb[k] * sk - ra[kn] * sn;
        v[r] = fh_cmul(tab[k], make_float2(A, B));
      }
    }
    fh_dft4<SIGN>(v[0], v[1], v[2], v[3]);             // stage Ns = 1 (no twiddles): out[4 j + r]
    float4* y4 = reinterpret_cast<float4*>(Y + FH_PAD(4 * j));      // 4 consecutive outputs stay inside one 16-block
    y4[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    y4[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
  }
  __syncthreads();
  if (act) fh_fft256_stage<SIGN, 4, CONJ>(Y, X, j, tw);
  __syncthreads();
  if (act) fh_fft256_stage<SIGN, 16, CONJ>(X, Y, j, tw);
  __syncthreads();
  if (act) fh_fft256_stage<SIGN, 64, CONJ>(Y, X, j, tw);
  __syncthreads();
  if (act) {
    if (!INV) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = j + 64 * r;
        const float2 zk = X[FH_PAD(k)];
        const float2 zn = X[FH_PAD((FH_N - k) & (FH_N - 1))];
        const float2 va = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        const float2 vb = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
        const float2 p = tab[k];
        oa[k] = p.x * va.x - p.y * va.y;
        ob[k] = p.x * vb.x - p.y * vb.y;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = j + 64 * r;
        const int pos = (n < 128) ? 2 * n : 511 - 2 * n;
        const float2 z = X[FH_PAD(n)];
        oa[pos] = z.x;
        ob[pos] = z.y;
      }
    }
  }
}

template <bool INV, int ROWS>
__device__ __forceinline__ void fh_dct_strip(const float* __restrict__ in, float* __restrict__ out, FhStripSmem<ROWS>& sm) {
  fh_dct_strip_core<INV, ROWS, false>(in, out, sm.X, sm.Y, sm.tw, sm.tab);
}

// ---- strip movers (all 256 threads) -------------------------------------------------------------
// A strip is ROWS*256 contiguous floats; thread t owns ROWS of them as one vector per 256*W floats:
//   ROWS = 8: float4 #t and #t+256      ROWS = 4: float4 #t      ROWS = 2: float2 #t      ROWS = 6: float2 #t, #t+256, #t+512
template <int ROWS> struct FhVec { static constexpr int W = (ROWS % 4 == 0) ? 4 : 2; static constexpr int NV = ROWS / W; };

template <int ROWS>
__device__ __forceinline__ void fh_g_get(const float* __restrict__ src, float (&r)[ROWS]) {
  if constexpr (FhVec<ROWS>::W == 4) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int h = 0; h < ROWS / 4; ++h) {
      const float4 v = s4[threadIdx.x + FH_PASS_THREADS * h];
      r[4 * h] = v.x; r[4 * h + 1] = v.y; r[4 * h + 2] = v.z; r[4 * h + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int h = 0; h < ROWS / 2; ++h) {
      const float2 v = reinterpret_cast<const float2*>(src)[threadIdx.x + FH_PASS_THREADS * h];
      r[2 * h] = v.x; r[2 * h + 1] = v.y;
    }
  }
}
template <int ROWS>
__device__ __forceinline__ void fh_g_put(float* __restrict__ dst, const float (&r)[ROWS]) {
  if constexpr (FhVec<ROWS>::W == 4) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int h = 0; h < ROWS / 4; ++h) d4[threadIdx.x + FH_PASS_THREADS * h] = make_float4(r[4 * h], r[4 * h + 1], r[4 * h + 2], r[4 * h + 3]);
  } else {
#pragma unroll
    for (int h = 0; h < ROWS / 2; ++h)
      reinterpret_cast<float2*>(dst)[threadIdx.x + FH_PASS_THREADS * h] = make_float2(r[2 * h], r[2 * h + 1]);
  }
}
// element offset (inside the strip) of the thread's vector h
template <int ROWS>
__device__ __forceinline__ int fh_vec_off(int h) { return FhVec<ROWS>::W * (threadIdx.x + FH_PASS_THREADS * h); }

template <int ROWS>
__device__ __forceinline__ void fh_strip_get(const float* sm_rows, float (&r)[ROWS]) {
  constexpr int W = FhVec<ROWS>::W;
#pragma unroll
  for (int h = 0; h < FhVec<ROWS>::NV; ++h) {
    const int e = fh_vec_off<ROWS>(h);
    const float* p = sm_rows + (e >> 8) * FH_RS + (e & 255);
    if constexpr (W == 4) {
      const float4 v = *reinterpret_cast<const float4*>(p);
      r[4 * h] = v.x; r[4 * h + 1] = v.y; r[4 * h + 2] = v.z; r[4 * h + 3] = v.w;
    } else {
      const float2 v = *reinterpret_cast<const float2*>(p);
      r[2 * h] = v.x; r[2 * h + 1] = v.y;
    }
  }
}
template <int ROWS>
__device__ __forceinline__ void fh_strip_put(float* sm_rows, const float (&r)[ROWS]) {
  constexpr int W = FhVec<ROWS>::W;
#pragma unroll
  for (int h = 0; h < FhVec<ROWS>::NV; ++h) {
    const int e = fh_vec_off<ROWS>(h);
    float* p = sm_rows + (e >> 8) * FH_RS + (e & 255);
    if constexpr (W == 4) *reinterpret_cast<float4*>(p) = make_float4(r[4 * h], r[4 * h + 1], r[4 * h + 2], r[4 * h + 3]);
    else *reinterpret_cast<float2*>(p) = make_float2(r[2 * h], r[2 * h + 1]);
  }
}
template <int ROWS>
__device__ __forceinline__ void fh_strip_load(float* sm_rows, const float* __restrict__ src) {
  float r[ROWS];
  fh_g_get<ROWS>(src, r);
  fh_strip_put<ROWS>(sm_rows, r);
}
template <int ROWS>
__device__ __forceinline__ void fh_strip_store(const float* sm_rows, float* __restrict__ dst) {
  float r[ROWS];
  fh_strip_get<ROWS>(sm_rows, r);
  fh_g_put<ROWS>(dst, r);
}
// transposed store of the strip: dst_plane[k][row0 + r] = rows[r][k]  (thread k writes ROWS contiguous floats)
template <int ROWS>
__device__ __forceinline__ void fh_strip_store_t(const float* sm_rows, float* __restrict__ dst_plane, int row0) {
  const int k = threadIdx.x;
  float* d = dst_plane + (size_t)k * FH_N + row0;
  if constexpr (ROWS >= 4) {
#pragma unroll
    for (int h = 0; h < ROWS / 4; ++h)
      reinterpret_cast<float4*>(d)[h] = make_float4(sm_rows[(4 * h) * FH_RS + k], sm_rows[(4 * h + 1) * FH_RS + k],
                                                    sm_rows[(4 * h + 2) * FH_RS + k], sm_rows[(4 * h + 3) * FH_RS + k]);
  } else {
    *reinterpret_cast<float2*>(d) = make_float2(sm_rows[k], sm_rows[FH_RS + k]);
  }
}

// In-place circular CORRELATION along the rows of the strip with taps t[0..n) centred at c:
//     out[j] = sum_a t[a] in[(j + a - c) mod 256]
// (A^T of the blur uses the PSF taps as they are; A uses them reversed with centre n-1-c -- see fh_cg.cuh.)
// `taps` is a shared-memory array zero-padded to a multiple of 4.  All 256 threads; contains two __syncthreads.
// Fast path (8-row strips, c % 4 == 0): each thread produces 8 consecutive outputs from a register window that
// slides over the row 4 taps at a time -- one LDS.128 per 32 FMAs.
template <int ROWS>
__device__ __forceinline__ void fh_strip_conv(float* sm_rows, const float* __restrict__ taps, int n, int c) {
  float acc[ROWS];
#pragma unroll
  for (int j = 0; j < ROWS; ++j) acc[j] = 0.f;
  if (ROWS == 8 && (c & 3) == 0) {
    const int row = threadIdx.x >> 5, l = threadIdx.x & 31;
    const float* rr = sm_rows + row * FH_RS;
    const int base = 8 * l - c;                           // multiple of 4
    float4 f0 = *reinterpret_cast<const float4*>(rr + (base & (FH_N - 1)));
    float4 f1 = *reinterpret_cast<const float4*>(rr + ((base + 4) & (FH_N - 1)));
    for (int a0 = 0; a0 < n; a0 += 4) {
      const float4 f2 = *reinterpret_cast<const float4*>(rr + ((base + a0 + 8) & (FH_N - 1)));
      const float4 tk = *reinterpret_cast<const float4*>(taps + a0);
      const float win[12] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w};
      const float tt[4] = {tk.x, tk.y, tk.z, tk.w};
#pragma unroll
      for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int e = 0; e < ROWS; ++e) acc[e] = fmaf(tt[aa], win[aa + e], acc[e]);
      f0 = f1; f1 = f2;
    }
    __syncthreads();
    float* wr = sm_rows + row * FH_RS + 8 * l;
    *reinterpret_cast<float4*>(wr) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(wr + 4) = make_float4(acc[4 % ROWS], acc[5 % ROWS], acc[6 % ROWS], acc[7 % ROWS]);
    __syncthreads();
    return;
  }
  constexpr int TPR = FH_PASS_THREADS / ROWS;        // threads per row; each produces ROWS outputs
  const int row = threadIdx.x / TPR, l = threadIdx.x % TPR;
  const float* rr = sm_rows + row * FH_RS;
  for (int a = 0; a < n; ++a) {
    const float w = taps[a];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) acc[j] = fmaf(w, rr[(l + TPR * j + a - c) & (FH_N - 1)], acc[j]);
  }
  __syncthreads();
  float* wr = sm_rows + row * FH_RS;
#pragma unroll
  for (int j = 0; j < ROWS; ++j) wr[l + TPR * j] = acc[j];
  __syncthreads();
}
// Stage separable taps for fh_strip_conv: forward = 0 (A^T): t = k, centre c;  forward = 1 (A): t = reversed k,
// centre n-1-c.  Zero-padded to a multiple of 4.  Caller syncs.  Returns the centre to pass to fh_strip_conv.
__device__ __forceinline__ int fh_stage_taps(float* s_taps, const float* __restrict__ k, int n, int c, int forward) {
  const int np = (n + 3) & ~3;
  for (int i = threadIdx.x; i < np; i += blockDim.x) s_taps[i] = (i < n) ? (forward ? k[n - 1 - i] : k[i]) : 0.f;
  return forward ? n - 1 - c : c;
}

// Plain row pass: out = (TR ? transpose : identity)( DCT/IDCT along rows of in ).  grid.x strides over the
// nplanes*(256/ROWS) strips.
template <bool INV, bool TR, int ROWS>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_pass(const float* __restrict__ in, float* __restrict__ out, int nstrips, FhDctTables tab) {
  __shared__ __align__(16) FhStripSmem<ROWS> sm;
  constexpr int SPP = FH_N / ROWS;                   // strips per plane
  fh_strip_tables<INV>(sm, tab);
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const size_t base = (size_t)s * ROWS * FH_N;
    __syncthreads();
    fh_strip_load<ROWS>(sm.a, in + base);
    __syncthreads();
    fh_dct_strip<INV, ROWS>(sm.a, sm.b, sm);
    __syncthreads();
    if (TR) fh_strip_store_t<ROWS>(sm.b, out + (size_t)(s / SPP) * FH_N * FH_N, (s % SPP) * ROWS);
    else fh_strip_store<ROWS>(sm.b, out + base);
  }
}

// Plane transpose (import/export of basis-layout vectors, not on the hot path). grid = (8, 8, nplanes), block (32, 8).
__global__ void fh_k_transpose(const float* __restrict__ in, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const size_t plane = (size_t)blockIdx.z * FH_N * FH_N;
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) tile[j][threadIdx.x] = in[plane + (size_t)(y0 + j) * FH_N + x];
  __syncthreads();
  const int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) out[plane + (size_t)(oy0 + j) * FH_N + ox] = tile[threadIdx.x][j];
}
