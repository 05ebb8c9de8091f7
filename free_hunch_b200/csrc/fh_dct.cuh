// Orthonormal 256-point DCT-II / DCT-III along rows of 256x256 planes.
//
// Core: two real rows are packed into one 256-point complex FFT (Makhoul reordering); the FFT is a 16 x 16
// four-step transform done by 16 threads, each holding a 16-point FFT in registers.  A CTA of 256 threads
// owns a strip of 8 consecutive rows (2048 contiguous floats) staged in shared memory: all 256 threads move
// data (coalesced float4) and do the fused element-wise / convolution / dot-product work of the CG kernels in
// fh_cg.cuh, threads 0..63 (4 groups of 16) run the FFTs.
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
  const float2* tw256;   // exp(-2 pi i j / 256), j = 0..255
  const float2* post;    // f_k * exp(-i pi k / 512): forward post-twiddle incl. orthonormal scale
  const float2* cw;      // exp(+i pi k / 512): inverse pre-twiddle
};

#define FH_STRIP 8            // rows per CTA strip
#define FH_RS 264             // shared-memory row stride in floats (16-B aligned rows, conflict-free FFT stores)
#define FH_PASS_THREADS 256
#define FH_STRIP_ELEMS (FH_STRIP * FH_N)

struct FhStripSmem {
  float a[FH_STRIP * FH_RS];
  float b[FH_STRIP * FH_RS];
  float2 grp[FH_STRIP / 2][16 * 17];
  float2 tw[FH_N];
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

// 16-point DFT in registers: in v[n], out v[k] (natural order), W = exp(SIGN*2 pi i/16).
template <int SIGN>
__device__ __forceinline__ void fh_fft16(float2 (&v)[16]) {
  // step A: for each n1 (0..3): DFT4 over n2 of v[n1 + 4 n2]  -> y[n1][k2] kept at v[n1 + 4 k2]
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) fh_dft4<SIGN>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
  // twiddles W16^(n1*k2)
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
  // W16^j = (cos(2 pi j/16), SIGN*sin(2 pi j/16))
  const float2 w1 = make_float2(c1, SIGN * s1), w2 = make_float2(r2, SIGN * r2), w3 = make_float2(s1, SIGN * c1);
  const float2 w6 = make_float2(-r2, SIGN * r2), w9 = make_float2(-c1, -SIGN * s1);
  v[1 + 4] = fh_cmul(v[1 + 4], w1);               // n1=1,k2=1
  v[1 + 8] = fh_cmul(v[1 + 8], w2);               // n1=1,k2=2
  v[1 + 12] = fh_cmul(v[1 + 12], w3);             // n1=1,k2=3
  v[2 + 4] = fh_cmul(v[2 + 4], w2);               // n1=2,k2=1
  v[2 + 8] = fh_rot90<SIGN>(v[2 + 8]);            // n1=2,k2=2 : W16^4 = exp(SIGN i pi/2)
  v[2 + 12] = fh_cmul(v[2 + 12], w6);             // n1=2,k2=3
  v[3 + 4] = fh_cmul(v[3 + 4], w3);               // n1=3,k2=1
  v[3 + 8] = fh_cmul(v[3 + 8], w6);               // n1=3,k2=2
  v[3 + 12] = fh_cmul(v[3 + 12], w9);             // n1=3,k2=3
  // step B: for each k2: DFT4 over n1 of v[n1 + 4 k2] -> X[4 k1 + k2]
  float2 o[16];
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    float2 a0 = v[4 * k2], a1 = v[4 * k2 + 1], a2 = v[4 * k2 + 2], a3 = v[4 * k2 + 3];
    fh_dft4<SIGN>(a0, a1, a2, a3);
    o[k2] = a0; o[4 + k2] = a1; o[8 + k2] = a2; o[12 + k2] = a3;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// Load the twiddle tables of one direction into the strip's shared memory (all threads; caller syncs).
template <bool INV>
__device__ __forceinline__ void fh_strip_tables(FhStripSmem& sm, const FhDctTables& tab) {
  for (int i = threadIdx.x; i < FH_N; i += blockDim.x) {
    const float2 w = tab.tw256[i];
    sm.tw[i] = INV ? make_float2(w.x, -w.y) : w;
    sm.tab[i] = INV ? tab.cw[i] : tab.post[i];
  }
}

// Transform the 8 rows in `in` (shared, stride FH_RS) into `out` (shared, stride FH_RS; must not alias `in`).
// Call with threadIdx.x < 64 only, from whole warps (it synchronises with __syncwarp).
template <bool INV>
__device__ __forceinline__ void fh_dct8_core(const float* __restrict__ in, float* __restrict__ out, FhStripSmem& sm) {
  const int tid = threadIdx.x;
  const int g = tid >> 4, t = tid & 15;
  const float* ra = in + (2 * g) * FH_RS;
  const float* rb = ra + FH_RS;
  float* oa = out + (2 * g) * FH_RS;
  float* ob = oa + FH_RS;
  float2 v[16];
  if (!INV) {
    // z[n] = va[n] + i vb[n],  v[n] = a[2n] (n<128) | a[511-2n] (n>=128);  this thread: n = t + 16*n2
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
      const int n = t + 16 * n2;
      const int src = (n < 128) ? 2 * n : 511 - 2 * n;
      v[n2] = make_float2(ra[src], rb[src]);
    }
  } else {
    // Z[k] = cw[k] * ((Ca[k] + Cb[N-k]) + i (Cb[k] - Ca[N-k])),  C[k] = X[k]*g_k/N, C[N] = 0
    const float sc0 = 0.0625f, sc1 = 0.044194173824159216f;     // 1/sqrt(N), 1/sqrt(2N)
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
      const int k = t + 16 * n2;
      const float sk = (k == 0) ? sc0 : sc1;
      const int kn = (FH_N - k) & (FH_N - 1);
      const float sn = (k == 0) ? 0.f : sc1;
      const float A = ra[k] * sk + rb[kn] * sn;
      const float B = rb[k] * sk - ra[kn] * sn;
      v[n2] = fh_cmul(sm.tab[k], make_float2(A, B));
    }
  }
  // ---- four-step FFT: n = n1 + 16 n2 (n1 = t), k = 16 k1 + k2
  constexpr int SIGN = INV ? +1 : -1;
  fh_fft16<SIGN>(v);                                   // over n2 -> Y[n1=t][k2]
  float2* S = sm.grp[g];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) S[t * 17 + k2] = fh_cmul(v[k2], sm.tw[(t * k2) & (FH_N - 1)]);
  __syncwarp();
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = S[n1 * 17 + t];   // this thread: k2 = t
  fh_fft16<SIGN>(v);                                   // over n1 -> X[16 k1 + t]

  if (!INV) {
    __syncwarp();                                      // every read of S done before it is reused as Z
    float2* Z = S;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) Z[16 * k1 + t] = v[k1];
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const int k = 16 * k1 + t;
      const float2 zk = v[k1];
      const float2 zn = Z[(FH_N - k) & (FH_N - 1)];
      const float2 va = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      const float2 vb = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
      const float2 p = sm.tab[k];
      oa[k] = p.x * va.x - p.y * va.y;
      ob[k] = p.x * vb.x - p.y * vb.y;
    }
    __syncwarp();                                      // Z (= S) is free again for the next strip
  } else {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const int n = 16 * k1 + t;
      const int pos = (n < 128) ? 2 * n : 511 - 2 * n;
      oa[pos] = v[k1].x;
      ob[pos] = v[k1].y;
    }
    __syncwarp();
  }
}

// ---- strip movers (all 256 threads) -------------------------------------------------------------
// thread owns float4 #tid and #tid+256 of the strip's 2048 contiguous floats: element offset e = 4*(tid+256h)
__device__ __forceinline__ void fh_strip_load(float* sm_rows, const float* __restrict__ src) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = threadIdx.x + FH_PASS_THREADS * h;
    const float4 v = s4[f];
    *reinterpret_cast<float4*>(sm_rows + (f >> 6) * FH_RS + 4 * (f & 63)) = v;
  }
}
__device__ __forceinline__ void fh_strip_get(const float* sm_rows, float (&r)[8]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = threadIdx.x + FH_PASS_THREADS * h;
    const float4 v = *reinterpret_cast<const float4*>(sm_rows + (f >> 6) * FH_RS + 4 * (f & 63));
    r[4 * h] = v.x; r[4 * h + 1] = v.y; r[4 * h + 2] = v.z; r[4 * h + 3] = v.w;
  }
}
__device__ __forceinline__ void fh_strip_put(float* sm_rows, const float (&r)[8]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = threadIdx.x + FH_PASS_THREADS * h;
    *reinterpret_cast<float4*>(sm_rows + (f >> 6) * FH_RS + 4 * (f & 63)) = make_float4(r[4 * h], r[4 * h + 1], r[4 * h + 2], r[4 * h + 3]);
  }
}
__device__ __forceinline__ void fh_g_get(const float* __restrict__ src, float (&r)[8]) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 v = s4[threadIdx.x + FH_PASS_THREADS * h];
    r[4 * h] = v.x; r[4 * h + 1] = v.y; r[4 * h + 2] = v.z; r[4 * h + 3] = v.w;
  }
}
__device__ __forceinline__ void fh_g_put(float* __restrict__ dst, const float (&r)[8]) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int h = 0; h < 2; ++h) d4[threadIdx.x + FH_PASS_THREADS * h] = make_float4(r[4 * h], r[4 * h + 1], r[4 * h + 2], r[4 * h + 3]);
}
__device__ __forceinline__ void fh_strip_store(const float* sm_rows, float* __restrict__ dst) {
  float r[8];
  fh_strip_get(sm_rows, r);
  fh_g_put(dst, r);
}
// transposed store of the strip: dst_plane[k][row0 + r] = rows[r][k]  (thread k writes 32 contiguous bytes)
__device__ __forceinline__ void fh_strip_store_t(const float* sm_rows, float* __restrict__ dst_plane, int row0) {
  const int k = threadIdx.x;
  float4* d = reinterpret_cast<float4*>(dst_plane + (size_t)k * FH_N + row0);
  d[0] = make_float4(sm_rows[k], sm_rows[FH_RS + k], sm_rows[2 * FH_RS + k], sm_rows[3 * FH_RS + k]);
  d[1] = make_float4(sm_rows[4 * FH_RS + k], sm_rows[5 * FH_RS + k], sm_rows[6 * FH_RS + k], sm_rows[7 * FH_RS + k]);
}

// In-place circular convolution along the rows of the strip with taps k[0..n) centred at c:
//   transpose = 0:  out[j] = sum_a k[a] in[(j - (a - c)) mod 256]       transpose = 1:  in[(j + (a - c)) mod 256]
// (all 256 threads; contains two __syncthreads)
__device__ __forceinline__ void fh_strip_conv(float* sm_rows, const float* __restrict__ taps, int n, int c, int transpose) {
  const int row = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* rr = sm_rows + row * FH_RS;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int a = 0; a < n; ++a) {
    const float w = taps[a];
    const int off = transpose ? (a - c) : -(a - c);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, rr[(lane + 32 * j + off) & (FH_N - 1)], acc[j]);
  }
  __syncthreads();
  float* wr = sm_rows + row * FH_RS;
#pragma unroll
  for (int j = 0; j < 8; ++j) wr[lane + 32 * j] = acc[j];
  __syncthreads();
}

// Plain row pass: out = (TR ? transpose : identity)( DCT/IDCT along rows of in ).  grid.x strides over the
// nplanes*32 strips.
template <bool INV, bool TR>
__global__ void __launch_bounds__(FH_PASS_THREADS)
fh_k_pass(const float* __restrict__ in, float* __restrict__ out, int nstrips, FhDctTables tab) {
  __shared__ __align__(16) FhStripSmem sm;
  fh_strip_tables<INV>(sm, tab);
  for (int s = blockIdx.x; s < nstrips; s += gridDim.x) {
    const size_t base = (size_t)s * FH_STRIP_ELEMS;
    __syncthreads();
    fh_strip_load(sm.a, in + base);
    __syncthreads();
    if (threadIdx.x < 64) fh_dct8_core<INV>(sm.a, sm.b, sm);
    __syncthreads();
    if (TR) fh_strip_store_t(sm.b, out + (size_t)(s >> 5) * FH_N * FH_N, (s & 31) * FH_STRIP);
    else fh_strip_store(sm.b, out + base);
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
