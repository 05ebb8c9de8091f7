// Orthonormal 256-point DCT-II / DCT-III along rows, two rows packed into one 256-point complex FFT
// (Makhoul reordering), FFT done as a 16 x 16 four-step transform with 16 threads per FFT and a
// 16-point FFT held in registers.  A CTA transforms R = 8 rows and stores the result TRANSPOSED, so
// running the same kernel twice yields the 2-D transform in the original orientation.
//
// Replaces torch_dct.dct_2d / idct_2d(norm='ortho') (reference call sites
// conditioning_utils/online_update_bfgs.py:352-374; torch-dct 0.1.6 computes the same transform through
// an FFT of the even/odd reordered sequence).
#pragma once
#include "fh_common.cuh"

struct FhDctTables {
  const float2* tw256;   // exp(-2 pi i j / 256), j = 0..255
  const float2* post;    // f_k * exp(-i pi k / 512): forward post-twiddle incl. orthonormal scale
  const float2* cw;      // exp(+i pi k / 512): inverse pre-twiddle
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

#define FH_DCT_ROWS 8        // rows per CTA
#define FH_DCT_THREADS 64    // 4 FFT groups of 16 threads, 2 rows each
#define FH_OUTT_STRIDE 10    // floats per transposed row in smem (8 used): conflict-free float2 stores

// in:  planes [nplanes][256][256] (row-major).  out: same shape, each plane TRANSPOSED after the row transform.
// grid = (256 / FH_DCT_ROWS, nplanes), block = FH_DCT_THREADS.
template <bool INV>
__global__ void __launch_bounds__(FH_DCT_THREADS)
fh_k_dct_rows_t(const float* __restrict__ in, float* __restrict__ out, FhDctTables tab) {
  __shared__ __align__(16) float rows[FH_DCT_ROWS][FH_N];
  __shared__ __align__(16) float2 scratch[4][16 * 17];
  __shared__ __align__(16) float2 zb[4][FH_N];
  __shared__ __align__(16) float outT[FH_N * FH_OUTT_STRIDE];
  __shared__ __align__(16) float2 s_tw[FH_N];
  __shared__ __align__(16) float2 s_tab[FH_N];

  const int tid = threadIdx.x;
  const int g = tid >> 4, t = tid & 15;
  const int row0 = blockIdx.x * FH_DCT_ROWS;
  const size_t plane = (size_t)blockIdx.y * FH_N * FH_N;

  // tables + input rows (coalesced float4)
  for (int i = tid; i < FH_N; i += FH_DCT_THREADS) {
    float2 w = tab.tw256[i];
    s_tw[i] = INV ? make_float2(w.x, -w.y) : w;
    s_tab[i] = INV ? tab.cw[i] : tab.post[i];
  }
  {
    const float4* src = reinterpret_cast<const float4*>(in + plane + (size_t)row0 * FH_N);
    float4* dst = reinterpret_cast<float4*>(&rows[0][0]);
    for (int i = tid; i < FH_DCT_ROWS * FH_N / 4; i += FH_DCT_THREADS) dst[i] = src[i];
  }
  __syncthreads();

  const float* ra = rows[2 * g];
  const float* rb = rows[2 * g + 1];
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
      v[n2] = fh_cmul(s_tab[k], make_float2(A, B));
    }
  }
  // ---- four-step FFT: n = n1 + 16 n2 (n1 = t), k = 16 k1 + k2
  constexpr int SIGN = INV ? +1 : -1;
  fh_fft16<SIGN>(v);                                   // over n2 -> Y[n1=t][k2]
  float2* S = scratch[g];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    float2 y = fh_cmul(v[k2], s_tw[(t * k2) & (FH_N - 1)]);
    S[t * 17 + k2] = y;
  }
  __syncwarp();
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = S[n1 * 17 + t];   // this thread: k2 = t
  fh_fft16<SIGN>(v);                                   // over n1 -> X[16 k1 + t]

  if (!INV) {
    float2* Z = zb[g];
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
      const float2 p = s_tab[k];
      const float xa = p.x * va.x - p.y * va.y;
      const float xb = p.x * vb.x - p.y * vb.y;
      *reinterpret_cast<float2*>(&outT[k * FH_OUTT_STRIDE + 2 * g]) = make_float2(xa, xb);
    }
  } else {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const int n = 16 * k1 + t;
      const int pos = (n < 128) ? 2 * n : 511 - 2 * n;
      *reinterpret_cast<float2*>(&outT[pos * FH_OUTT_STRIDE + 2 * g]) = v[k1];
    }
  }
  __syncthreads();
  // transposed store: out[plane][j][row0 .. row0+8)
  for (int j = tid; j < FH_N; j += FH_DCT_THREADS) {
    const float2* s = reinterpret_cast<const float2*>(&outT[j * FH_OUTT_STRIDE]);
    float2 a = s[0], b = s[1], c = s[2], d = s[3];
    float4* dst = reinterpret_cast<float4*>(out + plane + (size_t)j * FH_N + row0);
    dst[0] = make_float4(a.x, a.y, b.x, b.y);
    dst[1] = make_float4(c.x, c.y, d.x, d.y);
  }
}
