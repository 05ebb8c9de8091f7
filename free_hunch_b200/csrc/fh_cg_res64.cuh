// Resident CG solve in FLOAT64 arithmetic: the parity path of the guidance solve.
//
// Why float64.  The reference runs cg.py:230-292 in float64 on a system whose condition number reaches ~1e6
// (Sigma = dct_variance spans 7 decades, sigma_y^2 = 0.01).  A CPU experiment with the oracle shows
// that ANY float32 rounding inside the loop -- even rounding only q = A p to float32 with everything else in float64 --
// delays the stop by 20-50 % (34 instead of 28 iterations at sigma = 80, 2670 instead of 1937 at rtol 1e-6) and moves the
// first-call output by O(1) at the reference's default rtol.  So the vectors x, r, p, q, the 2-D DCTs, the blur taps, the
// low-rank products and every reduction are float64 here; only the STORED operands delta and Z stay float32 (they define
// the operator; widening them is exact).  B200 issues 57.6 DFMA/clk/SM (measured, tools/micro/fp64_rates.cu), half the
// FP32 rate, so the price is < 2x on the arithmetic and nothing on the exchanges.
//
// Structure (one cooperative launch per image, 128 CTAs x 256 threads, CTA s owns rows [6s, 6s+6) of the 768 image rows
// in the pixel domain and of the basis layout in the DCT domain, Z slices resident in shared memory -- see
// fh_cg_resident.cuh for the layout).  Float64 makes two algebraic restructurings safe that float32 did not allow (the
// float32 kernel lost conjugacy with them), and they remove two of the five grid-wide exchanges of an iteration:
//   * linearity: the forward 2-D DCT is applied to the new RESIDUAL r' as soon as it exists, and the direction's
//     transform follows by  w_hat(p') = w_hat(r') + beta w_hat(p)  once beta is known.  The r.r partials therefore ride on
//     the grid barrier of the first plane exchange instead of needing an exchange of their own;
//   * Parseval: p.Ap = sigma_y^2 p.p + w_hat . (Sigma_hat w_hat) is available in the DCT domain, BEFORE the inverse
//     transform, so its partials ride on the barrier of the second plane exchange.
// One pass k of the loop (k = 0: initial residual r0 = b - A_op b with x0 = b, cg.py:230-235):
//   F1  v = r (k = 0: b); mask / row-blur^T; row DCT; transposed store -> T0          | barrier 1 (+ r.r partials)
//   F2  T0 -> column blur^T; column DCT = w_hat(r); stop test of iteration k-1; beta; p = r + beta p; w_hat(p) likewise
//       partial Z (delta w_hat)                                                        | dot exchange (two hops)
//   B1  coef = Gs sums; v = delta (w_hat + Z^T coef); p.Ap partial; column IDCT; column blur; -> T1 | barrier 2 (+ p.Ap)
//   B2  T1 -> row IDCT; row blur / mask = t; q = sy2 p + t; alpha; x += alpha p; r -= alpha q; r.r partial
#pragma once
#include "fh_cg_resident.cuh"

#define FH_R64_LD 320                                // padded FFT buffer length (FH_PAD), in double2
#define FH_R64_RS 320                                // row stride of a staged strip, in doubles (256 + 16 B of padding per 64 B)
// column -> position inside a staged row: 2 doubles of padding after every 8, so that the 16-byte window loads of the
// convolution (lanes 64 B apart) fall on distinct bank groups (unpadded: 4-way conflicts, ncu: 35 % excess wavefronts)
#define FH_R64_POS(c) ((c) + (((c) >> 3) << 1))
#define FH_R64_TW 252                                // stage twiddles: Ns = 4 (3 x 4), 16 (3 x 16), 64 (3 x 64)
// Strips never straddle a channel plane here (unlike the float32 kernel's 128 x 6 rows): every plane is cut into 42 strips
// of 6 rows and one of 4, i.e. 43 CTAs per plane, 129 for an RGB image.  The gathers of the non-separable operators
// (halo rows wrap around inside ONE plane) would otherwise need two segments in the two straddling CTAs, which the whole
// grid would wait for at the next barrier.
#define FH_R64_CPP ((FH_N + FH_RES_ROWS - 1) / FH_RES_ROWS)          // CTAs per plane (43)
#define FH_R64_ZT_MIN_ITERS 6                        // Zt mode is not worth its prologue below this many expected iterations

struct FhRes64Smem {
  double2 X[FH_RES_ROWS / 2][FH_R64_LD];             // FFT ping-pong; X doubles as the strip's input rows, Y as its output
  double2 Y[FH_RES_ROWS / 2][FH_R64_LD];             // rows; Y also hosts the scratch of the low-rank stages (FhRes64Scratch)
  // The twiddle tables (FhRes64Tabs, 8 KB) follow this struct only in the TABS = true instantiation: with them 30 rows
  // of Z fit beside the FFT buffers, without them 32 -- the rank a 30-step Heun run ends with.  TABS = false reads them
  // through the read-only path instead (they stay in the ~28 KB of L1 that the 227 KB carve-out leaves: every other
  // global access of the kernel bypasses L1), which costs 1.7 us per iteration (measured) and is used only above rank 30.
  double taps[4][FH_RES_TAPS];                       // row A^T, column A^T, column A, row A (centre padded to a multiple of 4)
  double s_red[4][8];                                // block-sum slots: ride 1, p.Ap partial, ride 2, r.r partial
  double s_keep[3][FH_RES_MAXM];                     // Zt mode (survives the transforms): totals of Zt.h(r'), Z delta w_hat(p), coef
  fh_mbar_t bar;
  unsigned long long pad_;
};
struct FhRes64Tabs {
  double2 tw[FH_R64_TW + 4];                         // [stage][r = 1..3][k < Ns]: exp(-2 pi i k r / (4 Ns))
  double2 e512[FH_N];                                // exp(-i pi k / 512): post- (forward) and conjugated pre-twiddle (inverse)
};
struct FhRes64Scratch {                              // lives in Y while no transform is in flight
  double s_w[FH_PASS_THREADS / 32][FH_MAXM];
  double s_sum[FH_MAXM];
  double s_c[FH_MAXM];
};
static_assert(sizeof(FhRes64Scratch) <= sizeof(double2) * (FH_RES_ROWS / 2) * FH_R64_LD, "scratch must fit the FFT buffer");
static_assert(FH_RES_ROWS * FH_R64_RS * sizeof(double) <= sizeof(double2) * (FH_RES_ROWS / 2) * FH_R64_LD, "rows must fit the FFT buffer");

struct FhR64Gen {
  int mode;                       // 0: none (mask / separable blur / plain), 3: dense 2-D PSF as column runs, 4: x4 super-resolution
  // mode 3: the PSF as one dense vertical run per column (fh_op: fh_k_conv_runs), weights in window order per direction
  const int4* runs; const double* w_fwd; const double* w_tr; int nruns, ntaps, ymin, ymax, xmin, xmax;   // runs / weights padded to 6 taps
  int cw;                         // output columns per staging chunk (256, 128, 64 or 32)
  // mode 4: dense ks x ks PSF (row-major), centre ks / 2 (a multiple of 4), 256 x 256 planes, 64 x 64 measurements
  const float* psf; int ks;
  int gbuf_bytes;                 // shared-memory staging area the mode needs (between the tables and the Z rows)
};

struct FhCgRes64 {
  const double* bvec;          // (d) right-hand side b (pixel layout); x0 = b
  double* x_out;               // (d) solution
  double* T[2];                // (d) each: half-transformed planes (forward / inverse), transposed exchange through L2
  double* ride[2];             // (grid) each: per-CTA partial sums that ride on barrier 1 (r.r; pass 0: b.b) / barrier 2 (p.Ap)
  uint4* ll_dots;              // [FH_MAXM][grid] flag-in-data slots: per-CTA partial of (Z (delta w_hat))_i
  uint4* ll_sum;               // [FH_MAXM] slots, FH_LL_SPREAD apart: the totals, published by row i's owner CTA
  unsigned* grid_counter;
  unsigned seq_base;
  double rtol, atol, sy2;
  int use_stop, image, timing;
  FhR64Gen gen;                // non-separable operator (mode 0: none)
  int apply_only;              // 1: no solve -- x_out = Sigma bvec (one pass of the loop without operator, fh_cov_dot_f64)
  int mr_cap;                  // rows of Z that fit in shared memory
  int zt_mode;                 // Zt mode (see the kernel): 0 off, 1 on, 2 decided from *niter_hint (iterations of the image's previous solve)
  int zt_wide;                 // the launch reserved 2 m row slots: Zt mode keeps its rows as float64 (no widening in the two low-rank loops)
  int* niter_hint;             // device word of this image: written at the end of every solve, read at the start of the next
  FhCgInfo* info_out;
  const double2* tw;           // device tables (FH_R64_TW, FH_N entries)
  const double2* e512;
  const int* only_if;          // != null: the launch does nothing unless only_if[image] != 0 (every CTA reads the same word)
};

// ---- float64 pieces of the strip transform -----------------------------------------------------------------------
__device__ __forceinline__ double2 fh_cmul64(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 fh_cadd64(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 fh_csub64(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
template <int SIGN>
__device__ __forceinline__ void fh_dft4_64(double2& a0, double2& a1, double2& a2, double2& a3) {
  const double2 s02 = fh_cadd64(a0, a2), d02 = fh_csub64(a0, a2), s13 = fh_cadd64(a1, a3), t = fh_csub64(a1, a3);
  const double2 d13 = SIGN < 0 ? make_double2(t.y, -t.x) : make_double2(-t.y, t.x);
  a0 = fh_cadd64(s02, s13); a2 = fh_csub64(s02, s13); a1 = fh_cadd64(d02, d13); a3 = fh_csub64(d02, d13);
}
__device__ __forceinline__ double2 fh_ldg_d2(const double2* p) {
#ifdef FH_EMU
  return *p;
#else
  double2 v;
  asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#endif
}
template <bool TABS> __device__ __forceinline__ double2 fh_tab(const double2* p) { return TABS ? *p : fh_ldg_d2(p); }
template <int SIGN, int NS, bool TABS>
__device__ __forceinline__ void fh_fft256_stage64(const double2* __restrict__ in, double2* __restrict__ out, int j, const double2* __restrict__ tw) {
  const int k = j & (NS - 1);
  double2 v0 = in[FH_PAD(j)], v1 = in[FH_PAD(j + 64)], v2 = in[FH_PAD(j + 128)], v3 = in[FH_PAD(j + 192)];
  constexpr int OFF = (NS == 4) ? 0 : (NS == 16) ? 12 : 60;
  double2 w1 = fh_tab<TABS>(tw + OFF + k), w2 = fh_tab<TABS>(tw + OFF + NS + k), w3 = fh_tab<TABS>(tw + OFF + 2 * NS + k);
  if (SIGN > 0) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
  v1 = fh_cmul64(v1, w1); v2 = fh_cmul64(v2, w2); v3 = fh_cmul64(v3, w3);
  fh_dft4_64<SIGN>(v0, v1, v2, v3);
  const int j0 = ((j - k) << 2) + k;
  out[FH_PAD(j0)] = v0; out[FH_PAD(j0 + NS)] = v1; out[FH_PAD(j0 + 2 * NS)] = v2; out[FH_PAD(j0 + 3 * NS)] = v3;
}
// Barrier of ONE 64-thread FFT group (named barrier 1 + f): the three transforms of a strip are independent, so their
// stage exchanges need not wait for each other nor for the two idle warps (ncu: 39 % of the stall samples were CTA-wide
// barriers).  The emulation has one barrier kind; there every thread calls it.
#ifdef FH_EMU
#define FH_GROUP_SYNC(act, f) __syncthreads()
#else
#define FH_GROUP_SYNC(act, f) do { if (act) asm volatile("bar.sync %0, 64;" ::"r"(1 + (f)) : "memory"); } while (0)
#endif
// Orthonormal DCT-II (INV = false) / DCT-III (INV = true) of the 6 rows in `in` (stride FH_R64_RS) into `out`; same
// algorithm, aliasing rules and barrier contract as fh_dct_strip_core (fh_dct.cuh): `in` may alias Xb, `out` may alias Yb.
template <bool INV, bool TABS>
__device__ __forceinline__ void fh_dct64_core(const double* in, double* out, double2 (*Xb)[FH_R64_LD], double2 (*Yb)[FH_R64_LD],
                                              const double2* __restrict__ tw, const double2* __restrict__ e512, int npair) {
  constexpr int SIGN = INV ? +1 : -1;
  const int f = threadIdx.x >> 6, j = threadIdx.x & 63;
  const bool act = f < npair;
  const int fc = act ? f : 0;
  const double* ra = in + (2 * fc) * FH_R64_RS;
  const double* rb = ra + FH_R64_RS;
  double* oa = out + (2 * fc) * FH_R64_RS;
  double* ob = oa + FH_R64_RS;
  double2* X = Xb[fc];
  double2* Y = Yb[fc];
  if (act) {
    double2 v[4];
    if (!INV) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = j + 64 * r;
        const int src = (n < 128) ? 2 * n : 511 - 2 * n;
        v[r] = make_double2(ra[FH_R64_POS(src)], rb[FH_R64_POS(src)]);
      }
    } else {
      const double sc0 = 0.0625, sc1 = 0.044194173824159216;       // 1/sqrt(N), 1/sqrt(2N)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = j + 64 * r;
        const double sk = (k == 0) ? sc0 : sc1;
        const int kn = (FH_N - k) & (FH_N - 1);
        const double sn = (k == 0) ? 0.0 : sc1;
        const double A = ra[FH_R64_POS(k)] * sk + rb[FH_R64_POS(kn)] * sn;
        const double B = rb[FH_R64_POS(k)] * sk - ra[FH_R64_POS(kn)] * sn;
        const double2 e = fh_tab<TABS>(e512 + k);
        v[r] = fh_cmul64(make_double2(e.x, -e.y), make_double2(A, B));
      }
    }
    fh_dft4_64<SIGN>(v[0], v[1], v[2], v[3]);
    double2* y = Y + FH_PAD(4 * j);
    y[0] = v[0]; y[1] = v[1]; y[2] = v[2]; y[3] = v[3];
  }
  // `in` may alias Xb: EVERY group must have read its rows before ANY group writes Xb (rows 2f, 2f+1 coincide with Xb[f],
  // so in fact a group only overwrites its own rows -- the group barrier suffices)
  FH_GROUP_SYNC(act, fc);
  if (act) fh_fft256_stage64<SIGN, 4, TABS>(Y, X, j, tw);
  FH_GROUP_SYNC(act, fc);
  if (act) fh_fft256_stage64<SIGN, 16, TABS>(X, Y, j, tw);
  FH_GROUP_SYNC(act, fc);
  if (act) fh_fft256_stage64<SIGN, 64, TABS>(Y, X, j, tw);
  FH_GROUP_SYNC(act, fc);
  if (act) {
    if (!INV) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = j + 64 * r;
        const double2 zk = X[FH_PAD(k)];
        const double2 zn = X[FH_PAD((FH_N - k) & (FH_N - 1))];
        const double hf = (k == 0) ? 0.03125 : 0.044194173824159216;            // 0.5 * sqrt(1/N), 0.5 * sqrt(2/N)
        const double2 va = make_double2(hf * (zk.x + zn.x), hf * (zk.y - zn.y));
        const double2 vb = make_double2(hf * (zk.y + zn.y), -hf * (zk.x - zn.x));
        const double2 e = fh_tab<TABS>(e512 + k);
        oa[FH_R64_POS(k)] = e.x * va.x - e.y * va.y;
        ob[FH_R64_POS(k)] = e.x * vb.x - e.y * vb.y;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = j + 64 * r;
        const int pos = (n < 128) ? 2 * n : 511 - 2 * n;
        const double2 z = X[FH_PAD(n)];
        oa[FH_R64_POS(pos)] = z.x;
        ob[FH_R64_POS(pos)] = z.y;
      }
    }
  }
}

// the thread's 6 elements of a strip: double2 #t, #t+256, #t+512 (flat element e = 2 (t + 256 h) -> row e >> 8, column e & 255)
__device__ __forceinline__ void fh_r64_put(double* sm_rows, const double (&r)[FH_RES_ROWS]) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    const int e = 2 * (threadIdx.x + FH_PASS_THREADS * h);
    *reinterpret_cast<double2*>(sm_rows + (e >> 8) * FH_R64_RS + FH_R64_POS(e & 255)) = make_double2(r[2 * h], r[2 * h + 1]);
  }
}
__device__ __forceinline__ void fh_r64_get(const double* sm_rows, double (&r)[FH_RES_ROWS], int npair) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    const int e = 2 * (threadIdx.x + FH_PASS_THREADS * h);
    double2 v = make_double2(0.0, 0.0);
    if (h < npair) v = *reinterpret_cast<const double2*>(sm_rows + (e >> 8) * FH_R64_RS + FH_R64_POS(e & 255));
    r[2 * h] = v.x; r[2 * h + 1] = v.y;
  }
}
__device__ __forceinline__ double2 fh_ld_cg_d2(const double2* p) {
#ifdef FH_EMU
  return *p;
#else
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void fh_r64_load_plane(const double* src, double (&r)[FH_RES_ROWS], int npair) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    double2 v = make_double2(0.0, 0.0);
    if (h < npair) v = fh_ld_cg_d2(reinterpret_cast<const double2*>(src) + threadIdx.x + FH_PASS_THREADS * h);
    r[2 * h] = v.x; r[2 * h + 1] = v.y;
  }
}
__device__ __forceinline__ void fh_r64_load(const double* __restrict__ src, double (&r)[FH_RES_ROWS], int npair) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    double2 v = make_double2(0.0, 0.0);
    if (h < npair) v = reinterpret_cast<const double2*>(src)[threadIdx.x + FH_PASS_THREADS * h];
    r[2 * h] = v.x; r[2 * h + 1] = v.y;
  }
}
__device__ __forceinline__ void fh_r64_store(double* dst, const double (&r)[FH_RES_ROWS], int npair) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h)
    if (h < npair) reinterpret_cast<double2*>(dst)[threadIdx.x + FH_PASS_THREADS * h] = make_double2(r[2 * h], r[2 * h + 1]);
}
// float32 strip operands (delta, Z slices): float2 #t, #t+256, #t+512, zero beyond the strip's rows
__device__ __forceinline__ void fh_r64_getf(const float* src, float (&r)[FH_RES_ROWS], int npair) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    float2 v = make_float2(0.f, 0.f);
    if (h < npair) v = reinterpret_cast<const float2*>(src)[threadIdx.x + FH_PASS_THREADS * h];
    r[2 * h] = v.x; r[2 * h + 1] = v.y;
  }
}
// transposed store of the strip starting at global row grow0 into the exchange plane: element (global row g, column k)
// -> row (plane(g), k), column g & 255
__device__ __forceinline__ void fh_r64_store_t(const double* sm_rows, double* __restrict__ T, int grow0, int npair) {
  const int k = threadIdx.x;
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    if (h < npair) {
      const int g = grow0 + 2 * h;
      double* d = T + (size_t)(g >> 8) * FH_N * FH_N + (size_t)k * FH_N + (g & (FH_N - 1));
      *reinterpret_cast<double2*>(d) = make_double2(sm_rows[(2 * h) * FH_R64_RS + FH_R64_POS(k)], sm_rows[(2 * h + 1) * FH_R64_RS + FH_R64_POS(k)]);
    }
  }
}

// In-place circular correlation along the 6 rows, float64 (register window, 8 outputs per thread, warps 0..5 active).
__device__ __forceinline__ void fh_r64_conv(double* sm_rows, const double* __restrict__ taps, int n4, int c4, int nrows) {
  const int row = threadIdx.x >> 5, l = threadIdx.x & 31;
  const bool act = row < nrows;
  double acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0;
  if (act) {
    const double* rr = sm_rows + row * FH_R64_RS;
    const int base = 8 * l - c4;
    double win[12];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = (base + 2 * u) & (FH_N - 1);
      const double2 v = *reinterpret_cast<const double2*>(rr + FH_R64_POS(c));
      win[2 * u] = v.x; win[2 * u + 1] = v.y;
    }
    for (int a0 = 0; a0 < n4; a0 += 4) {
      const int c0 = (base + a0 + 8) & (FH_N - 1), c1 = (base + a0 + 10) & (FH_N - 1);
      const double2 f0 = *reinterpret_cast<const double2*>(rr + FH_R64_POS(c0));
      const double2 f1 = *reinterpret_cast<const double2*>(rr + FH_R64_POS(c1));
      win[8] = f0.x; win[9] = f0.y; win[10] = f1.x; win[11] = f1.y;
      const double2 t0 = *reinterpret_cast<const double2*>(taps + a0);
      const double2 t1 = *reinterpret_cast<const double2*>(taps + a0 + 2);
      const double tt[4] = {t0.x, t0.y, t1.x, t1.y};
#pragma unroll
      for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fma(tt[aa], win[aa + e], acc[e]);
#pragma unroll
      for (int u = 0; u < 8; ++u) win[u] = win[u + 4];
    }
  }
  __syncthreads();
  if (act) {
    double* wr = sm_rows + row * FH_R64_RS + FH_R64_POS(8 * l);
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<double2*>(wr + 2 * u) = make_double2(acc[2 * u], acc[2 * u + 1]);
  }
  __syncthreads();
}
// like fh_res_stage_taps, widened to float64
// Taps below 1e-18 of the largest one are left out at both ends: in a float64 sum they are under half an ulp of the
// result (the reference's 61-tap Gaussian of std 3 has 3 such taps per side: 64 -> 56 padded taps per convolution, and
// the four convolutions of an iteration run at the issue rate of the FP64 pipe).
__device__ __forceinline__ int fh_r64_stage_taps(double* s_taps, const float* __restrict__ k, int n, int c, int forward) {
  float mx = 0.f;
  for (int i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(k[i]));
  const float thr = mx * 1e-18f;
  int lo = 0, hi = n - 1;
  while (lo < c && fabsf(k[lo]) <= thr) ++lo;
  while (hi > c && fabsf(k[hi]) <= thr) --hi;
  k += lo; n = hi - lo + 1; c -= lo;
  const int cen = forward ? n - 1 - c : c;
  const int sh = (4 - (cen & 3)) & 3;
  const int n4 = (n + sh + 3) & ~3;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int a = i - sh;
    s_taps[i] = (a >= 0 && a < n) ? (double)(forward ? k[n - 1 - a] : k[a]) : 0.0;
  }
  return (n4 << 8) | (cen + sh);
}
// the thread's 6 elements of a Z row slice that is not resident: from L2, past L1 (L1 is left to the twiddle tables)
__device__ __forceinline__ void fh_r64_get_z_l2(const float* src, float (&r)[FH_RES_ROWS], int npair) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    float2 v = make_float2(0.f, 0.f);
    if (h < npair) v = fh_ld_cg(reinterpret_cast<const float2*>(src) + threadIdx.x + FH_PASS_THREADS * h);
    r[2 * h] = v.x; r[2 * h + 1] = v.y;
  }
}
// lane l ends with the warp total of v[l] (31 shuffles of 64 bits)
__device__ __forceinline__ double fh_warp_transpose_sum32_d(double (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < o; ++k) {
      const double send = hi ? v[k] : v[k + o];
      const double keep = hi ? v[k + o] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}
// 8 values per lane -> lane l ends with the warp total of v[4 b4 + 2 b3 + b2] (b_k = bit k of l): three halving exchanges,
// two butterfly steps.  The row loops of the low-rank stages run over chunks of 8 rows with this instead of one fully
// unrolled pass over 32 (fh_warp_transpose_sum32_d): the unrolled bodies had grown the kernel to 250 KB of SASS, more
// than the instruction caches hold, and every iteration walks most of it.
__device__ __forceinline__ double fh_warp_transpose_sum8_d(double (&v)[8]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16, w = 4; w >= 1; o >>= 1, w >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < w; ++k) {
      const double send = hi ? v[k] : v[k + w];
      const double keep = hi ? v[k + w] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}
#define FH_R64_T8_ROW(lane) (((lane) >> 2) & 7)        // which of the 8 values lane `lane` holds after fh_warp_transpose_sum8_d
// Block sums without barriers of their own: every warp posts its total into an 8-entry slot BEFORE a __syncthreads the
// caller needs anyway, and whoever needs the block total adds the 8 entries in a fixed order AFTER it (bit-identical in
// every thread; a slot is rewritten one pass later, many barriers after its last read).
__device__ __forceinline__ void fh_r64_post(double v, double* slot8) {
  v = fh_warp_sum(v);
  if ((threadIdx.x & 31) == 0) slot8[threadIdx.x >> 5] = v;
}
__device__ __forceinline__ double fh_r64_sum8(const double* slot8) {
  double t = slot8[0];
#pragma unroll
  for (int w = 1; w < FH_PASS_THREADS / 32; ++w) t += slot8[w];
  return t;
}
// per-thread share of the partials that rode on a grid barrier (read past L1); block total via fh_r64_post / fh_r64_sum8
__device__ __forceinline__ double fh_r64_ride_part(const double* ride, int nblk) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nblk; i += FH_PASS_THREADS) v += fh_ldcg(ride + i);
  return v;
}
// float32 -> float64 on the integer pipe (exact for normal numbers; zero becomes 2^-127, subnormals are garbage of that
// size: both vanish against the other operand).  F2F.F64.F32 issues at 16/clk/SM on B200 (measured) and so does this
// sequence on the ALU pipe: alternating them doubles the widening rate of the low-rank stages.
__device__ __forceinline__ double fh_widen_alu(float f) {
  const unsigned u = fh_f2u(f);
  const unsigned hi = (((unsigned)((int)u >> 3)) & 0x8fffffffu) + 0x38000000u;
  return fh_u2d(u << 29, hi);
}
#define FH_WIDEN(z, e) (((e) & 1) ? fh_widen_alu(z) : (double)(z))

// ---- non-separable operators (2-D PSF, x4 super-resolution): gathers through L2 --------------------------------------
// A CTA owns rows [grow0, grow0 + nrows) of one plane.  A^T r and A t need rows of OTHER CTAs, so the operand is first
// published to global memory (it stays in L2), the grid synchronises, and every CTA gathers its halo back into shared
// memory.  Two extra grid barriers per iteration compared with the mask / separable blur, which fold into the DCT passes.
#define FH_TAPS64_CHUNK 512       // taps staged per round in fh_k_conv_taps64 (6 KB of static shared memory)
#define FH_R64_SRP 28             // padded PSF row length of the x4 SR gathers (25 taps + 3 zeros: b0 + 4 t never leaves the row)

// 16-byte asynchronous copies global -> shared past L1 (the published planes are rewritten every iteration by other CTAs:
// an L1-allocating load could return last iteration's line), grouped so that the NEXT chunk's halo is in flight while
// the current one is being convolved.
#ifdef FH_EMU
static inline void fh_cp16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
static inline void fh_cp_commit() {}
template <int N> static inline void fh_cp_wait() {}
#else
__device__ __forceinline__ void fh_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(fh_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void fh_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void fh_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif
// tile[r][0 .. 2 npc) <- plane[(gi0 + r) & 255][(gj0 + ..) & 255], gj0 even; row stride tws (even).  Issues copies only.
__device__ __forceinline__ void fh_r64_stage_tile(double* tile, int tws, const double* plane, int gi0, int gj0, int th, int npc) {
  // a warp per tile row, lanes along it: no index division per copy (the flat loop spent ~40 instructions per 16 bytes)
  const int lane = threadIdx.x & 31;
  for (int rr = threadIdx.x >> 5; rr < th; rr += FH_PASS_THREADS / 32) {
    const double* prow = plane + (size_t)((gi0 + rr) & (FH_N - 1)) * FH_N;
    double* trow = tile + rr * tws;
    for (int pc = lane; pc < npc; pc += 32) fh_cp16(trow + 2 * pc, prow + ((gj0 + 2 * pc) & (FH_N - 1)));
  }
}

// mode 3: out rows (shared, staged-row layout) = A src (transpose = 0) or A^T src (transpose = 1) for this CTA's rows;
// src is a published (C, 256, 256) float64 image.  The halo is staged in chunks of G.cw output columns (256, 128, 64 or
// 32: the host gives the gather as much shared memory as the rank leaves, because a chunk is one L2 round trip -- ~2 us
// from a plane other SMs have just written -- and the bytes in flight are what hides it).  Thread = (output column of
// the chunk, run group): a group owns a subset of the PSF's column runs, every thread slides a 6-row window down its
// image column (one shared load per tap for 6 FMAs); the groups' partial sums meet in shared memory in a fixed order.
// Contains __syncthreads; all 256 threads.
__device__ __forceinline__ void fh_r64_gather_conv(const double* src, double* out_rows, unsigned char* gbuf, const FhR64Gen& G,
                                                   int pl, int row0, int transpose) {
  const int tid = threadIdx.x;
  const int cw = G.cw, ngrp = FH_PASS_THREADS / cw;
  const int oy0 = transpose ? G.ymin : -G.ymax, oy1 = transpose ? G.ymax : -G.ymin;
  const int ox0 = transpose ? G.xmin : -G.xmax, ox1 = transpose ? G.xmax : -G.xmin;
  const int th = FH_RES_ROWS + (oy1 - oy0) + FH_RES_ROWS;                  // + 6: the zero-weight taps that pad a run to a multiple of 6 read on
  const int tws = (cw + (ox1 - ox0) + 2) & ~1, npc = tws / 2;              // one spare column: the even-aligned origin may sit one left
  double* tile = reinterpret_cast<double*>(gbuf);                          // [th][tws]
  double* red = tile + (size_t)th * tws;                                   // [ngrp - 1][6][cw] partial sums of run groups 1..
  double* wts = red + (size_t)(ngrp - 1) * FH_RES_ROWS * cw;               // weights of this direction, widened, every run padded to a multiple of 6
  int4* srun = reinterpret_cast<int4*>(wts + ((G.ntaps + 1) & ~1));   // [nruns] (tile row of the window start, tile column offset, L6, woff6)
  const double* plane = src + (size_t)pl * FH_N * FH_N;
  const int sh = (ox0 & 1);                                                // chunk origins j0 are even: parity of the first column
  __syncthreads();                                                         // gbuf may still be in use by the previous gather
  fh_r64_stage_tile(tile, tws, plane, row0 + oy0, ox0 - sh, th, npc);
  fh_cp_commit();
  const double* wsrc = transpose ? G.w_tr : G.w_fwd;
  for (int t = tid; t < G.ntaps; t += FH_PASS_THREADS) wts[t] = wsrc[t];
  for (int t = tid; t < G.nruns; t += FH_PASS_THREADS) {
    const int4 r = G.runs[t];                                              // (dx, dy0, L, offset of the padded weights)
    const int L6 = (r.z + FH_RES_ROWS - 1) / FH_RES_ROWS * FH_RES_ROWS;
    srun[t] = make_int4(transpose ? (r.y - oy0) : (-r.y - (r.z - 1) - oy0), (transpose ? r.x : -r.x) - ox0 + sh, L6, r.w);
  }
  const int jc = tid % cw, grp = tid / cw;
  for (int j0 = 0; j0 < FH_N; j0 += cw) {
    fh_cp_wait<0>();
    __syncthreads();
    double acc[FH_RES_ROWS];
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) acc[e] = 0.0;
    for (int t = grp; t < G.nruns; t += ngrp) {
      const int4 r = srun[t];
      const double* q = tile + r.x * tws + r.y + jc;
      const double* w = wts + r.w;
      double win[FH_RES_ROWS];
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) win[e] = q[e * tws];
      q += FH_RES_ROWS * tws;
      // Six taps per round with a compile-time rotation of the window: before tap k slot j holds the input row congruent
      // to j (mod 6) among rows k .. k+5, so output e reads slot (e + k) % 6 and the dead slot k % 6 takes row k + 6.
      for (int k0 = 0; k0 < r.z; k0 += FH_RES_ROWS) {
        // (measured, 12.5 us per gather at rank 0: skipping the 125 zero-weight taps of the motion PSF's 342 padded ones or
        // loading the weights as 16-byte pairs changes nothing; two runs in flight per thread with separate accumulators
        // -- twice the loads in flight per round -- is SLOWER: 13.3 us, 64.6 -> 72.7 us per iteration at rank 56)
#pragma unroll
        for (int kk = 0; kk < FH_RES_ROWS; ++kk) {
          const double wv = w[k0 + kk];
#pragma unroll
          for (int e = 0; e < FH_RES_ROWS; ++e) acc[e] = fma(wv, win[(e + kk) % FH_RES_ROWS], acc[e]);
          win[kk] = q[(k0 + kk) * tws];
        }
      }
    }
    if (grp > 0) {
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) red[((size_t)(grp - 1) * FH_RES_ROWS + e) * cw + jc] = acc[e];
    }
    __syncthreads();                                                       // tile fully read, partials visible
    if (j0 + cw < FH_N) {                                                  // next chunk's halo flies while group 0 finishes this one
      fh_r64_stage_tile(tile, tws, plane, row0 + oy0, j0 + cw + ox0 - sh, th, npc);
      fh_cp_commit();
    }
    if (grp == 0) {
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) {
        double v = acc[e];
        for (int g = 0; g < ngrp - 1; ++g) v += red[((size_t)g * FH_RES_ROWS + e) * cw + jc];
        out_rows[e * FH_R64_RS + FH_R64_POS(j0 + jc)] = v;
      }
    }
  }
  __syncthreads();
}

// mode 4: PSF widened into shared memory, rows padded to FH_R64_SRP.  Caller syncs.
__device__ __forceinline__ void fh_r64_sr_stage_psf(double* spsf, const FhR64Gen& G) {
  for (int i = threadIdx.x; i < FH_SR_KS * FH_R64_SRP; i += FH_PASS_THREADS) {
    const int a = i / FH_R64_SRP, b = i - a * FH_R64_SRP;
    spsf[i] = (a < G.ks && b < G.ks) ? (double)G.psf[a * G.ks + b] : 0.0;
  }
}
// The full rounds of the A^T gather for a strip whose first row is 0 (PAR = 0) or 2 (PAR = 1) modulo 4 (strips start at
// multiples of 6; the PSF centre is a multiple of 4).  In round u row e of the strip multiplies PSF row A0[e] + 4 u with
// measurement row YR[e] + u (relative to the first staged one): rows of the strip that share a PSF row or a measurement row
// share the LOADS -- 7 x (3 + 4) or 7 x (2 + 4) per round instead of 7 x 12; the gather runs at the pace of the shared-memory
// pipe.  The tables are compile-time, so the sharing is plain common-subexpression elimination.
template <int PAR>
__device__ __forceinline__ void fh_r64_sr_at_core(const double* spsf, const double* ys, const int (&qt)[7], int b0, int ks,
                                                  double (&acc)[FH_RES_ROWS]) {
  constexpr int A0[2][FH_RES_ROWS] = {{0, 3, 2, 1, 0, 3}, {2, 1, 0, 3, 2, 1}};
  constexpr int YR[2][FH_RES_ROWS] = {{0, 1, 1, 1, 1, 2}, {1, 1, 1, 2, 2, 2}};
  const int nfull = ks >> 2;                               // rounds in which every row's tap row A0 + 4 u exists (A0 <= 3)
  for (int u = 0; u < nfull; ++u) {
    const double* yu = ys + (size_t)u * 64;
    const double* ku = spsf + (size_t)(4 * u) * FH_R64_SRP + b0;
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      double yv[3], kv[4];
#pragma unroll
      for (int g = PAR; g < 3; ++g) yv[g] = yu[g * 64 + qt[t]];
      if (PAR) yv[0] = 0.0;
#pragma unroll
      for (int h = 0; h < 4; ++h) kv[h] = ku[h * FH_R64_SRP + 4 * t];        // taps beyond ks are zero
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) acc[e] = fma(kv[A0[PAR][e]], yv[YR[PAR][e]], acc[e]);
    }
  }
}
// mode 4, A^T: out rows = k (*)^T (zero insertion of the published 64 x 64 measurement plane `pl` of ysrc)   (utils_sisr.py:44-52)
//   x[i][j] = sum over taps with (i + a - c) % 4 == 0 and (j + b - c) % 4 == 0 of k[a][b] y[((i + a - c) & 255) / 4][((j + b - c) & 255) / 4]
__device__ __forceinline__ void fh_r64_sr_gather_at(const double* ysrc, double* out_rows, unsigned char* gbuf, const FhR64Gen& G,
                                                    int pl, int row0, int nrows) {
  const int tid = threadIdx.x;
  const int c = G.ks / 2;
  double* spsf = reinterpret_cast<double*>(gbuf);
  double* ys = spsf + FH_SR_KS * FH_R64_SRP;                               // [10][64] measurement rows this strip touches
  const int Ilo = (row0 - c) >> 2;                                         // floor
  const int nI = ((row0 + FH_RES_ROWS - 1 + (G.ks - 1 - c)) >> 2) - Ilo + 1;    // <= 10
  const double* yp = ysrc + (size_t)pl * (FH_N / 4) * (FH_N / 4);
  __syncthreads();
  {
    // <= 640 values: all three loads of a thread in flight together (as a loop each load waited for the store of the one
    // before it: three L2 round trips instead of one)
    double v[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int idx = tid + FH_PASS_THREADS * u;
      v[u] = idx < nI * 64 ? fh_ldcg(yp + (size_t)((Ilo + (idx >> 6)) & 63) * 64 + (idx & 63)) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int idx = tid + FH_PASS_THREADS * u;
      if (idx < nI * 64) ys[idx] = v[u];
    }
  }
  __syncthreads();
  const int j = tid;
  const int b0 = (4 - ((j - c) & 3)) & 3;
  int qt[7];
#pragma unroll
  for (int t = 0; t < 7; ++t) qt[t] = ((j + b0 + 4 * t - c) >> 2) & 63;
  // All 6 rows of the strip at once: 6 independent accumulator chains (the first version walked the rows one after another
  // with ONE accumulator -- 49 dependent DFMAs per output, 294 per thread, every one behind its own two shared-memory loads:
  // 20 us of a 45 us iteration).  Row e takes the tap rows a = a0[e] + 4 u; per output the order of the sum is unchanged
  // (a ascending, t ascending), so the results are bit-identical.  Rows beyond nrows (the last strip of a plane) are computed
  // on whatever the staging area holds and never stored.
  double acc[FH_RES_ROWS];
#pragma unroll
  for (int e = 0; e < FH_RES_ROWS; ++e) acc[e] = 0.0;
  if ((row0 >> 1) & 1) fh_r64_sr_at_core<1>(spsf, ys, qt, b0, G.ks, acc);
  else fh_r64_sr_at_core<0>(spsf, ys, qt, b0, G.ks, acc);
  {
    const int nfull = G.ks >> 2;
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) {
      const int a0e = (4 - ((row0 + e - c) & 3)) & 3;
      for (int a = a0e + 4 * nfull; a < G.ks; a += 4) {
        const double* yr = ys + (size_t)(((row0 + e + a - c) >> 2) - Ilo) * 64;
        const double* kr = spsf + a * FH_R64_SRP + b0;
#pragma unroll
        for (int t = 0; t < 7; ++t) acc[e] = fma(kr[4 * t], yr[qt[t]], acc[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < FH_RES_ROWS; ++e)
    if (e < nrows) out_rows[e * FH_R64_RS + FH_R64_POS(j)] = acc[e];
  __syncthreads();
}
// mode 4, A: the CTA's own measurement elements of plane pl from the published image tsrc   (utils_sisr.py:55-61)
//   y[I][J] = sum_{a,b} k[a][b] t[(4 I - a + c) & 255][(4 J - b + c) & 255]
// Ownership: the 64 x 64 plane in 256 GROUPS of 16 elements, group g = (column block g >> 6, row g & 63) -- column-major, so
// that the 5 or 6 consecutive groups [g0, g0 + ng) of a CTA are vertical neighbours whose 25-row image windows overlap in 21
// rows: ONE window of (4 (n - 1) + 25) x 88 values per run of n groups inside a column block (at most two runs), staged by
// cp.async, instead of a 25 x 88 window per group (the first version owned whole measurement rows: 4 or 8 windows per CTA,
// each exposing its L2 round trip, 13.6 MB through L2 per iteration -- 10.8 of the 29.5 us of an iteration on the CTAs
// with two rows).  Thread t < 16 ng owns element (group g0 + t / 16, column 16 (g >> 6) + t % 16); 16 threads share the
// taps of an element.  Returns the thread's own element.
#define FH_R64_SR_WROWS (4 * 6 - 8 + 2 * FH_SR_KS)                          // rows of both windows together: 4 (n - 1) + 25 each, n1 + n2 <= 6
__device__ __forceinline__ int fh_r64_sr_yindex(int g, int el) { return 64 * (g & 63) + 16 * (g >> 6) + el; }
__device__ __forceinline__ double fh_r64_sr_gather_a(const double* tsrc, unsigned char* gbuf, const FhR64Gen& G, int pl, int g0, int ng) {
  const int tid = threadIdx.x;
  const int c = G.ks / 2;
  double* spsf = reinterpret_cast<double*>(gbuf);
  double* qy = spsf + FH_SR_KS * FH_R64_SRP + 10 * 64;                     // [128] results
  constexpr int TW = 64 + 2 * 12, NPC = TW / 2;
  double* tile0 = qy + 128;                                                // [FH_R64_SR_WROWS][88]: the windows of the two runs, back to back
  const double* plane = tsrc + (size_t)pl * FH_N * FH_N;
  const int el = tid >> 4, part = tid & 15;
  // runs: [g0, gm) and [gm, g0 + ng), gm = the column-block boundary inside the range (or its end)
  const int gend = g0 + ng;
  const int gm = min(gend, ((g0 >> 6) + 1) << 6);
  __syncthreads();
  fh_r64_stage_tile(tile0, TW, plane, 4 * (g0 & 63) - c, 64 * (g0 >> 6) - c, 4 * (gm - g0 - 1) + FH_SR_KS, NPC);
  fh_cp_commit();
  const int rows0 = 4 * (gm - g0 - 1) + FH_SR_KS;
  if (gm < gend) fh_r64_stage_tile(tile0 + rows0 * TW, TW, plane, 4 * (gm & 63) - c, 64 * (gm >> 6) - c, 4 * (gend - gm - 1) + FH_SR_KS, NPC);
  fh_cp_commit();
  const bool two = part + 16 < G.ks;
  const bool one = part < G.ks;
  for (int run = 0; run < 2; ++run) {
    const int ga = run ? gm : g0, gb = run ? gend : gm;
    if (ga >= gb) break;
    if (run == 0) fh_cp_wait<1>(); else fh_cp_wait<0>();
    __syncthreads();
    const double* tile = tile0 + (size_t)run * rows0 * TW;
    // All groups of the run at once: lane `part` of an element's 16 takes columns b = part and part + 16 of every PSF row and
    // feeds the row's two weights to the (up to 6) vertically adjacent elements it serves -- 2 + 2 n loads for 2 n FMAs.  Group
    // by group it was 2 loads per FMA (0.9 us per group); earlier still one accumulator and q / ks per tap.  Measured inside
    // (clock64, CTA 0, 5 groups): staging + barrier 0.8 us, this loop 4.5 us, reduction 0.6 us -- the loop runs at the pace of
    // the shared-memory pipe: 8 warps x 12 loads x 2 wavefronts (two elements per warp span 160 bytes) per PSF row.
    const int n = gb - ga;
    double s0[6], s1[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { s0[i] = 0.0; s1[i] = 0.0; }
    const double* kp = spsf + part;
    const double* tp0 = tile + (2 * c) * TW + 4 * el + 2 * c - part;
    for (int a = 0; a < G.ks; ++a) {
      const double k0 = one ? kp[a * FH_R64_SRP] : 0.0;
      const double k1 = two ? kp[a * FH_R64_SRP + 16] : 0.0;
      const double* tr = tp0 - a * TW;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (i < n) {
          if (one) s0[i] = fma(k0, tr[4 * i * TW], s0[i]);
          if (two) s1[i] = fma(k1, tr[4 * i * TW - 16], s1[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i < n) {
        double acc = s0[i] + s1[i];
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (part == 0) qy[16 * (ga - g0 + i) + el] = acc;
      }
    }
  }
  __syncthreads();
  return tid < 16 * ng ? qy[tid] : 0.0;
}

// grid = 43 CTAs per plane (cooperative), block = 256, dynamic shared memory = FhRes64Smem [+ tables] [+ gather staging]
// + min(m, mr_cap) * 6 KB.  BIG = false: every row of Z resident.  BIG = true: rows >= mr_cap are re-read from L2.
// TABS = true: twiddle tables in shared memory (costs two rows of Z); false: read through L1.
// First hop of the dot exchange, one warp: row i of the low-rank products is owned by CTA i * (grid / orows); its owner
// sums the grid's partials in a fixed order and publishes the total.
__device__ __forceinline__ void fh_r64_dots_hop1(const uint4* ll_dots, uint4* ll_sum, int s, int nblk, int m, int orows, unsigned seq) {
  const int lane = threadIdx.x & 31;
  const int own_step = nblk / orows > 0 ? nblk / orows : 1;
  if (s % own_step == 0 && s / own_step < m) {
    const int i = s / own_step;
    const uint4* col = ll_dots + (size_t)i * nblk;
    double t = 0.0;
    for (int k0 = lane; k0 < nblk; k0 += 128) {
      if (k0 + 96 < nblk) {
        uint4 v[4];
        fh_ll_wait<4>(col + k0, 32, seq, v);
#pragma unroll
        for (int u = 0; u < 4; ++u) t += fh_u2d(v[u].x, v[u].z);
      } else {
        for (int kk = k0; kk < nblk; kk += 32) {
          uint4 v[1];
          fh_ll_wait<1>(col + kk, 0, seq, v);
          t += fh_u2d(v[0].x, v[0].z);
        }
      }
    }
    t = fh_warp_sum(t);
    if (lane == 0) fh_ll_put_double(ll_sum + (size_t)i * FH_LL_SPREAD, t, seq);
  }
}

template <bool BIG, bool TABS>
__global__ void __launch_bounds__(FH_PASS_THREADS, 1)
fh_k_cg_res64(FhCgFused P, FhCgRes64 R) {
  FH_DYN_SMEM(unsigned char, smem_raw);
  FhRes64Smem& sm = *reinterpret_cast<FhRes64Smem*>(smem_raw);
  FhRes64Scratch& sc = *reinterpret_cast<FhRes64Scratch*>(sm.Y);
  FhRes64Tabs* tabs = reinterpret_cast<FhRes64Tabs*>(smem_raw + sizeof(FhRes64Smem));
  unsigned char* gbuf = smem_raw + sizeof(FhRes64Smem) + (TABS ? sizeof(FhRes64Tabs) : 0);
  float* Zs = reinterpret_cast<float*>(gbuf + R.gen.gbuf_bytes);
  const double2* tw = TABS ? tabs->tw : R.tw;
  const double2* e512 = TABS ? tabs->e512 : R.e512;
  double* rows_a = reinterpret_cast<double*>(sm.X);
  double* rows_b = reinterpret_cast<double*>(sm.Y);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x;
  const unsigned nblk = gridDim.x;
  const int m = P.m;
  const int mr = BIG ? min(m, R.mr_cap) : m;
  // geometry: CTA cc of plane pl owns rows [6 cc, 6 cc + nrows) of that plane, in the pixel AND in the basis layout
  const int pl = s / FH_R64_CPP, cc = s - pl * FH_R64_CPP;
  const int prow0 = cc * FH_RES_ROWS;                                       // first row inside the plane
  const int nrows = min(FH_RES_ROWS, FH_N - prow0), npair = nrows >> 1;
  const int grow0 = pl * FH_N + prow0;
  const long d = (long)P.C * FH_N * FH_N;
  const size_t img = (size_t)R.image * d;
  const size_t soff = (size_t)grow0 * FH_N;
  const int fuse = R.gen.mode ? R.gen.mode : P.fuse;
  // measurement-space ownership: the strip itself, except x4 super-resolution (64 x 64 planes: 1 or 2 rows per CTA)
  // (the 4096 elements of a 64 x 64 plane in 256 column-major groups of 16, 5 or 6 groups per CTA: fh_r64_sr_gather_a)
  const int yg0 = (cc * 256) / FH_R64_CPP, yng = ((cc + 1) * 256) / FH_R64_CPP - yg0;
  const bool yown = tid < 16 * yng;
  const size_t yidx = (size_t)pl * (FH_N / 4) * (FH_N / 4) + (size_t)fh_r64_sr_yindex(yg0 + (yown ? tid >> 4 : 0), tid & 15);   // fuse == 4 only

  if (R.only_if && R.only_if[R.image] == 0) return;   // uniform over the grid: no CTA reaches a grid barrier
  // ---- prologue ----
  if (tid == 0) { fh_mbar_init(&sm.bar, 1); fh_mbar_fence_init(); }
  __syncthreads();
  const float* zsrc = P.Z + (size_t)R.image * FH_MAXM * d + soff;
  if (mr > 0 && warp == 0) {
    if (lane == 0) fh_mbar_expect_tx(&sm.bar, (unsigned)mr * (unsigned)nrows * FH_N * 4u);
    __syncwarp();
    for (int i = lane; i < mr; i += 32) fh_tma_load_1d(Zs + (size_t)i * FH_RES_SLICE, zsrc + (size_t)i * d, (unsigned)nrows * FH_N * 4u, &sm.bar);
  }
  if (TABS) {
    for (int i = tid; i < FH_R64_TW; i += FH_PASS_THREADS) tabs->tw[i] = R.tw[i];
    for (int i = tid; i < FH_N; i += FH_PASS_THREADS) tabs->e512[i] = R.e512[i];
  }
  int tp[4] = {0, 0, 0, 0};
  if (fuse == 2) {
    tp[0] = fh_r64_stage_taps(sm.taps[0], P.kr, P.nr, P.cr, 0);
    tp[1] = fh_r64_stage_taps(sm.taps[1], P.kc, P.nc, P.cc, 0);
    tp[2] = fh_r64_stage_taps(sm.taps[2], P.kc, P.nc, P.cc, 1);
    tp[3] = fh_r64_stage_taps(sm.taps[3], P.kr, P.nr, P.cr, 1);
  }
  if (fuse == 4) fh_r64_sr_stage_psf(reinterpret_cast<double*>(gbuf), R.gen);
  float dl[FH_RES_ROWS], mk[FH_RES_ROWS];
  fh_r64_getf(P.delta + img + soff, dl, npair);
#pragma unroll
  for (int e = 0; e < FH_RES_ROWS; ++e) mk[e] = 1.f;
  if (fuse == 1) fh_r64_getf(P.mask + (size_t)(P.mask_batch > 1 ? R.image : 0) * FH_N * FH_N + (size_t)prow0 * FH_N, mk, npair);
  constexpr int CP = BIG ? 4 : 8, NG = BIG ? 16 : 4, CROWS = FH_PASS_THREADS / CP;
  const int ci = tid / CP, cpart = tid % CP;
  double gs[NG];
#pragma unroll
  for (int u = 0; u < NG; ++u) gs[u] = 0.0;
  if (ci < m) {
    const double* g = P.Gs + ((size_t)R.image * FH_MAXM + ci) * FH_MAXM;
#pragma unroll
    for (int u = 0; u < NG; ++u) gs[u] = (cpart + CP * u < m) ? g[cpart + CP * u] : 0.0;
  }
  unsigned bar_target = 0;
  // CG vectors of the measurement space.  fuse 1..3: the strip (6 per thread).  fuse 4: element 0 of threads < 16 yng.
  double x[FH_RES_ROWS], r[FH_RES_ROWS], p[FH_RES_ROWS], wp[FH_RES_ROWS];
  if (fuse == 4) {
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) x[e] = 0.0;
    if (yown) x[0] = R.bvec[yidx];
  } else {
    fh_r64_load(R.bvec + soff, x, npair);
  }
  {
    double bb = 0.0;                                     // what rides on the first barrier: pass 0 |b|^2, later |r|^2
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) { r[e] = x[e]; p[e] = 0.0; wp[e] = 0.0; bb = fma(x[e], x[e], bb); }
    fh_r64_post(bb, sm.s_red[3]);
  }
  double rr_old = 0.0, stop2 = -1.0, resid2 = 0.0;
  int niter = 0, optimal = 0;
  bool z_ready = (mr == 0);
  __syncthreads();
  // ---- Zt mode: the resident rows become Zt_i = column-IDCT(delta o z_i), the low-rank factor in the HALF-transformed
  // domain the strip holds between the two transposes.  By Parseval Z (delta w_hat) = Zt . h (h = the strip before the
  // column DCT) and column-IDCT(delta Z^T c) = Zt^T c, so the partial dots are posted BEFORE the forward column transform
  // and the correction is added AFTER the inverse one: the two-hop exchange of the dots (2.7 - 4.6 us) runs while the
  // strip does both transforms, polled by warp 7, which the 192-thread transforms leave idle.  By linearity the dots are
  // taken of the new RESIDUAL and follow the direction as t(p') = t(r') + beta t(p).  Costs m strip transforms per solve:
  // used when the image's previous solve ran long enough (every CTA reads the same word: a uniform decision).
  bool zt = false;
  if (!BIG && m > 0 && !R.apply_only) {
    if (R.zt_mode == 1) zt = true;
    else if (R.zt_mode == 2 && R.niter_hint) zt = *reinterpret_cast<volatile const int*>(R.niter_hint) >= max(FH_R64_ZT_MIN_ITERS, (5 * m) >> 3);
  }
  double beta = 0.0, lr_pAp = 0.0;
  const bool wide = !BIG && zt && R.zt_wide;
  double* Zd = reinterpret_cast<double*>(Zs);
  if (!BIG && zt) {
    fh_mbar_wait(&sm.bar, 0);
    z_ready = true;
    if (tid < FH_RES_MAXM) sm.s_keep[1][tid] = 0.0;
    // wide: row i moves from its float32 slot [6 i, 6 i + 6) KB to the float64 slot [12 i, 12 i + 12) KB, last row first
    // (the float64 slot of row i starts at or behind the end of the float32 rows below it)
    for (int ii = 0; ii < m; ++ii) {
      const int i = wide ? m - 1 - ii : ii;
      float z[FH_RES_ROWS];
      double v[FH_RES_ROWS];
      fh_r64_getf(Zs + (size_t)i * FH_RES_SLICE, z, npair);
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) v[e] = (double)z[e] * (double)dl[e];
      fh_r64_put(rows_a, v);
      __syncthreads();
      fh_dct64_core<true, TABS>(rows_a, rows_b, sm.X, sm.Y, tw, e512, npair);
      __syncthreads();
      fh_r64_get(rows_b, v, npair);
#pragma unroll
      for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
        if (wide) reinterpret_cast<double2*>(Zd + (size_t)i * FH_RES_SLICE)[tid + FH_PASS_THREADS * h] = make_double2(v[2 * h], v[2 * h + 1]);   // zeros beyond the strip's rows
        else if (h < npair) reinterpret_cast<float2*>(Zs + (size_t)i * FH_RES_SLICE)[tid + FH_PASS_THREADS * h] = make_float2((float)v[2 * h], (float)v[2 * h + 1]);
      }
    }
    __syncthreads();
  }

#define FH_R64TRACE(slot) do { if (k == 3) FH_TRACE(P, 6, slot); } while (0)
  for (int k = 0;; ++k) {
    const unsigned seq = R.seq_base + (unsigned)k + 1u;
    // ------------------------------------------------------------------ F1: rows forward of A^T r (pass 0: of A^T b)
    FH_R64TRACE(0);
    if (fuse >= 3) {
      // publish r (measurement space) and meet: the r.r partial rides on this barrier instead of the next one
      if (fuse == 4) { if (yown) R.T[1][yidx] = r[0]; }
      else fh_r64_store(R.T[1] + soff, r, npair);
      __syncthreads();
      if (tid == 0) R.ride[0][s] = fh_r64_sum8(sm.s_red[3]);
      fh_grid_sync(R.grid_counter, bar_target, nblk);
      FH_R64TRACE(11);
      if (fuse == 4) fh_r64_sr_gather_at(R.T[1], rows_a, gbuf, R.gen, pl, prow0, nrows);
      else fh_r64_gather_conv(R.T[1], rows_a, gbuf, R.gen, pl, prow0, 1);
      FH_R64TRACE(12);
    } else {
      double w[FH_RES_ROWS];
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) w[e] = r[e] * (double)mk[e];
      fh_r64_put(rows_a, w);
      __syncthreads();
      if (tid == 0) R.ride[0][s] = fh_r64_sum8(sm.s_red[3]);
      if (fuse == 2) fh_r64_conv(rows_a, sm.taps[0], tp[0] >> 8, tp[0] & 255, nrows);
    }
    fh_dct64_core<false, TABS>(rows_a, rows_b, sm.X, sm.Y, tw, e512, npair);
    __syncthreads();
    fh_r64_store_t(rows_b, R.T[0], grow0, npair);
    FH_R64TRACE(1);
    fh_grid_sync(R.grid_counter, bar_target, nblk);
    // ------------------------------------------------------------------ F2: columns forward; scalars; new direction
    FH_R64TRACE(2);
    double wr[FH_RES_ROWS];
    {
      double t[FH_RES_ROWS];
      fh_r64_load_plane(R.T[0] + soff, t, npair);
      fh_r64_post(fh_r64_ride_part(R.ride[0], (int)nblk), sm.s_red[0]);
      fh_r64_put(rows_a, t);
      __syncthreads();
      const double S1 = fh_r64_sum8(sm.s_red[0]);
      if (fuse == 2) fh_r64_conv(rows_a, sm.taps[1], tp[1] >> 8, tp[1] & 255, nrows);
      if (!BIG && zt) {
        // Zt . h(r'): partials posted now, summed by the owners' warp 7 while warps 0..5 run the column transform
        double hh[FH_RES_ROWS];
        fh_r64_get(rows_a, hh, npair);
        if (wide) {
          for (int c0 = 0; c0 < m; c0 += 8) {
            double acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              acc[u] = 0.0;
              if (c0 + u < m) {
#pragma unroll
                for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
                  const double2 zz = reinterpret_cast<const double2*>(Zd + (size_t)(c0 + u) * FH_RES_SLICE)[tid + FH_PASS_THREADS * h];
                  acc[u] = fma(zz.y, hh[2 * h + 1], fma(zz.x, hh[2 * h], acc[u]));
                }
              }
            }
            const double tot = fh_warp_transpose_sum8_d(acc);
            if ((lane & 3) == 0 && c0 + FH_R64_T8_ROW(lane) < m) sc.s_w[warp][c0 + FH_R64_T8_ROW(lane)] = tot;
          }
        } else {
          for (int c0 = 0; c0 < m; c0 += 8) {
            double acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              acc[u] = 0.0;
              if (c0 + u < m) {
                float z[FH_RES_ROWS];
                fh_r64_getf(Zs + (size_t)(c0 + u) * FH_RES_SLICE, z, npair);
#pragma unroll
                for (int e = 0; e < FH_RES_ROWS; ++e) acc[u] = fma(FH_WIDEN(z[e], e), hh[e], acc[u]);
              }
            }
            const double tot = fh_warp_transpose_sum8_d(acc);
            if ((lane & 3) == 0 && c0 + FH_R64_T8_ROW(lane) < m) sc.s_w[warp][c0 + FH_R64_T8_ROW(lane)] = tot;
          }
        }
        __syncthreads();
        if (tid < m) {
          double t2 = 0.0;
#pragma unroll
          for (int wv = 0; wv < FH_PASS_THREADS / 32; ++wv) t2 += sc.s_w[wv][tid];
          fh_ll_put_double(R.ll_dots + (size_t)tid * nblk + s, t2, seq);
        }
        __syncthreads();                                 // the scratch in Y is read: the transform may overwrite it
        if (warp == FH_PASS_THREADS / 32 - 1) fh_r64_dots_hop1(R.ll_dots, R.ll_sum, s, (int)nblk, m, FH_RES_MAXM, seq);
      }
      fh_dct64_core<false, TABS>(rows_a, rows_b, sm.X, sm.Y, tw, e512, npair);
      __syncthreads();
      fh_r64_get(rows_b, wr, npair);
      __syncthreads();                                   // Y is free for the scratch of the low-rank stages
      if (k == 0) {
        const double stop = fmax(R.rtol * sqrt(S1), R.atol);                    // cg.py:237
        stop2 = R.use_stop ? stop * stop : -1.0;
        if (s == 0 && tid == 0) { P.sc[R.image].bnorm2 = S1; P.sc[R.image].stop2 = stop2; }
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) { p[e] = x[e]; wp[e] = wr[e]; }
      } else {
        resid2 = S1;
        beta = 0.0;
        if (k >= 2) {                                                          // end of CG iteration k - 1 (cg.py:258-263)
          niter = k - 1;
          const bool converged = (stop2 >= 0.0) && (S1 <= stop2);
          if (converged) optimal = 1;
          if (converged || k - 1 >= P.last_allowed) break;
          beta = S1 / rr_old;
        }
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) { p[e] = fma(beta, p[e], r[e]); wp[e] = fma(beta, wp[e], wr[e]); }
        rr_old = S1;
      }
    }
    FH_R64TRACE(3);
    // ------------------------------------------------------------------ low-rank part: coef = Gs Z (delta w_hat)
    double vv[FH_RES_ROWS];
#pragma unroll
    for (int e = 0; e < FH_RES_ROWS; ++e) vv[e] = wp[e];
    if (m > 0 && !(!BIG && zt)) {
      if (!z_ready) { fh_mbar_wait(&sm.bar, 0); z_ready = true; }
      double wd[FH_RES_ROWS];
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) wd[e] = wp[e] * (double)dl[e];
      for (int c0 = 0; c0 < mr; c0 += 8) {
        double acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc[u] = 0.0;
          if (c0 + u < mr) {
            float z[FH_RES_ROWS];
            fh_r64_getf(Zs + (size_t)(c0 + u) * FH_RES_SLICE, z, npair);
#pragma unroll
            for (int e = 0; e < FH_RES_ROWS; ++e) acc[u] = fma(FH_WIDEN(z[e], e), wd[e], acc[u]);
          }
        }
        const double tot = fh_warp_transpose_sum8_d(acc);
        if ((lane & 3) == 0 && c0 + FH_R64_T8_ROW(lane) < mr) sc.s_w[warp][c0 + FH_R64_T8_ROW(lane)] = tot;
      }
      if (BIG) {
        for (int c0 = mr; c0 < m; c0 += 8) {             // rows that did not fit: straight from L2, 8 rows (24 loads) in flight
          float z[8][FH_RES_ROWS];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (c0 + u < m) fh_r64_get_z_l2(zsrc + (size_t)(c0 + u) * d, z[u], npair);
          }
          double acc[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            acc[u] = 0.0;
            if (c0 + u < m) {
#pragma unroll
              for (int e = 0; e < FH_RES_ROWS; ++e) acc[u] = fma(FH_WIDEN(z[u][e], e), wd[e], acc[u]);
            }
          }
          const double tot = fh_warp_transpose_sum8_d(acc);
          if ((lane & 3) == 0 && c0 + FH_R64_T8_ROW(lane) < m) sc.s_w[warp][c0 + FH_R64_T8_ROW(lane)] = tot;
        }
      }
      __syncthreads();
      if (tid < m) {
        double t = 0.0;
#pragma unroll
        for (int wv = 0; wv < FH_PASS_THREADS / 32; ++wv) t += sc.s_w[wv][tid];
        fh_ll_put_double(R.ll_dots + (size_t)tid * nblk + s, t, seq);
      }
      FH_R64TRACE(4);
      {
        constexpr int OROWS = BIG ? FH_MAXM : FH_RES_MAXM;
        if (warp == 0) {
          fh_r64_dots_hop1(R.ll_dots, R.ll_sum, s, (int)nblk, m, OROWS, seq);
#pragma unroll
          for (int g = 0; g < (BIG ? 2 : 1); ++g) {
            const int row = lane + 32 * g;
            double tot = 0.0;
            if (row < m) {
              uint4 v[1];
              fh_ll_wait<1>(R.ll_sum + (size_t)row * FH_LL_SPREAD, 0, seq, v);
              tot = fh_u2d(v[0].x, v[0].z);
            }
            sc.s_sum[row] = tot;
          }
        }
        FH_R64TRACE(5);
        __syncthreads();
        double a = 0.0;
#pragma unroll
        for (int u = 0; u < NG; ++u) a += gs[u] * sc.s_sum[cpart + CP * u];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (!BIG) a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (cpart == 0 && ci < CROWS) sc.s_c[ci] = a;
        __syncthreads();
      }
      FH_R64TRACE(6);
#pragma unroll 4
      for (int i = 0; i < mr; ++i) {
        const double c = sc.s_c[i];
        float z[FH_RES_ROWS];
        fh_r64_getf(Zs + (size_t)i * FH_RES_SLICE, z, npair);
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) vv[e] = fma(c, FH_WIDEN(z[e], e), vv[e]);
      }
      if (BIG) {
        for (int c0 = mr; c0 < m; c0 += 4) {
          float z[4][FH_RES_ROWS];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c0 + u < m) fh_r64_get_z_l2(zsrc + (size_t)(c0 + u) * d, z[u], npair);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c0 + u < m) {
              const double c = sc.s_c[c0 + u];
#pragma unroll
              for (int e = 0; e < FH_RES_ROWS; ++e) vv[e] = fma(c, FH_WIDEN(z[u][e], e), vv[e]);
            }
          }
        }
      }
    }
    FH_R64TRACE(7);
    // ------------------------------------------------------------------ B1: Sigma_hat w_hat; p.Ap partial; columns inverse
    {
      double loc = 0.0, pp = 0.0;
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) {
        vv[e] *= (double)dl[e];
        loc = fma(wp[e], vv[e], loc);
        pp = fma(p[e], p[e], pp);
      }
      fh_r64_post(fma(R.sy2, pp, loc), sm.s_red[1]);      // Parseval: p.Ap = sy2 p.p + w_hat.(Sigma_hat w_hat)
      fh_r64_put(rows_a, vv);
      __syncthreads();                                   // also: the scratch in Y is dead, the transform may use Y again
      if (tid == 0) R.ride[1][s] = fh_r64_sum8(sm.s_red[1]);
      if (!BIG && zt && warp == FH_PASS_THREADS / 32 - 1) {
        // second hop, overlapped with the inverse column transform: the totals of Zt . h(r')
        double tot = 0.0;
        if (lane < m) {
          uint4 v[1];
          fh_ll_wait<1>(R.ll_sum + (size_t)lane * FH_LL_SPREAD, 0, seq, v);
          tot = fh_u2d(v[0].x, v[0].z);
        }
        sm.s_keep[0][lane] = tot;
      }
      fh_dct64_core<true, TABS>(rows_a, rows_b, sm.X, sm.Y, tw, e512, npair);
      __syncthreads();
      if (!BIG && zt) {
        if (tid < FH_RES_MAXM) sm.s_keep[1][tid] = fma(beta, sm.s_keep[1][tid], sm.s_keep[0][tid]);      // t(p') = t(r') + beta t(p)
        __syncthreads();
        double a = 0.0;
#pragma unroll
        for (int u = 0; u < NG; ++u) a += gs[u] * sm.s_keep[1][(cpart + CP * u) & (FH_RES_MAXM - 1)];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (cpart == 0 && ci < FH_RES_MAXM) sm.s_keep[2][ci] = a;                                       // coef = Gs t(p')
        __syncthreads();
        lr_pAp = fh_warp_sum(sm.s_keep[1][lane] * sm.s_keep[2][lane]);                                  // t . Gs t, the same in every warp and CTA
        double cv[FH_RES_ROWS];
        fh_r64_get(rows_b, cv, npair);
        if (wide) {
#pragma unroll 4
          for (int i = 0; i < m; ++i) {
            const double c = sm.s_keep[2][i];
#pragma unroll
            for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
              const double2 zz = reinterpret_cast<const double2*>(Zd + (size_t)i * FH_RES_SLICE)[tid + FH_PASS_THREADS * h];
              cv[2 * h] = fma(c, zz.x, cv[2 * h]);
              cv[2 * h + 1] = fma(c, zz.y, cv[2 * h + 1]);
            }
          }
        } else {
#pragma unroll 4
          for (int i = 0; i < m; ++i) {
            const double c = sm.s_keep[2][i];
            float z[FH_RES_ROWS];
            fh_r64_getf(Zs + (size_t)i * FH_RES_SLICE, z, npair);
#pragma unroll
            for (int e = 0; e < FH_RES_ROWS; ++e) cv[e] = fma(c, FH_WIDEN(z[e], e), cv[e]);
          }
        }
        fh_r64_put(rows_b, cv);                            // every thread rewrites exactly the elements it read
        __syncthreads();
      }
      if (fuse == 2) fh_r64_conv(rows_b, sm.taps[2], tp[2] >> 8, tp[2] & 255, nrows);
      fh_r64_store_t(rows_b, R.T[1], grow0, npair);
    }
    FH_R64TRACE(8);
    fh_grid_sync(R.grid_counter, bar_target, nblk);
    // ------------------------------------------------------------------ B2: rows inverse; q; step
    FH_R64TRACE(9);
    {
      double t[FH_RES_ROWS];
      fh_r64_load_plane(R.T[1] + soff, t, npair);
      fh_r64_post(fh_r64_ride_part(R.ride[1], (int)nblk), sm.s_red[2]);
      fh_r64_put(rows_a, t);
      __syncthreads();
      const double pAp = fh_r64_sum8(sm.s_red[2]) + lr_pAp;
      fh_dct64_core<true, TABS>(rows_a, rows_b, sm.X, sm.Y, tw, e512, npair);
      __syncthreads();
      if (fuse == 2) fh_r64_conv(rows_b, sm.taps[3], tp[3] >> 8, tp[3] & 255, nrows);
      fh_r64_get(rows_b, t, npair);
      if (R.apply_only) {                                // Sigma v for fh_cov_dot_f64: t is the result
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) x[e] = t[e];
        break;
      }
      if (fuse >= 3) {
        // publish t = Sigma A^T p (image space), meet, gather A t for the own measurement elements
        fh_r64_store(R.T[0] + soff, t, npair);
        FH_R64TRACE(13);
        fh_grid_sync(R.grid_counter, bar_target, nblk);
        FH_R64TRACE(14);
        if (fuse == 4) {
          const double q0 = fh_r64_sr_gather_a(R.T[0], gbuf, R.gen, pl, yg0, yng);
#pragma unroll
          for (int e = 0; e < FH_RES_ROWS; ++e) t[e] = 0.0;
          t[0] = q0;
        } else {
          fh_r64_gather_conv(R.T[0], rows_a, gbuf, R.gen, pl, prow0, 0);
          fh_r64_get(rows_a, t, npair);
        }
        FH_R64TRACE(15);
      }
      double rr = 0.0;
      if (k == 0) {
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) {        // r0 = b - A_op b     (cg.py:233-235, x0 = b)
          const double q = fma(R.sy2, p[e], t[e] * (double)mk[e]);
          r[e] = x[e] - q;
          rr = fma(r[e], r[e], rr);
        }
      } else {
        if (pAp <= 1e-16 && !R.timing) { niter = k; break; }                    // cg.py:250
        const double alpha = rr_old / pAp;
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) {
          const double q = fma(R.sy2, p[e], t[e] * (double)mk[e]);
          x[e] = fma(alpha, p[e], x[e]);
          r[e] = fma(-alpha, q, r[e]);
          rr = fma(r[e], r[e], rr);
        }
      }
      fh_r64_post(rr, sm.s_red[3]);
    }
    FH_R64TRACE(10);
  }
  if (!z_ready) fh_mbar_wait(&sm.bar, 0);
  if (fuse == 4 && !R.apply_only) { if (yown) R.x_out[yidx] = x[0]; }
  else fh_r64_store(R.x_out + soff, x, npair);
  if (s == 0 && tid == 0) {
    FhCgScalars* w = P.sc + R.image;
    w->resid2 = resid2;
    w->rr[0] = rr_old; w->rr[1] = rr_old;
    w->niter = niter;
    w->optimal = optimal;
    fh_set_done(w, niter + 1);
    if (R.info_out) { R.info_out->residual_norm = sqrt(resid2); R.info_out->niter = niter; R.info_out->optimal = optimal; }
    if (!R.apply_only && R.niter_hint) *R.niter_hint = niter;
  }
}

// ---- float64 helpers around the solve (once per solve, not per iteration) --------------------------------------------
// Circular convolution with a sparse tap list in float64 (the operator of fh_k_conv_taps):
//   transpose = 0:  out[i][j] = sum_t w_t in[(i - dy_t) mod H][(j - dx_t) mod W]      transpose = 1: offsets negated
// 32 x 32 output tile per CTA, halo staged in shared memory.  grid = (W/32, H/32, planes), block (32, 8).
__global__ void __launch_bounds__(256)
fh_k_conv_taps64(const double* __restrict__ in, double* __restrict__ out, FhTaps taps, int H, int W, int transpose) {
  FH_DYN_SMEM(double, sm);
  // taps staged per CTA in chunks of FH_TAPS64_CHUNK: (offset inside the tile, float64 weight) -- the first version read
  // dy / dx / w from global memory and widened the weight inside the loop (38 us per call for the 217-tap motion PSF)
  __shared__ int s_off[FH_TAPS64_CHUNK];
  __shared__ double s_w[FH_TAPS64_CHUNK];
  const int oy0 = transpose ? taps.ymin : -taps.ymax, oy1 = transpose ? taps.ymax : -taps.ymin;
  const int ox0 = transpose ? taps.xmin : -taps.xmax, ox1 = transpose ? taps.xmax : -taps.xmin;
  const int th = 32 + (oy1 - oy0), tw = 32 + (ox1 - ox0);
  const int tws = tw | 1;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const size_t plane = (size_t)blockIdx.z * H * W;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int base = tid; base < th * tws; base += 256 * 8) {        // 8 halo loads in flight per thread
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + 256 * u, r = idx / tws, c = idx - r * tws;
      v[u] = 0.0;
      if (r < th && c < tw) {
        int gi = (i0 + oy0 + r) % H, gj = (j0 + ox0 + c) % W;
        gi += (gi < 0) ? H : 0; gj += (gj < 0) ? W : 0;
        v[u] = in[plane + (size_t)gi * W + gj];
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + 256 * u;
      if (idx < th * tws) sm[idx] = v[u];
    }
  }
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const double* base = sm + threadIdx.y * tws + threadIdx.x;
  for (int t0 = 0; t0 < taps.n; t0 += FH_TAPS64_CHUNK) {
    const int nt = min(FH_TAPS64_CHUNK, taps.n - t0);
    __syncthreads();                                              // tile complete / previous chunk consumed
    for (int t = tid; t < nt; t += 256) {
      const int oy = transpose ? taps.dy[t0 + t] : -taps.dy[t0 + t];
      const int ox = transpose ? taps.dx[t0 + t] : -taps.dx[t0 + t];
      s_off[t] = (oy - oy0) * tws + (ox - ox0);
      s_w[t] = (double)taps.w[t0 + t];
    }
    __syncthreads();
    for (int t = 0; t < nt; ++t) {
      const double wv = s_w[t];
      const double* p = base + s_off[t];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fma(wv, p[8 * q * tws], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) out[plane + (size_t)(i0 + threadIdx.y + 8 * q) * W + j0 + threadIdx.x] = acc[q];
}
// x4 super-resolution in float64, once per solve (right-hand side b = y - A mu and result A^T u; utils_sisr.py:44-61).
//   forward:   out[I][J] = sum_t w_t in[(sf I - dy_t) mod H][(sf J - dx_t) mod W]          (blur, then keep every sf-th pixel)
//   transpose: out[i][j] = sum_t w_t up(in)[(i + dy_t) mod H][(j + dx_t) mod W],  up = zero insertion
// The first versions walked the whole tap list per output with one accumulator (625 dependent FMAs and as many index
// loads; the transpose tested every tap for divisibility by sf): 60 and 211 us per call, 16 ms of a 70 ms image.
// forward: one WARP per measurement pixel, the taps spread over its lanes, shuffle reduction.
__global__ void __launch_bounds__(256)
fh_k_sr_forward64(const double* __restrict__ in, double* __restrict__ out, FhTaps taps, int H, int W, int sf) {
  const int h = H / sf, w = W / sf;
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t p = blockIdx.y;
  if (idx >= h * w) return;                                   // warp-uniform
  const int I = idx / w, J = idx % w;
  const double* src = in + p * (size_t)H * W;
  double acc = 0.0;
  for (int t = lane; t < taps.n; t += 32) {
    const int gi = (sf * I - taps.dy[t] + H) & (H - 1), gj = (sf * J - taps.dx[t] + W) & (W - 1);
    acc = fma((double)taps.w[t], src[(size_t)gi * W + gj], acc);
  }
  acc = fh_warp_sum(acc);
  if (lane == 0) out[p * (size_t)h * w + idx] = acc;
}
// transpose: the PSF as a dense (ymax - ymin + 1) x (xmax - xmin + 1) float64 array in shared memory; an output pixel visits
// only the taps of its own polyphase component (dy = -i, dx = -j mod sf: 1 / sf^2 of them), in the tap list's row-major order.
#define FH_SR_PSF_MAX 33
__global__ void __launch_bounds__(256)
fh_k_sr_transpose64(const double* __restrict__ in, double* __restrict__ out, FhTaps taps, int H, int W, int sf) {
  __shared__ double s_psf[FH_SR_PSF_MAX * FH_SR_PSF_MAX];
  const int ph = taps.ymax - taps.ymin + 1, pw = taps.xmax - taps.xmin + 1;
  for (int t = threadIdx.x; t < ph * pw; t += blockDim.x) s_psf[t] = 0.0;
  __syncthreads();
  for (int t = threadIdx.x; t < taps.n; t += blockDim.x) s_psf[(taps.dy[t] - taps.ymin) * pw + (taps.dx[t] - taps.xmin)] = (double)taps.w[t];
  __syncthreads();
  const int h = H / sf, w = W / sf;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t p = blockIdx.y;
  if (idx >= H * W) return;
  const int i = idx / W, j = idx % W;
  const double* src = in + p * (size_t)h * w;
  // first offsets with (i + dy) and (j + dx) divisible by sf
  int dy0 = taps.ymin + ((((-i - taps.ymin) % sf) + sf) % sf);
  int dx0 = taps.xmin + ((((-j - taps.xmin) % sf) + sf) % sf);
  double acc = 0.0;
  for (int dy = dy0; dy <= taps.ymax; dy += sf) {
    const int gi = ((i + dy + H) & (H - 1)) / sf;
    const double* prow = s_psf + (dy - taps.ymin) * pw - taps.xmin;
    const double* srow = src + (size_t)gi * w;
    for (int dx = dx0; dx <= taps.xmax; dx += sf) {
      const int gj = ((j + dx + W) & (W - 1)) / sf;
      acc = fma(prow[dx], srow[gj], acc);
    }
  }
  out[p * (size_t)H * W + idx] = acc;
}
// b = mask .* y - mask .* mu (inpainting, conditioning_mechanisms.py:407) or y - A mu (:515, :663), in float64
__global__ void __launch_bounds__(256)
fh_k_rhs64(const double* __restrict__ y, const double* __restrict__ amu, const float* __restrict__ mask, int mask_batch, long hw,
           double* __restrict__ bvec, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t b = blockIdx.y;
  if (i >= n) return;
  const double yv = y[b * n + i], av = amu[b * n + i];
  if (mask) {
    const double mk = (double)mask[(mask_batch > 1 ? b * hw : 0) + (i % hw)];
    bvec[b * n + i] = mk * yv - mk * av;
  } else {
    bvec[b * n + i] = yv - av;
  }
}
// deterministic right-hand side of the timing entry point (fh_cg_resident_time): values in [-0.5, 0.5) from a hash of the index
template <typename T>
__global__ void __launch_bounds__(256) fh_k_fill_pattern(T* __restrict__ out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (T)((double)((((unsigned)i * 2654435761u) >> 8) & 0xffffu) / 65536.0 - 0.5);
}
__global__ void __launch_bounds__(256) fh_k_f2d(const float* __restrict__ in, double* __restrict__ out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}
__global__ void __launch_bounds__(256) fh_k_d2f(const double* __restrict__ in, float* __restrict__ out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}
// space update from float64 inputs: dx = x' - x, de = sigma^2 (mu' - mu_hat) formed in float64, stored as float32
__global__ void __launch_bounds__(256)
fh_k_su_diffs64(const double* __restrict__ x, const double* __restrict__ xn, const double* __restrict__ mu, const double* __restrict__ mun,
                double sig2, float* __restrict__ dx, float* __restrict__ de, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { dx[i] = (float)(xn[i] - x[i]); de[i] = (float)(sig2 * (mun[i] - mu[i])); }
}
// time update, last step in float64: mu' = x + sigma'^2 s'
__global__ void __launch_bounds__(256)
fh_k_tu_finish64(const double* __restrict__ x, const float* __restrict__ s_new, double sig2n, double* __restrict__ mean, double* __restrict__ score, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const double sv = (double)s_new[i]; score[i] = sv; mean[i] = fma(sig2n, sv, x[i]); }
}
// Separable circular convolution in float64 (the Gaussian PSF is rank 1): out = conv_cols(kc) conv_rows(kr) in, taps at
// offsets (a - cc) / (b - cr); transpose flips the offsets.  32 x 32 output tile per CTA, halo in shared memory, row pass
// into a second tile, column pass from it.  grid = (W/32, H/32, planes), block (32, 8).  (The 625-tap dense kernel
// fh_k_conv_taps64 took 80 us per call for the 25 x 25 Gaussian; this one does 50 taps per pixel.)  The taps are widened
// ONCE per CTA into shared memory, in the order the passes walk them (the first version converted one float per FMA:
// F2F.F64.F32 issues at 16 / clk / SM against 57.6 DFMA, and the kernel took 18.6 us), every thread carries 4 independent
// accumulators per pass, and the halo wraps by compare-and-subtract instead of an integer modulo.
__global__ void __launch_bounds__(256)
fh_k_conv_sep64(const double* __restrict__ in, double* __restrict__ out, FhSepTaps k, int H, int W, int transpose) {
  FH_DYN_SMEM(double, sm);
  const int oy0 = transpose ? -k.cc : -(k.nc - 1 - k.cc), oy1 = transpose ? (k.nc - 1 - k.cc) : k.cc;
  const int ox0 = transpose ? -k.cr : -(k.nr - 1 - k.cr), ox1 = transpose ? (k.nr - 1 - k.cr) : k.cr;
  const int th = 32 + (oy1 - oy0), tw = 32 + (ox1 - ox0), tws = tw | 1;
  double* tile = sm;                      // th x tws
  double* tmp = sm + th * tws;            // th x 33   (row-convolved)
  __shared__ double s_kr[FH_MAX_SEP_TAPS], s_kc[FH_MAX_SEP_TAPS];      // s_kr[t] multiplies tile column c + t, s_kc[t] row r + t
  const int tid = threadIdx.y * 32 + threadIdx.x;
  // offset of tap b is ox = +-(b - cr); window position t = ox - ox0  ->  b = transpose ? t : nr - 1 - t
  if (tid < k.nr) s_kr[tid] = (double)k.kr[transpose ? tid : k.nr - 1 - tid];
  if (tid >= 64 && tid - 64 < k.nc) s_kc[tid - 64] = (double)k.kc[transpose ? tid - 64 : k.nc - 1 - (tid - 64)];
  const size_t plane = (size_t)blockIdx.z * H * W;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  // halo tile: 8 loads in flight per thread (a plain r / c loop nest serialised them: one L2 round trip per tile row)
  if (tw <= 64) {
    for (int base = tid; base < th * 64; base += 256 * 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + 256 * u, r = idx >> 6, c = idx & 63;
        v[u] = 0.0;
        if (r < th && c < tw) {
          int gi = i0 + oy0 + r, gj = j0 + ox0 + c;
          gi += (gi < 0) ? H : 0; gi -= (gi >= H) ? H : 0;
          gj += (gj < 0) ? W : 0; gj -= (gj >= W) ? W : 0;
          v[u] = in[plane + (size_t)gi * W + gj];
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = base + 256 * u, r = idx >> 6, c = idx & 63;
        if (r < th && c < tw) tile[r * tws + c] = v[u];
      }
    }
  } else {
    for (int r = threadIdx.y; r < th; r += 8) {
      int gi = i0 + oy0 + r;
      gi += (gi < 0) ? H : 0; gi -= (gi >= H) ? H : 0;
      const double* srow = in + plane + (size_t)gi * W;
      for (int c = threadIdx.x; c < tw; c += 32) {
        int gj = j0 + ox0 + c;
        gj += (gj < 0) ? W : 0; gj -= (gj >= W) ? W : 0;
        tile[r * tws + c] = srow[gj];
      }
    }
  }
  __syncthreads();
  // row pass with a register window (like fh_r64_conv): item = (tile row r, 8-column segment sg), consecutive lanes take
  // consecutive rows (row stride tws is odd: conflict-free 8-byte loads); 8 + nr loads for 8 x nr FMAs
  const int nr4 = (k.nr + 3) & ~3, nc4 = (k.nc + 3) & ~3;
  for (int item = tid; item < 4 * th; item += 256) {
    const int r = item % th, sg = item / th;
    const double* trow = tile + r * tws + 8 * sg;
    const int lim = tw - 8 * sg;                           // valid window positions of this segment
    double acc[8], win[12];
#pragma unroll
    for (int e = 0; e < 8; ++e) { acc[e] = 0.0; win[e] = trow[e]; }
    for (int t0 = 0; t0 < nr4; t0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) win[8 + u] = (t0 + 8 + u < lim) ? trow[t0 + 8 + u] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double w = (t0 + u < k.nr) ? s_kr[t0 + u] : 0.0;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fma(w, win[u + e], acc[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) win[e] = win[e + 4];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) tmp[r * 33 + 8 * sg + e] = acc[e];
  }
  __syncthreads();
  // column pass: thread = (column, group of 4 consecutive output rows), window sliding down the column of tmp
  {
    const int c = threadIdx.x, g = threadIdx.y;
    const double* tcol = tmp + (4 * g) * 33 + c;
    const int lim = th - 4 * g;
    double acc[4], win[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[e] = 0.0; win[e] = tcol[e * 33]; }
    for (int t0 = 0; t0 < nc4; t0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) win[4 + u] = (t0 + 4 + u < lim) ? tcol[(t0 + 4 + u) * 33] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double w = (t0 + u < k.nc) ? s_kc[t0 + u] : 0.0;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fma(w, win[u + e], acc[e]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) win[e] = win[e + 4];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) out[plane + (size_t)(i0 + 4 * g + e) * W + j0 + c] = acc[e];
  }
}
