// Resident CG solve: ONE cooperative launch runs a whole solve of  (sigma_y^2 I + A Sigma A^T) u = b  for one image
// (conditioning_utils/cg.py:230-292 with the operators of conditioning_mechanisms.py:384-419 / 489-527), for the
// operators that fold into the DCT passes (mask, separable blur).
//
// Why: at one image per GPU (BASELINE configs[1], [2]) the five-kernel iteration of fh_cg.cuh is bound by launch
// latency and by re-reading Z (m x 786 KB) from L2 twice per iteration.  Z does not change during a solve, and the
// 148 SMs of a B200 hold 33 MB of shared memory -- more than Z at the rank a 30-step Heun run reaches (m = 32:
// 25 MB).  So the 768 rows (3 channels x 256) of the image are cut into 128 strips of 6 rows; CTA s of a 128-CTA
// cooperative grid
//   * owns pixel rows [6s, 6s+6) of x, r, p, q, b, mask                       -> registers, for the whole solve
//   * owns basis rows [6s, 6s+6) of delta (registers) and of every row of Z   -> shared memory (m x 6 KB, one TMA
//     bulk copy per row, issued once in the prologue), for the whole solve
// and the only global traffic of an iteration is the half-transformed plane T that has to be transposed between the
// row pass and the column pass of the 2-D DCT (written transposed, read back after a grid barrier: 4 x 786 KB through
// L2) plus a few KB of fp64 partial sums.  An iteration is five phases with five all-to-all exchanges between them:
//
//   P1  w = mask.p | row-blur^T p ; row DCT ; transposed store -> T                                   | exchange T
//   P2  T -> column blur^T ; column DCT = w_hat (registers) ; partial Z (delta w_hat) -> partials     | exchange dots
//   P3  coef = Gs sum(partials) ; delta (w_hat + Z^T coef) ; column IDCT ; column blur ; T' (transp.) | exchange T'
//   P4  T' -> row IDCT ; row blur | mask = t ; q = sy2 p + t ; partial p.q                             | exchange p.q
//   P5  alpha = rr / sum ; (p.Ap <= 1e-16 -> break) ; x += alpha p ; r -= alpha q ; partial r.r        | exchange r.r
//       rr' = sum ; stop test ; beta = rr'/rr ; p = r + beta p
//
// Exchanges use a flag-in-data protocol instead of grid barriers (the first version had five barriers per iteration:
// release fence + atomic + acquire polling cost 1.1-1.7 us each, 7 of 20 us): every 16-byte slot carries two 32-bit
// payload words and two copies of a sequence number (= iteration + 1), written with ONE vector store; the consumer
// polls the slot itself until both flags match.  Payload and flag arrive together, so there is no fence, no atomic and
// no separate barrier round trip, and a consumer starts as soon as ITS inputs are there.  Reuse of a slot is safe
// without further synchronisation because every exchange is all-to-all: a CTA can only pass exchange k once every CTA
// has produced its part of k, i.e. has finished reading exchange k-1 (see the hazard notes in the kernel).
// All scalars are fixed-order sums of the same partials in every CTA (bit-identical), so every CTA takes the same
// branch and the loop needs no flag.  The initial residual r0 = b - A_op b (x0 = b, cg.py:233-235) is iteration 0 of the
// same loop.  Polls are bounded (__trap instead of a hang).
#pragma once
#include "fh_cg.cuh"

#define FH_RES_ROWS 6
#define FH_RES_SLICE (FH_RES_ROWS * FH_N)          // floats per strip (1536)
#define FH_RES_CTAS_PER_PLANE_X3 128               // 3 * 256 / 6
#define FH_RES_MAXM 32                             // rows of Z one CTA can keep in shared memory (32 x 6 KB = 192 KB)
#define FH_RES_TAPS (FH_MAX_SEP_TAPS + 8)

struct FhResSmem {
  float2 X[FH_RES_ROWS / 2][FH_FFT_LD];            // FFT ping-pong; X doubles as the strip's input rows (stride FH_RS),
  float2 Y[FH_RES_ROWS / 2][FH_FFT_LD];            // Y as its output rows (see fh_dct_strip_core)
  float2 tw[3 * 3 * 64];
  float2 post[FH_N];
  float2 cw[FH_N];
  float taps[4][FH_RES_TAPS];                      // row A^T, column A^T, column A, row A (centre padded to a multiple of 4)
  float s_w[FH_PASS_THREADS / 32][FH_MAXM];
  double s_sum[FH_MAXM];
  double s_red[36];                                // [0..32] fh_block_sum scratch, [33], [34] broadcast slots of fh_ll_reduce
  float s_c[FH_MAXM];
  fh_mbar_t bar;
  unsigned long long pad_;
};

struct FhCgInfo { double residual_norm; int niter; int optimal; };      // == fh_cg_info of include/free_hunch_b200.h

struct FhCgResident {
  const float* bvec;           // (d) right-hand side b (pixel layout); x0 = b
  float* x_out;                // (d) solution
  uint4* ll_T[2];              // (d/2) slots each: half-transformed plane, forward / inverse (slot = 2 floats + 2 flags)
  uint4* ll_dots;              // [FH_MAXM][grid] slots: per-CTA partial of (Z (delta w_hat))_i as a double
  uint4* ll_sum;               // [FH_MAXM] slots, FH_LL_SPREAD apart: (Z (delta w_hat))_i summed over the grid by row i's owner CTA
  uint4* ll_a;                 // [grid] slots, FH_LL_SPREAD apart: per-CTA partial of p.q  (iteration 0: |r0|^2)
  uint4* ll_b;                 // [grid] slots, FH_LL_SPREAD apart: per-CTA partial of r.r  (iteration 0: |b|^2)
  unsigned* grid_counter;      // arrival counter of the grid barrier, zeroed by the host before the launch
  unsigned seq_base;           // sequence numbers of this launch are seq_base + 1 ... ; every flag in the buffers is
                               // <= seq_base (the host hands out increasing ranges and zeroes the buffers on wrap-around)
  double rtol, atol;
  int use_stop;                // 0: run exactly `last_allowed` iterations
  int image;                   // which image of the batch (offsets into the state / scalars)
  FhCgInfo* info_out;          // optional device record of the outcome (fh_cg_solve_async); written by CTA 0 at the end
  int timing;                  // fh_cg_resident_time only: keep iterating through the p.Ap <= 1e-16 break
};

// ---- flag-in-data exchange ---------------------------------------------------------------------------------
// slot = {payload0, seq, payload1, seq}: one 16-byte vector store / load (each 8-byte half carries its own flag, so
// the protocol holds even if the memory system splits the access).
#define FH_LL_SPIN_LIMIT 8000000u
#ifdef FH_EMU
#include <atomic>
static inline void fh_ll_store(uint4* p, unsigned a, unsigned b, unsigned seq) {
  volatile unsigned* q = reinterpret_cast<volatile unsigned*>(p);
  q[0] = a; q[2] = b;
  std::atomic_thread_fence(std::memory_order_release);
  q[1] = seq; q[3] = seq;
}
static inline uint4 fh_ll_load(const uint4* p) {
  volatile const unsigned* q = reinterpret_cast<volatile const unsigned*>(p);
  uint4 v;
  v.y = q[1]; v.w = q[3];
  std::atomic_thread_fence(std::memory_order_acquire);
  v.x = q[0]; v.z = q[2];
  return v;
}
static inline void fh_ll_relax() { emu::relax(); }
static inline void fh_ll_fail() { std::abort(); }
static inline unsigned fh_f2u(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float fh_u2f(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline void fh_d2u(double d, unsigned& lo, unsigned& hi) { unsigned long long b; std::memcpy(&b, &d, 8); lo = (unsigned)b; hi = (unsigned)(b >> 32); }
static inline double fh_u2d(unsigned lo, unsigned hi) { unsigned long long b = ((unsigned long long)hi << 32) | lo; double d; std::memcpy(&d, &b, 8); return d; }
#else
__device__ __forceinline__ void fh_ll_store(uint4* p, unsigned a, unsigned b, unsigned seq) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(seq), "r"(b), "r"(seq) : "memory");
}
__device__ __forceinline__ uint4 fh_ll_load(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fh_ll_relax() {}
__device__ __forceinline__ void fh_ll_fail() { __trap(); }
__device__ __forceinline__ unsigned fh_f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float fh_u2f(unsigned u) { return __uint_as_float(u); }
__device__ __forceinline__ void fh_d2u(double d, unsigned& lo, unsigned& hi) { lo = (unsigned)__double2loint(d); hi = (unsigned)__double2hiint(d); }
__device__ __forceinline__ double fh_u2d(unsigned lo, unsigned hi) { return __hiloint2double((int)hi, (int)lo); }
#endif

// N slots p[0], p[stride], ... : all loads are issued before the first flag is looked at; slots that are not there
// yet are re-polled one by one (bounded).
template <int N>
__device__ __forceinline__ void fh_ll_wait(const uint4* p, int stride, unsigned seq, uint4 (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = fh_ll_load(p + (size_t)i * stride);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    unsigned spins = 0;
    while (v[i].y != seq || v[i].w != seq) {
      if (++spins > FH_LL_SPIN_LIMIT) fh_ll_fail();     // seconds: a producer is missing -> fail the launch, do not hang
      fh_ll_relax();
      v[i] = fh_ll_load(p + (size_t)i * stride);
    }
  }
}

// ---- grid barrier (used for the two plane exchanges when FH_RES_T_LL == 0) ---------------------------------------
#ifndef FH_RES_T_LL
#define FH_RES_T_LL 0
#endif
#ifdef FH_EMU
static inline void fh_grid_sync(unsigned*, unsigned&, unsigned) {
  __syncthreads();
  if (threadIdx.x == 0) emu::g_grid_bar.wait();
  __syncthreads();
}
template <typename T> static inline T fh_ld_cg(const T* p) { return *p; }
#else
// All threads call it.  Thread 0 arrives with release semantics (the CTA's earlier writes, ordered by the bar.sync,
// become visible device-wide) and polls with acquire loads until every CTA of the grid has arrived.
__device__ __forceinline__ void fh_grid_sync(unsigned* counter, unsigned& target, unsigned nblocks) {
  target += nblocks;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned v = 0;
    for (unsigned spins = 0;; ++spins) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if (spins > FH_LL_SPIN_LIMIT) __trap();          // seconds: a CTA is missing -> fail the launch, do not hang
    }
  }
  __syncthreads();
}
template <typename T> __device__ __forceinline__ T fh_ld_cg(const T* p) { return __ldcg(p); }
#endif
// plain (flag-less) variants of the plane exchange: float2 per row pair, read past L1 after the grid barrier
__device__ __forceinline__ void fh_res_get_cg(const float* src, float (&r)[FH_RES_ROWS]) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    const float2 v = fh_ld_cg(reinterpret_cast<const float2*>(src) + threadIdx.x + FH_PASS_THREADS * h);
    r[2 * h] = v.x; r[2 * h + 1] = v.y;
  }
}
__device__ __forceinline__ void fh_res_store_t(const float* sm_rows, float* __restrict__ T, int s) {
  const int k = threadIdx.x;
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    const int g = FH_RES_ROWS * s + 2 * h;
    float* d = T + (size_t)(g >> 8) * FH_N * FH_N + (size_t)k * FH_N + (g & (FH_N - 1));
    *reinterpret_cast<float2*>(d) = make_float2(sm_rows[(2 * h) * FH_RS + k], sm_rows[(2 * h + 1) * FH_RS + k]);
  }
}

// The thread's 6 elements (3 pairs: slots #t, #t+256, #t+512) of its strip of an exchanged plane.
__device__ __forceinline__ void fh_res_get_ll(const uint4* strip, unsigned seq, float (&r)[FH_RES_ROWS]) {
  uint4 v[FH_RES_ROWS / 2];
  fh_ll_wait<FH_RES_ROWS / 2>(strip + threadIdx.x, FH_PASS_THREADS, seq, v);
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) { r[2 * h] = fh_u2f(v[h].x); r[2 * h + 1] = fh_u2f(v[h].z); }
}

// Transposed store of strip s (global rows 6s .. 6s+5 of the 768) into the exchange plane: element (row g, column k)
// goes to row (plane(g), k), column g & 255 of the transposed plane; the pair (g, g+1) is one slot.  Row pairs never
// straddle a plane (256 is even); the strip itself may.
__device__ __forceinline__ void fh_res_store_t_ll(const float* sm_rows, uint4* T, int s, unsigned seq) {
  const int k = threadIdx.x;
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    const int g = FH_RES_ROWS * s + 2 * h;
    uint4* d = T + ((size_t)(g >> 8) * FH_N + k) * (FH_N / 2) + ((g & (FH_N - 1)) >> 1);
    fh_ll_store(d, fh_f2u(sm_rows[(2 * h) * FH_RS + k]), fh_f2u(sm_rows[(2 * h + 1) * FH_RS + k]), seq);
  }
}

// Slots that many CTAs poll at once sit FH_LL_SPREAD slots (128 bytes) apart, so that they land in different L2 slices:
// with 2 KB of contiguous slots polled by 16 384 threads one slice serialises everything (measured: 10 us per exchange).
#define FH_LL_SPREAD 8
// Sum of one double per CTA over the grid: thread t < nblk polls slot t, fixed-order block sum (bit-identical in
// every CTA).  Contains __syncthreads.  (Polling with warp 0 only -- 4 slots per lane, shuffle reduce, one barrier --
// measured slower: 20.0 instead of 18.7 us per iteration.)
__device__ __forceinline__ double fh_ll_reduce(const uint4* slots, int nblk, unsigned seq, double* s_red, int /*which*/) {
  double v = 0.0;
  for (int i = threadIdx.x; i < nblk; i += FH_PASS_THREADS) {
    uint4 w[1];
    fh_ll_wait<1>(slots + (size_t)i * FH_LL_SPREAD, 0, seq, w);
    v += fh_u2d(w[0].x, w[0].z);
  }
  return fh_block_sum(v, s_red);
}
__device__ __forceinline__ void fh_ll_put_double(uint4* slot, double v, unsigned seq) {
  unsigned lo, hi;
  fh_d2u(v, lo, hi);
  fh_ll_store(slot, lo, hi, seq);
}

// the thread's 6 elements of a contiguous (unpadded) 1536-float slice in shared memory: conflict-free LDS.64
__device__ __forceinline__ void fh_strip_get_flat(const float* src, float (&r)[FH_RES_ROWS]) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
    const float2 v = reinterpret_cast<const float2*>(src)[threadIdx.x + FH_PASS_THREADS * h];
    r[2 * h] = v.x; r[2 * h + 1] = v.y;
  }
}

// In-place circular correlation along the 6 rows (register window, 8 outputs per thread, warps 0..5 active).
// `taps`: n4 (multiple of 4) taps with centre c4 (multiple of 4).  Contains two __syncthreads.
__device__ __forceinline__ void fh_res_conv(float* sm_rows, const float* __restrict__ taps, int n4, int c4) {
  const int row = threadIdx.x >> 5, l = threadIdx.x & 31;
  const bool act = row < FH_RES_ROWS;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (act) {
    const float* rr = sm_rows + row * FH_RS;
    const int base = 8 * l - c4;
    float4 f0 = *reinterpret_cast<const float4*>(rr + (base & (FH_N - 1)));
    float4 f1 = *reinterpret_cast<const float4*>(rr + ((base + 4) & (FH_N - 1)));
    for (int a0 = 0; a0 < n4; a0 += 4) {
      const float4 f2 = *reinterpret_cast<const float4*>(rr + ((base + a0 + 8) & (FH_N - 1)));
      const float4 tk = *reinterpret_cast<const float4*>(taps + a0);
      const float win[12] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w};
      const float tt[4] = {tk.x, tk.y, tk.z, tk.w};
#pragma unroll
      for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(tt[aa], win[aa + e], acc[e]);
      f0 = f1; f1 = f2;
    }
  }
  __syncthreads();
  if (act) {
    float* wr = sm_rows + row * FH_RS + 8 * l;
    *reinterpret_cast<float4*>(wr) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(wr + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
}

// Stage one tap set: forward = 0 (A^T): t = k, centre c; forward = 1 (A): t = reversed k, centre n-1-c; then shifted
// right by (4 - centre % 4) % 4 zeros so that the centre is a multiple of 4.  Returns (n4 << 8) | c4.  Caller syncs.
__device__ __forceinline__ int fh_res_stage_taps(float* s_taps, const float* __restrict__ k, int n, int c, int forward) {
  const int cen = forward ? n - 1 - c : c;
  const int sh = (4 - (cen & 3)) & 3;
  const int n4 = (n + sh + 3) & ~3;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int a = i - sh;
    s_taps[i] = (a >= 0 && a < n) ? (forward ? k[n - 1 - a] : k[a]) : 0.f;
  }
  return (n4 << 8) | (cen + sh);
}

// Sum over the 32 lanes of 32 values per lane with 31 shuffles: afterwards lane l holds the warp total of v[l] in v[0].
__device__ __forceinline__ float fh_warp_transpose_sum32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < o; ++k) {
      const float send = hi ? v[k] : v[k + o];
      const float keep = hi ? v[k + o] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// the thread's 6 elements of a Z row slice that is NOT resident: straight from global memory (L2), read-only path
__device__ __forceinline__ void fh_strip_get_global(const float* __restrict__ src, float (&r)[FH_RES_ROWS]) {
#pragma unroll
  for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
#ifdef FH_EMU
    const float2 v = reinterpret_cast<const float2*>(src)[threadIdx.x + FH_PASS_THREADS * h];
#else
    const float2 v = __ldg(reinterpret_cast<const float2*>(src) + threadIdx.x + FH_PASS_THREADS * h);
#endif
    r[2 * h] = v.x; r[2 * h + 1] = v.y;
  }
}

// grid = C*256/6 CTAs (cooperative launch), block = 256, dynamic shared memory = sizeof(FhResSmem) + min(m, 32) * 6 KB.
// BIG = false: m <= 32, every row of Z resident.  BIG = true: 32 < m <= 64 (a 100-step Euler run reaches 56): rows
// 0..31 resident, rows 32..m-1 re-read from L2 in the two low-rank phases of every iteration (their slices are only
// read, each by one CTA: plain read-only loads, 12 in flight per thread, no staging).
template <bool BIG>
__global__ void __launch_bounds__(FH_PASS_THREADS, 1)
fh_k_cg_resident(FhCgFused P, FhCgResident R) {
  FH_DYN_SMEM(unsigned char, smem_raw);
  FhResSmem& sm = *reinterpret_cast<FhResSmem*>(smem_raw);
  float* Zs = reinterpret_cast<float*>(smem_raw + sizeof(FhResSmem));       // [m][1536]
  // rows_a aliases the FFT buffer X and rows_b aliases Y: a strip is put into rows_a only after the barrier that follows
  // the last read of X, and Y is rewritten by the next transform only after the barrier that follows the put.
  float* rows_a = reinterpret_cast<float*>(sm.X);
  float* rows_b = reinterpret_cast<float*>(sm.Y);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x;
  const unsigned nblk = gridDim.x;
  const int m = P.m;
  const int mr = BIG ? FH_RES_MAXM : m;                  // rows of Z resident in shared memory
  const long d = (long)nblk * FH_RES_SLICE;
  const size_t img = (size_t)R.image * d;
  const size_t soff = (size_t)s * FH_RES_SLICE;

  // ---- prologue: Z slices by TMA (one 6 KB bulk copy per row), tables, taps, per-solve constants in registers ----
  if (tid == 0) { fh_mbar_init(&sm.bar, 1); fh_mbar_fence_init(); }
  __syncthreads();
  const float* zsrc = P.Z + (size_t)R.image * FH_MAXM * d + soff;        // this CTA's slice of row 0 of Z
  if (m > 0 && warp == 0) {
    if (lane == 0) fh_mbar_expect_tx(&sm.bar, (unsigned)mr * FH_RES_SLICE * 4u);
    __syncwarp();
    for (int i = lane; i < mr; i += 32) fh_tma_load_1d(Zs + (size_t)i * FH_RES_SLICE, zsrc + (size_t)i * d, FH_RES_SLICE * 4u, &sm.bar);
  }
  for (int i = tid; i < 3 * 3 * 64; i += FH_PASS_THREADS) sm.tw[i] = P.tab.tw256[i];
  for (int i = tid; i < FH_N; i += FH_PASS_THREADS) { sm.post[i] = P.tab.post[i]; sm.cw[i] = P.tab.cw[i]; }
  int tp[4] = {0, 0, 0, 0};
  if (P.fuse == 2) {
    tp[0] = fh_res_stage_taps(sm.taps[0], P.kr, P.nr, P.cr, 0);
    tp[1] = fh_res_stage_taps(sm.taps[1], P.kc, P.nc, P.cc, 0);
    tp[2] = fh_res_stage_taps(sm.taps[2], P.kc, P.nc, P.cc, 1);
    tp[3] = fh_res_stage_taps(sm.taps[3], P.kr, P.nr, P.cr, 1);
  }
  float bv[FH_RES_ROWS], dl[FH_RES_ROWS], mk[FH_RES_ROWS];
  fh_g_get<FH_RES_ROWS>(R.bvec + soff, bv);
  fh_g_get<FH_RES_ROWS>(P.delta + img + soff, dl);
#pragma unroll
  for (int e = 0; e < FH_RES_ROWS; ++e) mk[e] = 1.f;
  if (P.fuse == 1) {
    // the mask is one 256 x 256 plane per image, shared by the channels: global row g -> plane row g & 255
    const float* mp = P.mask + (size_t)(P.mask_batch > 1 ? R.image : 0) * FH_N * FH_N;
#pragma unroll
    for (int h = 0; h < FH_RES_ROWS / 2; ++h) {
      const int e = 2 * (tid + FH_PASS_THREADS * h);                    // element offset inside the strip
      const int g = FH_RES_ROWS * s + (e >> 8);
      const float2 v = *reinterpret_cast<const float2*>(mp + (size_t)(g & (FH_N - 1)) * FH_N + (e & (FH_N - 1)));
      mk[2 * h] = v.x; mk[2 * h + 1] = v.y;
    }
  }
  // coefficient stage: thread (i = tid / CP, part = tid % CP) owns Gs[i][part + CP u], u < NG  (32 rows x 8 parts x 4
  // columns, or 64 rows x 4 parts x 16 columns)
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
  float* Tp = P.T + img;                                                   // plain plane (FH_RES_T_LL == 0)
  const uint4* my_T0 = R.ll_T[0] + (size_t)s * (FH_RES_SLICE / 2);          // this CTA's strip of the exchanged planes
  const uint4* my_T1 = R.ll_T[1] + (size_t)s * (FH_RES_SLICE / 2);
  __syncthreads();

  float x[FH_RES_ROWS], r[FH_RES_ROWS], p[FH_RES_ROWS], q[FH_RES_ROWS];
#pragma unroll
  for (int e = 0; e < FH_RES_ROWS; ++e) { x[e] = bv[e]; p[e] = bv[e]; r[e] = 0.f; q[e] = 0.f; }
  double rr = 0.0, stop2 = -1.0, resid2 = 0.0;
  int niter = 0, optimal = 0;
  bool z_ready = (m == 0);

#define FH_RTRACE(slot) do { if (it == 3) FH_TRACE(P, 5, slot); } while (0)
  for (int it = 0;; ++it) {
    const unsigned seq = R.seq_base + (unsigned)it + 1u;
    // Slot reuse: T (forward) is rewritten in P1 of it+1, after this CTA has passed the r.r exchange of `it`, which every
    // CTA joins only after its own P2 read of T; likewise T' (P3 / P4, p.q exchange), the dot slots (P2 / P3, then p.q and
    // r.r), ll_a (P4 / P5, r.r) and ll_b (P5, then the dot exchange of it+1).
    // ------------------------------------------------------------------ P1: rows, forward
    FH_RTRACE(0);
    {
      float w[FH_RES_ROWS];
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) w[e] = p[e] * mk[e];
      fh_strip_put<FH_RES_ROWS>(rows_a, w);
      __syncthreads();
      if (P.fuse == 2) fh_res_conv(rows_a, sm.taps[0], tp[0] >> 8, tp[0] & 255);
      fh_dct_strip_core<false, FH_RES_ROWS, false>(rows_a, rows_b, sm.X, sm.Y, sm.tw, sm.post);
      __syncthreads();
#if FH_RES_T_LL
      fh_res_store_t_ll(rows_b, R.ll_T[0], s, seq);
#else
      fh_res_store_t(rows_b, Tp, s);
#endif
    }
    FH_RTRACE(1);
#if !FH_RES_T_LL
    fh_grid_sync(R.grid_counter, bar_target, nblk);
#endif
    // ------------------------------------------------------------------ P2: columns, forward; partial Z (delta w_hat)
    float wh[FH_RES_ROWS];
    {
      float t[FH_RES_ROWS];
#if FH_RES_T_LL
      fh_res_get_ll(my_T0, seq, t);
#else
      fh_res_get_cg(Tp + soff, t);
#endif
      FH_RTRACE(2);
      fh_strip_put<FH_RES_ROWS>(rows_a, t);
      __syncthreads();
      if (P.fuse == 2) fh_res_conv(rows_a, sm.taps[1], tp[1] >> 8, tp[1] & 255);
      fh_dct_strip_core<false, FH_RES_ROWS, false>(rows_a, rows_b, sm.X, sm.Y, sm.tw, sm.post);
      __syncthreads();
      fh_strip_get<FH_RES_ROWS>(rows_b, wh);
    }
    FH_RTRACE(3);
    if (m > 0) {
      if (!z_ready) { fh_mbar_wait(&sm.bar, 0); z_ready = true; }
      float wd[FH_RES_ROWS];
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) wd[e] = wh[e] * dl[e];
      {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          acc[i] = 0.f;
          if (i < mr) {
            float z[FH_RES_ROWS];
            fh_strip_get_flat(Zs + (size_t)i * FH_RES_SLICE, z);
#pragma unroll
            for (int e = 0; e < FH_RES_ROWS; ++e) acc[i] = fmaf(z[e], wd[e], acc[i]);
          }
        }
        sm.s_w[warp][lane] = fh_warp_transpose_sum32(acc);
      }
      if (BIG) {
        float acc[32];
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += 4) {                 // 4 rows = 12 independent 8-byte loads in flight per thread
          float z[4][FH_RES_ROWS];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (mr + i0 + u < m) fh_strip_get_global(zsrc + (size_t)(mr + i0 + u) * d, z[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[i0 + u] = 0.f;
            if (mr + i0 + u < m) {
#pragma unroll
              for (int e = 0; e < FH_RES_ROWS; ++e) acc[i0 + u] = fmaf(z[u][e], wd[e], acc[i0 + u]);
            }
          }
        }
        sm.s_w[warp][32 + lane] = fh_warp_transpose_sum32(acc);
      }
      __syncthreads();
      if (tid < m) {
        double t = 0.0;
#pragma unroll
        for (int wv = 0; wv < FH_PASS_THREADS / 32; ++wv) t += (double)sm.s_w[wv][tid];
        fh_ll_put_double(R.ll_dots + (size_t)tid * nblk + s, t, seq);
      }
      FH_RTRACE(4);
      // ---------------------------------------------------------------- P3: coef = Gs sum(partials); w_hat + Z^T coef
      {
        // Two hops, each a few KB: row i's partials (one slot per CTA, contiguous) are summed in a fixed order by warp 0
        // of its owner CTA (i * grid / 32: spread over the chip), which publishes the total; then warp 0 of EVERY CTA
        // fetches the m totals.  (One hop -- every CTA summing all grid x m partials itself -- moves 8 MB through L2 per
        // iteration and took 2.2 us.)
        constexpr int OROWS = BIG ? FH_MAXM : FH_RES_MAXM;
        const int own_step = (int)nblk / OROWS > 0 ? (int)nblk / OROWS : 1;
        if (warp == 0) {
          if (s % own_step == 0 && s / own_step < m) {
            const int i = s / own_step;
            const uint4* col = R.ll_dots + (size_t)i * nblk;
            double t = 0.0;
            for (int k0 = lane; k0 < (int)nblk; k0 += 128) {
              if (k0 + 96 < (int)nblk) {
                uint4 v[4];
                fh_ll_wait<4>(col + k0, 32, seq, v);
#pragma unroll
                for (int u = 0; u < 4; ++u) t += fh_u2d(v[u].x, v[u].z);
              } else {
                for (int k = k0; k < (int)nblk; k += 32) {
                  uint4 v[1];
                  fh_ll_wait<1>(col + k, 0, seq, v);
                  t += fh_u2d(v[0].x, v[0].z);
                }
              }
            }
            t = fh_warp_sum(t);
            if (lane == 0) fh_ll_put_double(R.ll_sum + (size_t)i * FH_LL_SPREAD, t, seq);
          }
#pragma unroll
          for (int g = 0; g < (BIG ? 2 : 1); ++g) {
            const int row = lane + 32 * g;
            double tot = 0.0;
            if (row < m) {
              uint4 v[1];
              fh_ll_wait<1>(R.ll_sum + (size_t)row * FH_LL_SPREAD, 0, seq, v);
              tot = fh_u2d(v[0].x, v[0].z);
            }
            sm.s_sum[row] = tot;
          }
        }
        FH_RTRACE(5);
        __syncthreads();
        double a = 0.0;
#pragma unroll
        for (int u = 0; u < NG; ++u) a += gs[u] * sm.s_sum[cpart + CP * u];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (!BIG) a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (cpart == 0 && ci < CROWS) sm.s_c[ci] = (float)a;
        __syncthreads();
      }
      FH_RTRACE(6);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i < mr) {
          const float c = sm.s_c[i];
          float z[FH_RES_ROWS];
          fh_strip_get_flat(Zs + (size_t)i * FH_RES_SLICE, z);
#pragma unroll
          for (int e = 0; e < FH_RES_ROWS; ++e) wh[e] = fmaf(c, z[e], wh[e]);
        }
      }
      if (BIG) {
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += 4) {
          float z[4][FH_RES_ROWS];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (mr + i0 + u < m) fh_strip_get_global(zsrc + (size_t)(mr + i0 + u) * d, z[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (mr + i0 + u < m) {
              const float c = sm.s_c[mr + i0 + u];
#pragma unroll
              for (int e = 0; e < FH_RES_ROWS; ++e) wh[e] = fmaf(c, z[u][e], wh[e]);
            }
          }
        }
      }
    }
    FH_RTRACE(7);
    {
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) wh[e] *= dl[e];
      fh_strip_put<FH_RES_ROWS>(rows_a, wh);
      __syncthreads();
      fh_dct_strip_core<true, FH_RES_ROWS, true>(rows_a, rows_b, sm.X, sm.Y, sm.tw, sm.cw);
      __syncthreads();
      if (P.fuse == 2) fh_res_conv(rows_b, sm.taps[2], tp[2] >> 8, tp[2] & 255);
#if FH_RES_T_LL
      fh_res_store_t_ll(rows_b, R.ll_T[1], s, seq);
#else
      fh_res_store_t(rows_b, Tp, s);
#endif
    }
    FH_RTRACE(8);
#if !FH_RES_T_LL
    fh_grid_sync(R.grid_counter, bar_target, nblk);
#endif
    // ------------------------------------------------------------------ P4: rows, inverse; q = sy2 p + t; partial p.q
    {
      float t[FH_RES_ROWS];
#if FH_RES_T_LL
      fh_res_get_ll(my_T1, seq, t);
#else
      fh_res_get_cg(Tp + soff, t);
#endif
      FH_RTRACE(9);
      fh_strip_put<FH_RES_ROWS>(rows_a, t);
      __syncthreads();
      fh_dct_strip_core<true, FH_RES_ROWS, true>(rows_a, rows_b, sm.X, sm.Y, sm.tw, sm.cw);
      __syncthreads();
      if (P.fuse == 2) fh_res_conv(rows_b, sm.taps[3], tp[3] >> 8, tp[3] & 255);
      fh_strip_get<FH_RES_ROWS>(rows_b, t);
      float a32 = 0.f;
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) {
        q[e] = fmaf(P.sy2, p[e], t[e] * mk[e]);
        a32 = fmaf(p[e], q[e], a32);
      }
      if (it == 0) {
        // initial residual r0 = b - A_op x0 (x0 = b); |b|^2 and |r0|^2; p1 = r0     (cg.py:233-237)
        float rr32 = 0.f, bb32 = 0.f;
#pragma unroll
        for (int e = 0; e < FH_RES_ROWS; ++e) {
          r[e] = bv[e] - q[e];
          rr32 = fmaf(r[e], r[e], rr32);
          bb32 = fmaf(bv[e], bv[e], bb32);
          p[e] = r[e];
        }
        const double t1 = fh_block_sum((double)rr32, sm.s_red);
        const double t2 = fh_block_sum((double)bb32, sm.s_red);
        if (tid == 0) { fh_ll_put_double(R.ll_a + (size_t)s * FH_LL_SPREAD, t1, seq); fh_ll_put_double(R.ll_b + (size_t)s * FH_LL_SPREAD, t2, seq); }
        rr = fh_ll_reduce(R.ll_a, (int)nblk, seq, sm.s_red, 0);
        const double bb = fh_ll_reduce(R.ll_b, (int)nblk, seq, sm.s_red, 1);
        const double stop = fmax(R.rtol * sqrt(bb), R.atol);
        stop2 = R.use_stop ? stop * stop : -1.0;
        resid2 = rr;
        if (s == 0 && tid == 0) { P.sc[R.image].bnorm2 = bb; P.sc[R.image].stop2 = stop2; }
        continue;
      }
      const double tot = fh_block_sum((double)a32, sm.s_red);
      if (tid == 0) fh_ll_put_double(R.ll_a + (size_t)s * FH_LL_SPREAD, tot, seq);
    }
    FH_RTRACE(10);
    // ------------------------------------------------------------------ P5: step, residual, stop test, new direction
    {
      const double pAp = fh_ll_reduce(R.ll_a, (int)nblk, seq, sm.s_red, 0);
      FH_RTRACE(12);
      if (pAp <= 1e-16 && !R.timing) { niter = it; break; }              // cg.py:250
      const float alpha = (float)(rr / pAp);
      float rr32 = 0.f;
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) {
        x[e] = fmaf(alpha, p[e], x[e]);
        r[e] = fmaf(-alpha, q[e], r[e]);
        rr32 = fmaf(r[e], r[e], rr32);
      }
      const double tot = fh_block_sum((double)rr32, sm.s_red);
      if (tid == 0) fh_ll_put_double(R.ll_b + (size_t)s * FH_LL_SPREAD, tot, seq);
      FH_RTRACE(13);
      const double rr_new = fh_ll_reduce(R.ll_b, (int)nblk, seq, sm.s_red, 1);
      FH_RTRACE(15);
      resid2 = rr_new;
      niter = it;
      const bool converged = (stop2 >= 0.0) && (rr_new <= stop2);
      if (converged) optimal = 1;
      if (converged || it >= P.last_allowed) break;                      // cg.py:258-263
      const float beta = (float)(rr_new / rr);
      rr = rr_new;
#pragma unroll
      for (int e = 0; e < FH_RES_ROWS; ++e) p[e] = fmaf(beta, p[e], r[e]);
    }
  }
  if (!z_ready) fh_mbar_wait(&sm.bar, 0);                  // never leave with bulk copies in flight
  fh_g_put<FH_RES_ROWS>(R.x_out + soff, x);
  if (s == 0 && tid == 0) {
    FhCgScalars* w = P.sc + R.image;
    w->resid2 = resid2;
    w->rr[0] = rr; w->rr[1] = rr;
    w->niter = niter;
    w->optimal = optimal;
    fh_set_done(w, niter + 1);
    if (R.info_out) { R.info_out->residual_norm = sqrt(resid2); R.info_out->niter = niter; R.info_out->optimal = optimal; }
  }
}
