// Host side of libfreehunch_b200: C ABI (include/free_hunch_b200.h) over the kernels in fh_dct.cuh and
// fh_kernels.cuh.  No torch types, no hidden host<->device copies on the hot path except the CG
// convergence poll (one 4-byte flag per image every few iterations).
#include "fh_common.cuh"
#include "fh_dct.cuh"
#include "fh_kernels.cuh"
#include "fh_cg.cuh"
#include "fh_cg_resident.cuh"
#include "fh_cg_res64.cuh"
#include "fh_upd64.cuh"
#include "fh_mech.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/free_hunch_b200.h"

long long fh_g_launches = 0;
static thread_local std::string g_err;

// ------------------------------------------------------------------------------------------- profiling
#ifndef FH_EMU
#include <map>
int fh_g_prof_on = 0;
namespace {
struct ProfPair { const char* name; cudaEvent_t a, b; };
std::vector<ProfPair> g_prof_pairs;
std::vector<cudaEvent_t> g_prof_free;
std::map<std::string, std::pair<long long, double>> g_prof_acc;   // name -> (launches, total ms)
cudaEvent_t prof_event() {
  if (!g_prof_free.empty()) { cudaEvent_t e = g_prof_free.back(); g_prof_free.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_drain() {
  for (auto& p : g_prof_pairs) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    auto& acc = g_prof_acc[p.name];
    acc.first += 1; acc.second += ms;
    g_prof_free.push_back(p.a); g_prof_free.push_back(p.b);
  }
  g_prof_pairs.clear();
}
}  // namespace
void fh_prof_begin(const char* name, cudaStream_t stream) {
  ProfPair p{name, prof_event(), prof_event()};
  cudaEventRecord(p.a, stream);
  g_prof_pairs.push_back(p);
}
void fh_prof_end(cudaStream_t stream) {
  cudaEventRecord(g_prof_pairs.back().b, stream);
  if (g_prof_pairs.size() >= (1u << 20)) prof_drain();      // draining stalls the host: only far beyond one bench step
}
extern "C" int fh_prof_enable(int on) {
  if (on && !fh_g_prof_on) g_prof_acc.clear();
  if (!on && fh_g_prof_on) prof_drain();
  fh_g_prof_on = on ? 1 : 0;
  return FH_OK;
}
// Writes "name count total_ms\n" lines; returns the number of bytes needed (call with buf = NULL to size).
extern "C" int fh_prof_report(char* buf, int buflen) {
  prof_drain();
  std::string out;
  char line[256];
  for (auto& kv : g_prof_acc) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && buflen > 0) { strncpy(buf, out.c_str(), (size_t)buflen - 1); buf[buflen - 1] = 0; }
  return (int)out.size() + 1;
}
#else
extern "C" int fh_prof_enable(int) { return FH_OK; }
extern "C" int fh_prof_report(char* buf, int buflen) { if (buf && buflen > 0) buf[0] = 0; return 1; }
#endif

#define FH_FAIL(code, ...)                         \
  do {                                             \
    char _b[512];                                  \
    snprintf(_b, sizeof(_b), __VA_ARGS__);         \
    g_err = _b;                                    \
    return (code);                                 \
  } while (0)
#define FH_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) FH_FAIL(FH_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));  \
  } while (0)
#define FH_KCHECK()                                                                            \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) FH_FAIL(FH_ECUDA, "kernel launch failed: %s", cudaGetErrorString(_e)); \
  } while (0)

#define FH_STATE_POOL 3                // parked states per context (one image's state is ~60 MB)
#define FH_STATE_SPARES 2              // spare buffer sets allocated together with a state that missed the pool
struct fh_ctx {
  int device, max_batch, C, H, W, max_rank;
  long d;
  float2 *tw256, *post, *cw;          // DCT tables
  float* plane_tmp;                   // (max_batch, d) scratch for the 2-pass DCT
  float* scratch[7];                  // (max_batch, d) general scratch vectors
  double* Gtmp;                       // (max_batch, FH_MAXM*FH_MAXM) core of a Hessian-side Sigma that differs from the state's
  double* dot_partial;                // (max_batch, nchunk, FH_MAXM+2)
  double* csum;                       // (max_batch, FH_MAXM+2)
  double* coef;                       // (max_batch, FH_MAXM)
  double* gram_partial;               // (max_batch, FH_GRAM_CTAS, LD*LD)
  double* gramK;                      // (max_batch, LD*LD)
  double *part_a, *part_b;            // (max_batch, nchunk) scalar-dot partials
  FhCgScalars* cg_sc;                 // (max_batch)
  FhCgScalars* cg_sc_host;            // pinned
  volatile int* mirror_host;          // pinned + mapped (max_batch, 2): progress / done written by the CG kernels
  int* mirror_dev;                    // device alias of mirror_host
  int* status;                        // device error counter (singular small matrices)
  float* cgv[6];                      // CG vectors b,x,r,p,q,t  (max_batch, d)
  double* strip_dots;                 // (max_batch, C*32, FH_MAXM) per-strip partial dots of the fused CG
  int sms;                            // multiprocessors of the device (grid sizing)
  unsigned* grid_counter;             // (max_batch) arrival counters of the resident CG kernel's grid barrier
  uint4* res_ll;                      // exchange slots of the resident CG kernel (fh_cg_resident.cuh): T, T', dots, p.q, r.r
  size_t res_ll_bytes;
  unsigned res_seq;                   // sequence numbers handed out so far (every flag in res_ll is <= res_seq)
  int cg_mode;                        // FH_CG_AUTO / FH_CG_STREAMING / FH_CG_RESIDENT (fh_ctx_set_cg_mode)
  int cg_last_mode;                   // path taken by the last DCT-basis fh_cg_solve
  int res_ok;                         // the resident kernel's cooperative grid fits this device
  // Device buffers of destroyed states, kept for the next fh_state_create of the same batch: a sampler builds one state
  // per image, and cudaMalloc / cudaFree of its ~60 MB synchronise the device (measured: sporadic 10-100 ms stalls).
  struct fh_state* state_pool[FH_STATE_POOL];
  long long* trace;                   // FH_TRACE=1: clock64 checkpoints of the fused CG kernels (8 kernels x 16 slots)
  // float64 resident CG solve (fh_cg_res64.cuh)
  double2 *tw64, *e512;               // float64 stage twiddles / exp(-i pi k / 512)
  double* r64_T[2];                   // (d) exchange planes
  double* r64_ride[2];                // per-CTA partials that ride on the two grid barriers
  double* d64[3];                     // (max_batch, d) float64 vectors: b, solution, A mu / converted inputs
  double* d64_out;                    // (max_batch, d) float64 result of the float32 entry points
  double* d64_y;                      // (max_batch, d) widened y of the float32 entry points
  int r64_smem_max;                   // opt-in shared memory per block: the float64 kernel's Z capacity is what its fixed part, tables and gather staging leave
  int r64_ok;
  // fused covariance updates (fh_upd64.cuh): one cooperative launch per update and image
  double *upd_gpart, *upd_K, *upd_dpart;
  int upd_ok, upd_cap, upd_smem;      // the grid (strips + 2 core CTAs) fits; rows of Z resident per CTA; dynamic shared memory
};

struct fh_state {
  fh_ctx* ctx;
  int B, basis, m;
  long d;
  double var2;                        // sigma^2 of the Hessian side: drifts by the reference's float32-rounded shifts
  double sig2;                        // sigma^2 of the Sigma side (exact: the reference's Sigma^-1 shifts are ~1e-7 accurate)
  float *pi, *delta, *Z;              // (B,d), (B,d), (B,FH_MAXM,d)
  double *Gz, *Gs;                    // (B,FH_MAXM,FH_MAXM)
  int* hint;                          // (B) iterations of each image's previous float64 CG solve on this state, 0 after init: the next
                                      // solve reads it to choose its arithmetic path (Zt mode of fh_k_cg_res64) -- a function of the
                                      // state's own history, so a batch and its images solved one by one take the same path
};

struct fh_op {
  fh_ctx* ctx;
  int kind, sf;
  const float* mask; int mask_batch;
  bool separable;
  int* t_dy; int* t_dx; float* t_w; int nt, ymin, ymax, xmin, xmax;
  float *kr, *kc; int nr, nc, cr, cc;
  float* psf_dense; int ks;           // dense PSF on the device (fast x4 super-resolution kernels)
  int4* runs; float *run_wf, *run_wt; int nruns, nrun_taps;   // the PSF as one dense vertical run per column (fh_k_conv_runs)
  int4* runs6; double *run_wf6, *run_wt6; int nrun_taps6;     // the same runs with float64 weights, every run padded to a multiple of 6 (fh_r64_gather_conv)
};

static void state_free_buffers(fh_state* p);

static inline dim3 ew_grid(long d, int B) { return dim3((unsigned)fh_nchunks(d), (unsigned)B, 1); }
static inline int vec_ok(long d) { return (d % 4 == 0) ? 1 : 0; }

// back-off of the host poll on mapped memory (streaming CG): a pause instruction per probe, the core is yielded after
// ~1000 probes (a solve iteration is 100-250 us at the batch sizes that take this path)
static inline void fh_host_relax(long spins) {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
  if (spins > 1000 && (spins & 63) == 0) std::this_thread::yield();
}
extern "C" const char* fh_last_error(void) { return g_err.c_str(); }
extern "C" int fh_version(void) { return 100; }
extern "C" long long fh_launch_count(int reset) {
  long long v = fh_g_launches;
  if (reset) fh_g_launches = 0;
  return v;
}

// --------------------------------------------------------------------------------------------- ctx
extern "C" int fh_ctx_create(fh_ctx** out, int device, int max_batch, int channels, int height, int width, int max_rank) {
  if (!out || max_batch < 1 || channels < 1 || height < 1 || width < 1) FH_FAIL(FH_EINVAL, "bad ctx arguments");
  // Z is allocated with max_rank rows per image and indexed with stride FH_MAXM by every kernel: they must agree
  if (max_rank != FH_MAXM) FH_FAIL(FH_EINVAL, "max_rank must be %d in this build (got %d)", FH_MAXM, max_rank);
  FH_CUDA(cudaSetDevice(device));
  fh_ctx* c = (fh_ctx*)calloc(1, sizeof(fh_ctx));
  c->device = device; c->max_batch = max_batch; c->C = channels; c->H = height; c->W = width; c->max_rank = max_rank;
  c->d = (long)channels * height * width;
  std::vector<float2> tw(3 * 3 * 64), post(FH_N), cw(FH_N);
  for (int s3 = 0; s3 < 3; ++s3) {                       // Stockham stages Ns = 4, 16, 64 (fh_dct.cuh)
    const int Ns = 4 << (2 * s3);
    for (int r = 1; r <= 3; ++r)
      for (int j = 0; j < 64; ++j) {
        const double a = -2.0 * M_PI * (double)(j % Ns) * r / (4.0 * Ns);
        tw[(s3 * 3 + (r - 1)) * 64 + j] = make_float2((float)cos(a), (float)sin(a));
      }
  }
  for (int j = 0; j < FH_N; ++j) {
    const double th = -M_PI * j / (2.0 * FH_N);
    const double f = (j == 0) ? sqrt(1.0 / FH_N) : sqrt(2.0 / FH_N);
    post[j] = make_float2((float)(f * cos(th)), (float)(f * sin(th)));
    cw[j] = make_float2((float)cos(-th), (float)sin(-th));
  }
  FH_CUDA(cudaMalloc((void**)&c->tw256, tw.size() * sizeof(float2)));
  FH_CUDA(cudaMalloc((void**)&c->post, FH_N * sizeof(float2)));
  FH_CUDA(cudaMalloc((void**)&c->cw, FH_N * sizeof(float2)));
  FH_CUDA(cudaMemcpy(c->tw256, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  FH_CUDA(cudaMemcpy(c->post, post.data(), FH_N * sizeof(float2), cudaMemcpyHostToDevice));
  FH_CUDA(cudaMemcpy(c->cw, cw.data(), FH_N * sizeof(float2), cudaMemcpyHostToDevice));
  const size_t vec = (size_t)max_batch * c->d * sizeof(float);
  FH_CUDA(cudaMalloc((void**)&c->plane_tmp, vec));
  for (int i = 0; i < 7; ++i) FH_CUDA(cudaMalloc((void**)&c->scratch[i], vec));
  FH_CUDA(cudaMalloc((void**)&c->Gtmp, (size_t)max_batch * FH_MAXM * FH_MAXM * sizeof(double)));
  for (int i = 0; i < 6; ++i) FH_CUDA(cudaMalloc((void**)&c->cgv[i], vec));
  const int nch = fh_nchunks(c->d);
  FH_CUDA(cudaMalloc((void**)&c->dot_partial, (size_t)max_batch * nch * (FH_MAXM + 2) * sizeof(double)));
  FH_CUDA(cudaMalloc((void**)&c->csum, (size_t)max_batch * (FH_MAXM + 2) * sizeof(double)));
  FH_CUDA(cudaMalloc((void**)&c->coef, (size_t)max_batch * FH_MAXM * sizeof(double)));
  FH_CUDA(cudaMalloc((void**)&c->gram_partial, (size_t)max_batch * FH_GRAM_CTAS * FH_GRAM_LD * FH_GRAM_LD * sizeof(double)));
  FH_CUDA(cudaMalloc((void**)&c->gramK, (size_t)max_batch * FH_GRAM_LD * FH_GRAM_LD * sizeof(double)));
  const size_t npart = std::max((size_t)nch, (size_t)channels * (FH_N / 2));
  FH_CUDA(cudaMalloc((void**)&c->part_a, (size_t)max_batch * npart * sizeof(double)));
  FH_CUDA(cudaMalloc((void**)&c->part_b, (size_t)max_batch * npart * sizeof(double)));
  FH_CUDA(cudaMalloc((void**)&c->cg_sc, (size_t)max_batch * sizeof(FhCgScalars)));
  FH_CUDA(cudaMallocHost((void**)&c->cg_sc_host, (size_t)max_batch * sizeof(FhCgScalars)));
  FH_CUDA(cudaHostAlloc((void**)&c->mirror_host, (size_t)max_batch * 2 * sizeof(int), cudaHostAllocMapped));
  FH_CUDA(cudaHostGetDevicePointer((void**)&c->mirror_dev, (void*)c->mirror_host, 0));
  FH_CUDA(cudaMalloc((void**)&c->status, sizeof(int)));
  FH_CUDA(cudaMemset(c->status, 0, sizeof(int)));
  FH_CUDA(cudaMalloc((void**)&c->strip_dots, (size_t)max_batch * channels * (FH_N / 2) * FH_MAXM * sizeof(double)));
  c->sms = 148;
  FH_CUDA(cudaMalloc((void**)&c->grid_counter, (size_t)max_batch * sizeof(unsigned)));
  {
    const size_t nblk = (size_t)(c->d / FH_RES_SLICE) + 1;
    c->res_ll_bytes = ((size_t)c->d + FH_MAXM * nblk + (FH_MAXM + 2 * nblk) * FH_LL_SPREAD) * sizeof(uint4);   // 2 x d/2, dots, sums, a, b
    FH_CUDA(cudaMalloc((void**)&c->res_ll, c->res_ll_bytes));
    FH_CUDA(cudaMemset(c->res_ll, 0, c->res_ll_bytes));
    c->res_seq = 0;
  }
  c->cg_mode = FH_CG_AUTO; c->cg_last_mode = 0; c->res_ok = 1;
  {
    // float64 tables of the resident float64 solve: stage twiddles [Ns = 4, 16, 64][r = 1..3][k < Ns] and exp(-i pi k / 512)
    std::vector<double2> tw64(FH_R64_TW), e512(FH_N);
    int off = 0;
    for (int Ns = 4; Ns <= 64; Ns *= 4) {
      for (int r = 1; r <= 3; ++r)
        for (int k = 0; k < Ns; ++k) {
          const double a = -2.0 * M_PI * (double)k * r / (4.0 * Ns);
          tw64[off + (r - 1) * Ns + k] = make_double2(cos(a), sin(a));
        }
      off += 3 * Ns;
    }
    for (int k = 0; k < FH_N; ++k) { const double th = -M_PI * k / (2.0 * FH_N); e512[k] = make_double2(cos(th), sin(th)); }
    FH_CUDA(cudaMalloc((void**)&c->tw64, tw64.size() * sizeof(double2)));
    FH_CUDA(cudaMalloc((void**)&c->e512, e512.size() * sizeof(double2)));
    FH_CUDA(cudaMemcpy(c->tw64, tw64.data(), tw64.size() * sizeof(double2), cudaMemcpyHostToDevice));
    FH_CUDA(cudaMemcpy(c->e512, e512.data(), e512.size() * sizeof(double2), cudaMemcpyHostToDevice));
    for (int i = 0; i < 2; ++i) {
      FH_CUDA(cudaMalloc((void**)&c->r64_T[i], (size_t)c->d * sizeof(double)));
      FH_CUDA(cudaMalloc((void**)&c->r64_ride[i], ((size_t)c->d / FH_RES_SLICE + 1) * sizeof(double)));
    }
    for (int i = 0; i < 3; ++i) FH_CUDA(cudaMalloc((void**)&c->d64[i], (size_t)max_batch * c->d * sizeof(double)));
    FH_CUDA(cudaMalloc((void**)&c->d64_out, (size_t)max_batch * c->d * sizeof(double)));
    FH_CUDA(cudaMalloc((void**)&c->d64_y, (size_t)max_batch * c->d * sizeof(double)));
    c->r64_smem_max = 227 * 1024; c->r64_ok = c->res_ok;
    const size_t nstrip = (size_t)channels * FH_R64_CPP;
    FH_CUDA(cudaMalloc((void**)&c->upd_gpart, 2 * nstrip * FH_U64_CP * FH_U64_NBLK2 * 64 * sizeof(double)));
    FH_CUDA(cudaMalloc((void**)&c->upd_K, 2 * (size_t)FH_U64_LD * FH_U64_LD * sizeof(double)));
    FH_CUDA(cudaMalloc((void**)&c->upd_dpart, nstrip * FH_U64_DP * sizeof(double)));
    FH_CUDA(cudaMemset(c->upd_dpart, 0, nstrip * FH_U64_DP * sizeof(double)));
    c->upd_ok = 0; c->upd_cap = 0; c->upd_smem = 0;
  }
  c->trace = nullptr;
  if (getenv("FH_TRACE")) { FH_CUDA(cudaMalloc((void**)&c->trace, 8 * 16 * sizeof(long long))); FH_CUDA(cudaMemset(c->trace, 0, 8 * 16 * sizeof(long long))); }
#ifndef FH_EMU
  { cudaDeviceProp pr; FH_CUDA(cudaGetDeviceProperties(&pr, device)); c->sms = pr.multiProcessorCount; }
#endif
#ifndef FH_EMU
  const int core_smem = (2 * FH_MAXM * FH_MAXM + 3 * FH_MAXM) * (int)sizeof(double);
  FH_CUDA(cudaFuncSetAttribute(fh_k_core, cudaFuncAttributeMaxDynamicSharedMemorySize, core_smem));
  FH_CUDA(cudaFuncSetAttribute(fh_k_core_from_cov, cudaFuncAttributeMaxDynamicSharedMemorySize, core_smem));
  // 5 rounds x 32 KB of Z in flight + ~42 KB static: the most one CTA can hold
  FH_CUDA(cudaFuncSetAttribute(fh_k_cg_fwd_cols<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_ZMAX_STAGES * FH_ZROUND_FLOATS * 4));
  FH_CUDA(cudaFuncSetAttribute(fh_k_cg_fwd_cols<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_ZMAX_STAGES * FH_ZROUND_FLOATS * 4));
  FH_CUDA(cudaFuncSetAttribute(fh_k_cg_fwd_cols<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_ZMAX_STAGES * FH_ZROUND_FLOATS * 4));
  FH_CUDA(cudaFuncSetAttribute(fh_k_cg_inv_cols<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_ZMAX_STAGES * FH_ZROUND_FLOATS * 4));
  FH_CUDA(cudaFuncSetAttribute(fh_k_cg_inv_cols<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_ZMAX_STAGES * FH_ZROUND_FLOATS * 4));
  FH_CUDA(cudaFuncSetAttribute(fh_k_cg_inv_cols<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_ZMAX_STAGES * FH_ZROUND_FLOATS * 4));
  {
    // resident CG solve: one CTA per SM holding up to FH_RES_MAXM row slices of Z; needs C*256/6 co-resident CTAs
    const int res_smem = (int)sizeof(FhResSmem) + FH_RES_MAXM * FH_RES_SLICE * 4;
    FH_CUDA(cudaFuncSetAttribute(fh_k_cg_resident<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, res_smem));
    FH_CUDA(cudaFuncSetAttribute(fh_k_cg_resident<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, res_smem));
    int per_sm = 0, per_sm_big = 0;
    FH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fh_k_cg_resident<false>, FH_PASS_THREADS, res_smem));
    FH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_big, fh_k_cg_resident<true>, FH_PASS_THREADS, res_smem));
    per_sm = std::min(per_sm, per_sm_big);
    int coop = 0;
    FH_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    const long rows = (long)channels * height;
    c->res_ok = (coop && height == FH_N && width == FH_N && rows % FH_RES_ROWS == 0 && rows / FH_RES_ROWS <= (long)per_sm * c->sms) ? 1 : 0;
    // float64 variant: as many rows of Z as fit beside its fixed shared memory
    int smem_max = 0;
    FH_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    c->r64_smem_max = smem_max;
    c->r64_ok = 0;
    if (c->res_ok && smem_max >= (int)(sizeof(FhRes64Smem) + sizeof(FhRes64Tabs))) {
      FH_CUDA(cudaFuncSetAttribute(fh_k_cg_res64<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      FH_CUDA(cudaFuncSetAttribute(fh_k_cg_res64<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      FH_CUDA(cudaFuncSetAttribute(fh_k_cg_res64<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      int a = 0, b = 0, e = 0;
      FH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, fh_k_cg_res64<false, true>, FH_PASS_THREADS, smem_max));
      FH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fh_k_cg_res64<false, false>, FH_PASS_THREADS, smem_max));
      FH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e, fh_k_cg_res64<true, false>, FH_PASS_THREADS, smem_max));
      c->r64_ok = ((long)channels * FH_R64_CPP <= (long)std::min(std::min(a, b), e) * c->sms) ? 1 : 0;
    }
    if (c->r64_ok) {
      // fused updates: strips + 2 core CTAs, one CTA per SM; the kernel's static shared memory (Gauss-Jordan staging) counts
      cudaFuncAttributes fa;
      FH_CUDA(cudaFuncGetAttributes(&fa, fh_k_upd64));
      const int room = smem_max - (int)fa.sharedSizeBytes;
      FH_CUDA(cudaFuncSetAttribute(fh_k_upd64, cudaFuncAttributeMaxDynamicSharedMemorySize, room));
      c->upd_cap = std::max(0, std::min(FH_RES_MAXM, (room - (int)sizeof(FhRes64Smem)) / (FH_RES_SLICE * (int)sizeof(float))));
      c->upd_smem = room;
      int per = 0;
      FH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fh_k_upd64, FH_PASS_THREADS, room));
      c->upd_ok = (room >= (int)sizeof(FhUpd64Core) && room >= (int)sizeof(FhRes64Smem) && (long)channels * FH_R64_CPP + 2 <= (long)per * c->sms) ? 1 : 0;
    }
  }
#else
  c->res_ok = (height == FH_N && width == FH_N && ((long)channels * height) % FH_RES_ROWS == 0) ? 1 : 0;
  c->r64_ok = c->res_ok;
  if (const char* e = getenv("FH_EMU_R64_SMEM")) c->r64_smem_max = atoi(e);      // CPU emulation: the test shrinks the shared memory to reach the L2 / L1 variants at small rank
  c->upd_ok = c->r64_ok;
  c->upd_smem = std::max(c->r64_smem_max, (int)sizeof(FhUpd64Core));
  c->upd_cap = std::max(0, std::min(FH_RES_MAXM, (c->r64_smem_max - (int)sizeof(FhRes64Smem)) / (FH_RES_SLICE * (int)sizeof(float))));
#endif
  *out = c;
  return FH_OK;
}

extern "C" int fh_ctx_set_cg_mode(fh_ctx* c, int mode) {
  if (!c || mode < FH_CG_AUTO || mode > FH_CG_RESIDENT_F32) FH_FAIL(FH_EINVAL, "bad fh_ctx_set_cg_mode arguments");
  c->cg_mode = mode;
  return FH_OK;
}
extern "C" int fh_ctx_cg_last_mode(const fh_ctx* c) { return c ? c->cg_last_mode : FH_EINVAL; }
extern "C" int fh_ctx_fused_updates(const fh_ctx* c) { return c ? c->upd_ok : FH_EINVAL; }

// Debug: copies the 8 x 16 clock64 checkpoints recorded by CTA 0 of the fused CG kernels (FH_TRACE=1) to the host.
extern "C" int fh_trace_read(fh_ctx* c, long long* out_host) {
  if (!c || !c->trace || !out_host) FH_FAIL(FH_EINVAL, "tracing is off (set FH_TRACE=1 before fh_ctx_create)");
  FH_CUDA(cudaDeviceSynchronize());
  FH_CUDA(cudaMemcpy(out_host, c->trace, 8 * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return FH_OK;
}

extern "C" int fh_ctx_destroy(fh_ctx* c) {
  if (!c) return FH_OK;
  for (int i = 0; i < FH_STATE_POOL; ++i)
    if (c->state_pool[i]) state_free_buffers(c->state_pool[i]);
  cudaFree(c->tw256); cudaFree(c->post); cudaFree(c->cw); cudaFree(c->plane_tmp);
  for (int i = 0; i < 7; ++i) cudaFree(c->scratch[i]);
  for (int i = 0; i < 6; ++i) cudaFree(c->cgv[i]);
  cudaFree(c->Gtmp); cudaFree(c->strip_dots); cudaFree(c->trace); cudaFree(c->res_ll); cudaFree(c->grid_counter);
  cudaFree(c->upd_gpart); cudaFree(c->upd_K); cudaFree(c->upd_dpart);
  cudaFree(c->dot_partial); cudaFree(c->csum); cudaFree(c->coef); cudaFree(c->gram_partial); cudaFree(c->gramK);
  cudaFree(c->tw64); cudaFree(c->e512);
  for (int i = 0; i < 2; ++i) { cudaFree(c->r64_T[i]); cudaFree(c->r64_ride[i]); }
  for (int i = 0; i < 3; ++i) cudaFree(c->d64[i]);
  cudaFree(c->d64_out); cudaFree(c->d64_y);
  cudaFree(c->part_a); cudaFree(c->part_b); cudaFree(c->cg_sc); cudaFreeHost(c->cg_sc_host); cudaFreeHost((void*)c->mirror_host); cudaFree(c->status);
  free(c);
  return FH_OK;
}

// --------------------------------------------------------------------------------------------- DCT
// Strip height by problem size: short strips for small batches so that every SM gets work.
static inline int pick_rows(const fh_ctx* c, int nplanes) {
#ifdef FH_EMU
  if (const char* e = getenv("FH_EMU_ROWS")) return atoi(e);      // CPU emulation: the test picks the instantiation
#endif
  // 8-row strips everywhere: the consumer-side reductions of the CG kernels read one partial per strip, and the
  // shorter instantiations (more, smaller CTAs) measured slower on B200 even at batch 1 (profiles/).
  (void)c; (void)nplanes;
  return 8;
}
static inline int pick_zstages(const fh_ctx* c, int nstrips) {
  if (const char* e = getenv("FH_ZSTAGES")) { int v = atoi(e); if (v >= 1 && v <= FH_ZMAX_STAGES) return v; }
  // one CTA per SM with a deep Z pipeline (4 x 32 KB leaves room for the next kernel's prologue to co-reside),
  // or two CTAs per SM
  return nstrips <= c->sms ? 4 : 2;
}
#define FH_ROWS_SWITCH(R, ...)                                    \
  switch (R) {                                                    \
    case 8: { constexpr int RR = 8; __VA_ARGS__; } break;         \
    case 4: { constexpr int RR = 4; __VA_ARGS__; } break;         \
    default: { constexpr int RR = 2; __VA_ARGS__; } break;        \
  }
// grid for a strip-loop kernel: enough CTAs to fill the machine, never more than there are strips
static inline dim3 strip_grid(const fh_ctx* c, int nstrips) { return dim3((unsigned)std::min(nstrips, c->sms * 6)); }

template <bool INV, bool TR>
static void launch_pass(fh_ctx* c, const float* in, float* out, int nplanes, void* stream) {
  FhDctTables tab{c->tw256, c->post, c->cw};
  const int R = pick_rows(c, nplanes);
  const int nstrips = nplanes * (FH_N / R);
  const dim3 grid = strip_grid(c, nstrips), block(FH_PASS_THREADS);
  FH_ROWS_SWITCH(R, {
    constexpr auto fh_k_pass_ = fh_k_pass<INV, TR, RR>;
    FH_LAUNCH(fh_k_pass_, grid, block, 0, stream, in, out, nstrips, tab);
  });
}

// second_transposed = 1: natural layout out (public fh_dct2); 0: the library's basis layout (see fh_dct.cuh)
static int dct2_planes(fh_ctx* c, const float* in, float* out, int nplanes, int inverse, int second_transposed, void* stream) {
  if (c->H != FH_N || c->W != FH_N) FH_FAIL(FH_EINVAL, "DCT basis needs 256x256 planes (got %dx%d)", c->H, c->W);
  if ((size_t)nplanes * FH_N * FH_N > (size_t)c->max_batch * c->d) FH_FAIL(FH_EINVAL, "too many planes for the scratch buffer");
  if (inverse) {
    launch_pass<true, true>(c, in, c->plane_tmp, nplanes, stream);
    if (second_transposed) launch_pass<true, true>(c, c->plane_tmp, out, nplanes, stream);
    else launch_pass<true, false>(c, c->plane_tmp, out, nplanes, stream);
  } else {
    launch_pass<false, true>(c, in, c->plane_tmp, nplanes, stream);
    if (second_transposed) launch_pass<false, true>(c, c->plane_tmp, out, nplanes, stream);
    else launch_pass<false, false>(c, c->plane_tmp, out, nplanes, stream);
  }
  FH_KCHECK();
  return FH_OK;
}
// natural <-> basis layout of DCT-domain vectors (import/export only)
static int transpose_planes(fh_ctx* c, const float* in, float* out, int nplanes, void* stream) {
  FH_LAUNCH(fh_k_transpose, dim3(8, 8, nplanes), dim3(32, 8), 0, stream, in, out);
  FH_KCHECK();
  return FH_OK;
}
extern "C" int fh_dct2(fh_ctx* c, const float* in, float* out, int nplanes, int inverse, void* stream) {
  if (!c || !in || !out || nplanes < 1) FH_FAIL(FH_EINVAL, "bad fh_dct2 arguments");
  return dct2_planes(c, in, out, nplanes, inverse, 1, stream);
}

// ------------------------------------------------------------------------------------------- state
static void state_free_buffers(fh_state* p) {
  cudaFree(p->pi); cudaFree(p->delta); cudaFree(p->Z); cudaFree(p->Gz); cudaFree(p->Gs); cudaFree(p->hint);
  free(p);
}
// A pool slot for a state of `batch` images: a free one, else one parked for a different batch size (evicted).
static int pool_slot_for(fh_ctx* c, int batch) {
  for (int i = 0; i < FH_STATE_POOL; ++i) if (!c->state_pool[i]) return i;
  for (int i = 0; i < FH_STATE_POOL; ++i)
    if (c->state_pool[i]->B != batch) { state_free_buffers(c->state_pool[i]); c->state_pool[i] = nullptr; return i; }
  return -1;
}

static int state_alloc(fh_ctx* c, fh_state* s, int batch) {
  const size_t vec = (size_t)batch * c->d * sizeof(float);
  FH_CUDA(cudaMalloc((void**)&s->pi, vec));
  FH_CUDA(cudaMalloc((void**)&s->delta, vec));
  FH_CUDA(cudaMalloc((void**)&s->Z, (size_t)batch * c->max_rank * c->d * sizeof(float) + 16));
  FH_CUDA(cudaMalloc((void**)&s->Gz, (size_t)batch * FH_MAXM * FH_MAXM * sizeof(double)));
  FH_CUDA(cudaMalloc((void**)&s->Gs, (size_t)batch * FH_MAXM * FH_MAXM * sizeof(double)));
  FH_CUDA(cudaMalloc((void**)&s->hint, (size_t)batch * sizeof(int)));
  FH_CUDA(cudaMemset(s->hint, 0, (size_t)batch * sizeof(int)));
  return FH_OK;
}

extern "C" int fh_state_create(fh_ctx* c, fh_state** out, int batch, int basis) {
  if (!c || !out || batch < 1 || batch > c->max_batch) FH_FAIL(FH_EINVAL, "bad fh_state_create arguments");
  if (basis == FH_BASIS_DCT && (c->H != FH_N || c->W != FH_N)) FH_FAIL(FH_EINVAL, "DCT basis needs 256x256 images");
  fh_state* s = (fh_state*)calloc(1, sizeof(fh_state));
  s->ctx = c; s->B = batch; s->basis = basis; s->m = 0; s->d = c->d; s->var2 = 1.0; s->sig2 = 1.0;
  for (int i = 0; i < FH_STATE_POOL; ++i) {
    fh_state* p = c->state_pool[i];
    if (p && p->B == batch) {                               // reuse the buffers of a destroyed state (no cudaMalloc)
      c->state_pool[i] = nullptr;
      s->pi = p->pi; s->delta = p->delta; s->Z = p->Z; s->Gz = p->Gz; s->Gs = p->Gs; s->hint = p->hint;
      free(p);
      *out = s;
      return FH_OK;
    }
  }
  int rc = state_alloc(c, s, batch);
  if (rc) { state_free_buffers(s); cudaGetLastError(); return rc; }      // partial allocation: nothing leaks
  // A sampler typically builds the next image's state while it still holds the previous one (or two: a caller that keeps
  // the last mechanism for its log while the next image is sampled): allocate spare sets of buffers together with a state
  // that missed the pool, so that the following images do not pay cudaMalloc inside their loop -- ~2 ms per state as a
  // rule, but 100 - 340 ms every now and then on a busy box (measured: bench.py's first end-to-end step).
  for (int k = 0; k < FH_STATE_SPARES; ++k) {
    int i = -1;
    for (int j = 0; j < FH_STATE_POOL; ++j) if (!c->state_pool[j]) { i = j; break; }
    if (i < 0) break;
    fh_state* spare = (fh_state*)calloc(1, sizeof(fh_state));
    spare->ctx = c; spare->B = batch; spare->basis = basis; spare->d = c->d;
    if (state_alloc(c, spare, batch) == FH_OK) c->state_pool[i] = spare;
    else { state_free_buffers(spare); cudaGetLastError(); break; }
  }
  *out = s;
  return FH_OK;
}
extern "C" int fh_state_destroy(fh_state* s) {
  if (!s) return FH_OK;
  if (s->ctx) {
    const int i = pool_slot_for(s->ctx, s->B);              // park the buffers; work queued on them is stream-ordered
    if (i >= 0) { s->ctx->state_pool[i] = s; return FH_OK; } // before whatever the next owner queues
  }
  state_free_buffers(s);
  return FH_OK;
}
extern "C" int fh_state_rank(const fh_state* s) { return s ? s->m : -1; }
extern "C" double fh_state_var2(const fh_state* s) { return s ? s->sig2 : -1.0; }

// Z is allocated with row capacity ctx->max_rank but indexed with stride FH_MAXM in the kernels: keep them equal.
static int check_rank_layout(fh_state* s) {
  if (s->ctx->max_rank != FH_MAXM) FH_FAIL(FH_EINVAL, "ctx max_rank must equal %d in this build", FH_MAXM);
  return FH_OK;
}

static int refresh_delta(fh_state* s, void* stream) {
  FH_LAUNCH(fh_k_delta, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->pi, s->delta, (float)(1.0 / s->sig2), s->d, vec_ok(s->d));
  FH_KCHECK();
  return FH_OK;
}

// K = Z diag(w) Z^T (+ extra rows) into ctx->gramK
static int gram(fh_state* s, int m, FhRowPtrs xr, int nx, const float* w, void* stream) {
  fh_ctx* c = s->ctx;
  const size_t smem = (size_t)(FH_GRAM_MAXR * (FH_GRAM_TILE + 1) + FH_GRAM_TILE) * sizeof(float);
  FH_LAUNCH(fh_k_gram, dim3(FH_GRAM_CTAS, s->B), dim3(256), smem, stream, s->Z, m, xr, nx, w, c->gram_partial, s->d);
  FH_KCHECK();
  FH_LAUNCH(fh_k_gram_reduce, dim3(s->B, FH_GRAM_RED_CTAS), dim3(256), 0, stream, c->gram_partial, FH_GRAM_CTAS, m + nx, c->gramK);
  FH_KCHECK();
  return FH_OK;
}

// Core of Sigma at diagonal `delta` (= 1/(pi + 1/sigma^2)):  Gs_out = -Gz (I + Z delta Z^T Gz)^-1.  If tu_s != null
// (score in the state's basis) also leaves the time-update coefficient vector in ctx->coef.
static int core_at(fh_state* s, const float* delta, double* Gs_out, const float* tu_s, double var2_old, void* stream) {
  fh_ctx* c = s->ctx;
  if (s->m == 0) return FH_OK;
  FhRowPtrs xr{};
  int nx = 0;
  if (tu_s) {
    FH_LAUNCH(fh_k_tu_rows, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, tu_s, s->pi, delta, c->scratch[4], c->scratch[5], s->d, vec_ok(s->d));
    FH_KCHECK();
    xr.p[0] = tu_s; xr.p[1] = c->scratch[4]; xr.p[2] = c->scratch[5];
    nx = 3;
  }
  int rc = gram(s, s->m, xr, nx, delta, stream);
  if (rc) return rc;
  const size_t smem = (size_t)(2 * FH_MAXM * FH_MAXM + 3 * FH_MAXM) * sizeof(double);
  FH_LAUNCH(fh_k_core, dim3(s->B), dim3(FH_CORE_THREADS), smem, stream, c->gramK, s->Gz, s->m, Gs_out, tu_s ? 1 : 0, var2_old, c->coef, c->status);
  FH_KCHECK();
  return FH_OK;
}

// Sigma side of the state at s->sig2: delta and Gs.
static int refresh_core(fh_state* s, void* stream) {
  int rc = refresh_delta(s, stream);
  if (rc) return rc;
  return core_at(s, s->delta, s->Gs, nullptr, 0.0, stream);
}

extern "C" int fh_state_init(fh_state* s, const float* var_dev, double init_noise_variance, void* stream) {
  if (!s || !var_dev || !(init_noise_variance > 0)) FH_FAIL(FH_EINVAL, "bad fh_state_init arguments");
  int rc = check_rank_layout(s);
  if (rc) return rc;
  s->m = 0;
  s->var2 = init_noise_variance;
  s->sig2 = init_noise_variance;
  if (s->basis == FH_BASIS_DCT) {              // init_variance is given in the natural (c, ki, kj) order
    if ((rc = transpose_planes(s->ctx, var_dev, s->ctx->scratch[0], s->ctx->C, stream))) return rc;
    var_dev = s->ctx->scratch[0];
  }
  FH_LAUNCH(fh_k_init_pi, dim3((unsigned)((s->d + 255) / 256), s->B), dim3(256), 0, stream, var_dev, s->pi, 1.0 / init_noise_variance, s->d);
  FH_KCHECK();
  FH_CUDA(cudaMemsetAsync(s->Gz, 0, (size_t)s->B * FH_MAXM * FH_MAXM * sizeof(double), (cudaStream_t)stream));
  FH_CUDA(cudaMemsetAsync(s->Gs, 0, (size_t)s->B * FH_MAXM * FH_MAXM * sizeof(double), (cudaStream_t)stream));
  FH_CUDA(cudaMemsetAsync(s->hint, 0, (size_t)s->B * sizeof(int), (cudaStream_t)stream));
  return refresh_delta(s, stream);
}

// to / from the state's basis
static int to_basis(fh_state* s, const float* in, float* out, void* stream) {
  if (s->basis == FH_BASIS_DCT) return dct2_planes(s->ctx, in, out, s->B * s->ctx->C, 0, 0, stream);
  FH_CUDA(cudaMemcpyAsync(out, in, (size_t)s->B * s->d * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return FH_OK;
}
static int from_basis(fh_state* s, const float* in, float* out, void* stream) {
  if (s->basis == FH_BASIS_DCT) return dct2_planes(s->ctx, in, out, s->B * s->ctx->C, 1, 0, stream);
  FH_CUDA(cudaMemcpyAsync(out, in, (size_t)s->B * s->d * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return FH_OK;
}

// out = Sigma v with v, out in the state's basis (out may alias v)
static int cov_dot_basis(fh_state* s, const float* v, float* out, void* stream) {
  fh_ctx* c = s->ctx;
  const int vec = vec_ok(s->d);
  if (s->m > 0) {
    FH_LAUNCH(fh_k_dots, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->Z, s->m, v, s->delta, (const float*)nullptr, (const float*)nullptr, c->dot_partial, s->d, vec);
    FH_KCHECK();
    FH_LAUNCH(fh_k_dots_finalize, dim3(s->B), dim3(128), 0, stream, c->dot_partial, fh_nchunks(s->d), s->m, 0, s->Gs, c->csum, c->coef);
    FH_KCHECK();
  }
  // out = delta .* (v + Z^T coef)
  FH_LAUNCH(fh_k_lowrank_apply, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->Z, s->m, c->coef, s->delta, v, 1.0f, out, s->d, vec);
  FH_KCHECK();
  return FH_OK;
}

static FhCgFused fused_params(fh_state* s, fh_op* o, int R) {
  fh_ctx* c = s->ctx;
  FhCgFused P{};
  P.B = s->B; P.C = c->C; P.m = s->m;
  P.fuse = 0;
  P.strips_per_image = c->C * (FH_N / R);
  P.ystrips = P.strips_per_image;
  P.zstages = pick_zstages(c, s->B * P.strips_per_image);
  P.last_allowed = 1 << 30;
  if (o && o->kind == FH_OP_INPAINT) { P.fuse = 1; P.mask = o->mask; P.mask_batch = o->mask_batch; }
  if (o && o->kind == FH_OP_BLUR && o->separable && o->nr <= FH_MAX_SEP_TAPS && o->nc <= FH_MAX_SEP_TAPS) {
    P.fuse = 2; P.kr = o->kr; P.kc = o->kc; P.nr = o->nr; P.nc = o->nc; P.cr = o->cr; P.cc = o->cc;
  }
  P.Z = s->Z; P.delta = s->delta; P.Gs = s->Gs;
  P.x = c->cgv[1]; P.r = c->cgv[2]; P.p = c->cgv[3]; P.q = c->cgv[4];
  P.T = c->plane_tmp; P.W = c->scratch[2];
  P.dots_partial = c->strip_dots; P.part_a = c->part_a; P.part_b = c->part_b;
  P.sc = nullptr; P.sy2 = 0.f; P.trace = c->trace; P.mirror = nullptr;
  P.tab = FhDctTables{c->tw256, c->post, c->cw};
  return P;
}

// out = Sigma in, both in pixel space, DCT basis: four fused launches (no operator folded in)
static int sigma_apply_pixels(fh_state* s, const float* in, float* out, void* stream) {
  const int R = pick_rows(s->ctx, s->B * s->ctx->C);
  FhCgFused P = fused_params(s, nullptr, R);
  const dim3 grid = strip_grid(s->ctx, P.B * P.strips_per_image), block(FH_PASS_THREADS);
  const size_t zbytes = (size_t)P.zstages * FH_ZROUND_FLOATS * 4;
  FH_ROWS_SWITCH(R, {
    FH_LAUNCH(fh_k_cg_fwd_rows<RR>, grid, block, 0, stream, P, in, (const float*)nullptr, 0);
    FH_LAUNCH(fh_k_cg_fwd_cols<RR>, grid, block, zbytes, stream, P);
    FH_LAUNCH(fh_k_cg_inv_cols<RR>, grid, block, zbytes, stream, P);
    FH_LAUNCH(fh_k_cg_inv_rows<RR>, grid, block, 0, stream, P, out, (const float*)nullptr, 0);
  });
  FH_KCHECK();
  return FH_OK;
}

extern "C" int fh_cov_dot(fh_state* s, const float* v, float* out, void* stream) {
  if (!s || !v || !out) FH_FAIL(FH_EINVAL, "bad fh_cov_dot arguments");
  if (s->basis == FH_BASIS_IDENTITY) return cov_dot_basis(s, v, out, stream);
  return sigma_apply_pixels(s, v, out, stream);
}

extern "C" int fh_time_update(fh_state* s, const float* x, const float* score, double sigma_t, double sigma_tnext,
                              int only_cov, float* mean_out, float* score_out, void* stream) {
  if (!s || !(sigma_t > 0) || !(sigma_tnext > 0)) FH_FAIL(FH_EINVAL, "bad fh_time_update arguments");
  if (!only_cov && (!x || !score || !mean_out || !score_out)) FH_FAIL(FH_EINVAL, "fh_time_update needs x, score and outputs");
  fh_ctx* c = s->ctx;
  const int vec = vec_ok(s->d);
  int rc;
  // Sigma side (online_update_bfgs.py:166): diag(Sigma^-1) += sigma'^-2 - sigma^-2.  The float32 rounding of that
  // shift is ~1e-7 relative at every step, so the state simply sits at the exact new sigma.
  s->sig2 = sigma_tnext * sigma_tnext;
  if (only_cov) return refresh_core(s, stream);
  // Hessian side (:172): diag(H^-1) -= float32(sigma'^2 - sigma^2), an ABSOLUTE rounding error that accumulates
  // until a space update rebuilds H from Sigma (:294).  var2 carries that drift; it only enters the score transfer.
  const double var2_old = s->var2;
  s->var2 = var2_old + (double)(float)(sigma_tnext * sigma_tnext - sigma_t * sigma_t);
  if (!(s->var2 > 0)) FH_FAIL(FH_ESTATE, "non-positive effective variance after time update");
  float* xh = c->scratch[0];
  float* sh = c->scratch[1];
  if ((rc = to_basis(s, x, xh, stream))) return rc;
  if ((rc = to_basis(s, score, sh, stream))) return rc;
  const bool consistent = fabs(s->var2 - s->sig2) <= 1e-7 * s->sig2;
  const float* dl = s->delta;
  if (consistent) {
    // one core serves both sides
    if ((rc = refresh_delta(s, stream))) return rc;
    if ((rc = core_at(s, s->delta, s->Gs, sh, var2_old, stream))) return rc;
  } else {
    float* dH = c->scratch[6];
    FH_LAUNCH(fh_k_delta, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->pi, dH, (float)(1.0 / s->var2), s->d, vec);
    FH_KCHECK();
    if ((rc = core_at(s, dH, c->Gtmp, sh, var2_old, stream))) return rc;
    dl = dH;
  }
  FH_LAUNCH(fh_k_tu_apply, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->Z, s->m, c->coef, sh, s->pi, dl, xh,
            (float)var2_old, (float)(1.0 / s->var2), (float)(sigma_tnext * sigma_tnext), c->scratch[2], c->scratch[3], s->d, vec);
  FH_KCHECK();
  if (!consistent && (rc = refresh_core(s, stream))) return rc;
  if ((rc = from_basis(s, c->scratch[2], score_out, stream))) return rc;
  return from_basis(s, c->scratch[3], mean_out, stream);
}

// ---- fused covariance updates (fh_upd64.cuh): one cooperative launch per update and image --------------------------------
// Applies to the DCT basis on a device whose SMs hold the strips plus two core CTAs; batches above 4 images keep the
// batched multi-launch path (its launches cover the whole batch at once).  FH_UPD_LEGACY=1 forces the old path (A/B runs).
static bool upd64_applies(const fh_state* s) {
  const char* e = getenv("FH_UPD_LEGACY");                   // read per call: the tests switch it inside one process
  const int legacy = e ? atoi(e) : 0;
  const fh_ctx* c = s->ctx;
  return !legacy && c->upd_ok && s->basis == FH_BASIS_DCT && s->B <= 4;
}
static int launch_upd64(fh_ctx* c, FhUpd64 U, int image, cudaStream_t st) {
  const dim3 grid((unsigned)(c->C * FH_R64_CPP + 2)), block(FH_PASS_THREADS);
  U.C = c->C; U.d = c->d; U.mr_cap = c->upd_cap;
  U.T[0] = c->r64_T[0]; U.T[1] = c->r64_T[1];
  U.gpart = c->upd_gpart; U.K = c->upd_K; U.dpart = c->upd_dpart; U.coef = c->coef + (size_t)image * FH_MAXM;
  U.status = c->status; U.tw = c->tw64; U.e512 = c->e512;
  U.grid_counter = c->grid_counter + image;
  FH_CUDA(cudaMemsetAsync(U.grid_counter, 0, sizeof(unsigned), st));
  const size_t strip = sizeof(FhRes64Smem) + (size_t)std::min(U.m, c->upd_cap) * FH_RES_SLICE * sizeof(float);
  const size_t smem = std::max(strip, sizeof(FhUpd64Core));
  if ((int)smem > c->upd_smem) FH_FAIL(FH_ESTATE, "fused update does not fit shared memory");
#ifdef FH_EMU
  emu::launch_coop([&] { fh_k_upd64(U); }, grid, block, smem);
  ++fh_g_launches;
#else
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (fh_g_prof_on) fh_prof_begin(U.mode == 1 ? "fh_k_upd64(time)" : "fh_k_upd64(space)", st);
  FH_CUDA(cudaLaunchKernelEx(&cfg, fh_k_upd64, U));
  if (fh_g_prof_on) fh_prof_end(st);
  ++fh_g_launches;
#endif
  return FH_OK;
}
static FhUpd64 upd64_state(fh_state* s, int b) {
  FhUpd64 U{};
  U.m = s->m;
  U.pi = s->pi + (size_t)b * s->d; U.delta = s->delta + (size_t)b * s->d; U.Z = s->Z + (size_t)b * FH_MAXM * s->d;
  U.Gz = s->Gz + (size_t)b * FH_MAXM * FH_MAXM; U.Gs = s->Gs + (size_t)b * FH_MAXM * FH_MAXM;
  return U;
}
// time update of every image: score = score_in, or (mean_prev - x) / sigma_t^2 when mean_prev is given
static int time_update_fused(fh_state* s, const double* x, const double* mean_prev, const double* score_in, double sigma_t, double sigma_tnext,
                             double* mean_out, double* score_out, void* stream) {
  // the bookkeeping of fh_time_update: Sigma side at the exact new sigma, Hessian side with the reference's float32-rounded shift
  s->sig2 = sigma_tnext * sigma_tnext;
  const double var2_old = s->var2;
  s->var2 = var2_old + (double)(float)(sigma_tnext * sigma_tnext - sigma_t * sigma_t);
  if (!(s->var2 > 0)) FH_FAIL(FH_ESTATE, "non-positive effective variance after time update");
  const bool consistent = fabs(s->var2 - s->sig2) <= 1e-7 * s->sig2;
  for (int b = 0; b < s->B; ++b) {
    FhUpd64 U = upd64_state(s, b);
    U.mode = 1; U.consistent = consistent ? 1 : 0;
    U.x_prev = x + (size_t)b * s->d;
    U.mean_prev = mean_prev ? mean_prev + (size_t)b * s->d : nullptr;
    U.score_in = score_in ? score_in + (size_t)b * s->d : nullptr;
    U.sig2_t = sigma_t * sigma_t; U.var2_old = var2_old; U.var2_new = s->var2; U.sig2_new = s->sig2;
    U.mean_out = mean_out + (size_t)b * s->d; U.score_out = score_out + (size_t)b * s->d;
    int rc = launch_upd64(s->ctx, U, b, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return FH_OK;
}
static int space_update_fused(fh_state* s, const double* mean_at_x, const double* mean_at_xnext, double sigma_t, const double* x,
                              const double* xnext, void* stream) {
  s->var2 = sigma_t * sigma_t;                 // H is rebuilt from Sigma at the true sigma (online_update_bfgs.py:294)
  for (int b = 0; b < s->B; ++b) {
    FhUpd64 U = upd64_state(s, b);
    U.mode = 2; U.sig2 = sigma_t * sigma_t;
    U.x = x + (size_t)b * s->d; U.xnext = xnext + (size_t)b * s->d;
    U.mean_at_x = mean_at_x + (size_t)b * s->d; U.mean_at_xnext = mean_at_xnext + (size_t)b * s->d;
    int rc = launch_upd64(s->ctx, U, b, (cudaStream_t)stream);
    if (rc) return rc;
  }
  s->m += 2;
  return FH_OK;
}

// update_time_step with float64 in and out: the transfer itself runs in the float32 state (DCT, low-rank algebra), but the
// result mu' = x + sigma'^2 s' is formed in float64 from the float64 x (|x| ~ sigma >> |mu'|: a float32 x would put its
// rounding error, 6e-8 |x|, straight into mu').
extern "C" int fh_time_update_f64(fh_state* s, const double* x, const double* score, double sigma_t, double sigma_tnext,
                                  int only_cov, double* mean_out, double* score_out, void* stream) {
  if (only_cov) return fh_time_update(s, nullptr, nullptr, sigma_t, sigma_tnext, 1, nullptr, nullptr, stream);
  if (!s || !x || !score || !mean_out || !score_out || !(sigma_t > 0) || !(sigma_tnext > 0)) FH_FAIL(FH_EINVAL, "bad fh_time_update_f64 arguments");
  if (upd64_applies(s)) return time_update_fused(s, x, nullptr, score, sigma_t, sigma_tnext, mean_out, score_out, stream);
  fh_ctx* c = s->ctx;
  const long tot = (long)s->B * s->d;
  const dim3 g((unsigned)((tot + 255) / 256)), b(256);
  FH_LAUNCH(fh_k_d2f, g, b, 0, stream, x, c->cgv[0], tot);
  FH_LAUNCH(fh_k_d2f, g, b, 0, stream, score, c->cgv[1], tot);
  FH_KCHECK();
  int rc = fh_time_update(s, c->cgv[0], c->cgv[1], sigma_t, sigma_tnext, 0, c->cgv[2], c->cgv[3], stream);
  if (rc) return rc;
  FH_LAUNCH(fh_k_tu_finish64, g, b, 0, stream, x, (const float*)c->cgv[3], sigma_tnext * sigma_tnext, mean_out, score_out, tot);
  FH_KCHECK();
  return FH_OK;
}

// Everything of the space update after dx (pixel space, scratch[0]) and de = sigma^2 (mu' - mu_hat) (scratch[1]) exist.
static int space_update_from_diffs(fh_state* s, double sigma_t, void* stream) {
  fh_ctx* c = s->ctx;
  const int vec = vec_ok(s->d);
  int rc;
  // H is rebuilt from Sigma at the true sigma (online_update_bfgs.py:294): the float32 drift of var2 is dropped.
  // Sigma itself stays where the last time update left it (s->sig2).
  s->var2 = sigma_t * sigma_t;
  float* dxp = c->scratch[0];
  float* dep = c->scratch[1];
  float* dx = c->scratch[2];
  float* de = c->scratch[3];
  if ((rc = to_basis(s, dxp, dx, stream))) return rc;
  if ((rc = to_basis(s, dep, de, stream))) return rc;
  // dots:  Z.de (rows 0..m-1) and dx.de (extra row dx)
  FH_LAUNCH(fh_k_dots, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->Z, s->m, de, (const float*)nullptr, dx, (const float*)nullptr, c->dot_partial, s->d, vec);
  FH_KCHECK();
  FH_LAUNCH(fh_k_dots_finalize, dim3(s->B), dim3(128), 0, stream, c->dot_partial, fh_nchunks(s->d), s->m, 1, s->Gz, c->csum, c->coef);
  FH_KCHECK();
  // rows m (dx) and m+1 (g = Sigma^-1 de = de./delta + Z^T Gz Z de); de.g partials -> part_b
  FH_LAUNCH(fh_k_su_rows, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->Z, s->m, c->coef, de, s->delta, dx, c->part_b, s->d, vec);
  FH_KCHECK();
  FH_LAUNCH(fh_k_bfgs_block, dim3((s->B + 63) / 64), dim3(64), 0, stream, s->Gz, s->m, c->csum, c->part_b, fh_nchunks(s->d), s->B);
  FH_KCHECK();
  s->m += 2;
  return refresh_core(s, stream);
}

// project_to_diagonal (online_update_bfgs.py:274-277): the BFGS pair is not appended but folded into the DIAGONAL part
// of Sigma's representation, diag <- diag + u.u - v.v with u = de sqrt(gamma), v = Sigma dx / sqrt(dx.Sigma dx), i.e.
//     D' = delta + gamma de^2 - (Sigma dx)^2 / (dx.Sigma dx)          (real also when gamma < 0: only squares enter)
// while the low-rank part L = delta Z^T Gs Z delta of Sigma stays.  In the precision form Sigma'^-1 = (D' + L)^-1 that is
//     Z' = (delta / D') Z (row-wise),  pi' = 1 / D' - sigma^-2,  Gz' = -Gs (I + K' Gs)^-1,  K' = Z (delta^2 / D') Z^T
// (Woodbury with the SAME small algebra as the per-sigma core, roles of Gz and Gs swapped).  The rank does not change.
// One deliberate difference: the reference appends (u, v) / sigma^2 to its Hessian factors ALSO in this mode (:296-298 run
// unconditionally), which makes its H inconsistent with its own Sigma; here H always follows from Sigma.
static int space_update_project(fh_state* s, double sigma_t, void* stream) {
  fh_ctx* c = s->ctx;
  const int vec = vec_ok(s->d);
  int rc;
  s->var2 = sigma_t * sigma_t;
  float *dxp = c->scratch[0], *dep = c->scratch[1], *dx = c->scratch[2], *de = c->scratch[3], *sdx = c->scratch[4], *wgt = c->scratch[5];
  if ((rc = to_basis(s, dxp, dx, stream))) return rc;
  if ((rc = to_basis(s, dep, de, stream))) return rc;
  if ((rc = cov_dot_basis(s, dx, sdx, stream))) return rc;                      // Sigma dx in the state's basis
  // dx.Sigma dx and dx.de: two "rows" against dx
  FH_LAUNCH(fh_k_dots, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->Z, 0, dx, (const float*)nullptr, (const float*)sdx, (const float*)de, c->dot_partial, s->d, vec);
  FH_KCHECK();
  FH_LAUNCH(fh_k_dots_finalize, dim3(s->B), dim3(128), 0, stream, c->dot_partial, fh_nchunks(s->d), 0, 2, (const double*)nullptr, c->csum, (double*)nullptr);
  FH_KCHECK();
  // D' and the weights delta^2 / D' (Gram) ; Z rows are rescaled by delta / D' after the Gram
  FH_LAUNCH(fh_k_su_project, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->delta, de, sdx, c->csum, (float)(1.0 / s->sig2), s->pi, wgt, c->scratch[6], s->d, vec);
  FH_KCHECK();
  if (s->m > 0) {
    FhRowPtrs xr{};
    if ((rc = gram(s, s->m, xr, 0, wgt, stream))) return rc;                    // K' = Z (delta^2 / D') Z^T
    const size_t smem = (size_t)(2 * FH_MAXM * FH_MAXM + 3 * FH_MAXM) * sizeof(double);
    FH_LAUNCH(fh_k_core, dim3(s->B), dim3(FH_CORE_THREADS), smem, stream, c->gramK, s->Gs, s->m, c->Gtmp, 0, 0.0, c->coef, c->status);   // -Gs (I + K' Gs)^-1
    FH_KCHECK();
    FH_CUDA(cudaMemcpyAsync(s->Gz, c->Gtmp, (size_t)s->B * FH_MAXM * FH_MAXM * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    FH_LAUNCH(fh_k_scale_rows, dim3((unsigned)fh_nchunks(s->d), s->B, s->m), dim3(FH_EW_THREADS), 0, stream, s->Z, c->scratch[6], s->d, vec);
    FH_KCHECK();
  }
  return refresh_core(s, stream);
}

extern "C" int fh_space_update(fh_state* s, const float* mean_at_x, const float* mean_at_xnext, double sigma_t,
                               const float* x, const float* xnext, int project_to_diagonal, void* stream) {
  if (!s || !mean_at_x || !mean_at_xnext || !x || !xnext || !(sigma_t > 0)) FH_FAIL(FH_EINVAL, "bad fh_space_update arguments");
  if (!project_to_diagonal && s->m + 2 > FH_MAXM) FH_FAIL(FH_ERANK, "low-rank capacity %d exhausted", FH_MAXM);
  fh_ctx* c = s->ctx;
  FH_LAUNCH(fh_k_su_diffs, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, x, xnext, mean_at_x, mean_at_xnext,
            (float)(sigma_t * sigma_t), c->scratch[0], c->scratch[1], s->d, vec_ok(s->d));
  FH_KCHECK();
  return project_to_diagonal ? space_update_project(s, sigma_t, stream) : space_update_from_diffs(s, sigma_t, stream);
}
// The same update from float64 inputs (the reference's x and denoiser means are float64, generate_conditional.py:116):
// dx = x' - x and de = sigma^2 (mu' - mu_hat) are DIFFERENCES of nearly equal images, so they are formed in float64 and
// only then rounded to the float32 the state stores (forming them from float32-rounded inputs costs two to three digits).
extern "C" int fh_space_update_f64(fh_state* s, const double* mean_at_x, const double* mean_at_xnext, double sigma_t,
                                   const double* x, const double* xnext, int project_to_diagonal, void* stream) {
  if (!s || !mean_at_x || !mean_at_xnext || !x || !xnext || !(sigma_t > 0)) FH_FAIL(FH_EINVAL, "bad fh_space_update_f64 arguments");
  if (!project_to_diagonal && s->m + 2 > FH_MAXM) FH_FAIL(FH_ERANK, "low-rank capacity %d exhausted", FH_MAXM);
  if (!project_to_diagonal && upd64_applies(s)) return space_update_fused(s, mean_at_x, mean_at_xnext, sigma_t, x, xnext, stream);
  fh_ctx* c = s->ctx;
  const long tot = (long)s->B * s->d;
  FH_LAUNCH(fh_k_su_diffs64, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, stream, x, xnext, mean_at_x, mean_at_xnext, sigma_t * sigma_t,
            c->scratch[0], c->scratch[1], tot);
  FH_KCHECK();
  return project_to_diagonal ? space_update_project(s, sigma_t, stream) : space_update_from_diffs(s, sigma_t, stream);
}

extern "C" int fh_state_set_cov_vectors(fh_state* s, const float* U, int ku, const float* V, int kv, double sigma, void* stream) {
  if (!s || ku < 0 || kv < 0 || ku + kv > FH_MAXM || !(sigma > 0)) FH_FAIL(FH_EINVAL, "bad fh_state_set_cov_vectors arguments");
  fh_ctx* c = s->ctx;
  const int vec = vec_ok(s->d);
  int rc;
  // current Sigma diagonal stays (delta at the current var2); Z = W ./ delta, Gz = -J (I + Kp J)^-1, Kp = W P W^T
  if ((rc = refresh_delta(s, stream))) return rc;
  const int m = ku + kv;
  for (int b = 0; b < s->B; ++b)
    for (int i = 0; i < m; ++i) {
      const float* src = (i < ku) ? U + (size_t)i * s->d : V + (size_t)(i - ku) * s->d;
      float* dst = s->Z + ((size_t)b * FH_MAXM + i) * s->d;
      if (s->basis == FH_BASIS_DCT) { if ((rc = transpose_planes(c, src, dst, c->C, stream))) return rc; }
      else FH_CUDA(cudaMemcpyAsync(dst, src, (size_t)s->d * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
  // w = 1/delta = P  -> scratch[0];  Kp = W diag(P) W^T
  FH_LAUNCH(fh_k_delta, ew_grid(s->d, s->B), dim3(FH_EW_THREADS), 0, stream, s->delta, c->scratch[0], 0.0f, s->d, vec);   // 1/(delta+0)
  FH_KCHECK();
  s->m = m;
  FhRowPtrs xr{};
  if ((rc = gram(s, m, xr, 0, c->scratch[0], stream))) return rc;
  const size_t smem = (size_t)(2 * FH_MAXM * FH_MAXM + 3 * FH_MAXM) * sizeof(double);
  FH_LAUNCH(fh_k_core_from_cov, dim3(s->B), dim3(256), smem, stream, c->gramK, ku, kv, s->Gz, c->status);
  FH_KCHECK();
  // Z <- W .* P  (row by row, in place)
  for (int b = 0; b < s->B; ++b)
    for (int i = 0; i < m; ++i) {
      float* row = s->Z + ((size_t)b * FH_MAXM + i) * s->d;
      FH_LAUNCH(fh_k_lowrank_apply, dim3((unsigned)fh_nchunks(s->d), 1), dim3(FH_EW_THREADS), 0, stream, s->Z, 0, c->coef, c->scratch[0] + (size_t)b * s->d, row, 1.0f, row, s->d, vec);
      FH_KCHECK();
    }
  // the diagonal of Sigma^-1 is P = 1/delta exactly as before: pi unchanged.  sigma fixes the H relation.
  (void)sigma;         // Sigma's diagonal is defined at the current var2; `sigma` only fixes the H relation
  s->var2 = s->sig2;
  return refresh_core(s, stream);
}

extern "C" int fh_state_drop_all(fh_state* s, double sigma, void* stream) {
  if (!s) FH_FAIL(FH_EINVAL, "null state");
  // Sigma <- its diagonal part delta at the current sigma:  pi = 1/delta - tau = pi (unchanged), Z dropped.
  (void)sigma;
  s->m = 0;
  FH_CUDA(cudaMemsetAsync(s->Gz, 0, (size_t)s->B * FH_MAXM * FH_MAXM * sizeof(double), (cudaStream_t)stream));
  FH_CUDA(cudaMemsetAsync(s->Gs, 0, (size_t)s->B * FH_MAXM * FH_MAXM * sizeof(double), (cudaStream_t)stream));
  return FH_OK;
}

extern "C" int fh_state_gram(fh_state* s, double* K_out, void* stream) {
  if (!s || !K_out) FH_FAIL(FH_EINVAL, "bad fh_state_gram arguments");
  if (s->m == 0) return FH_OK;
  fh_ctx* c = s->ctx;
  FhRowPtrs xr{};
  int rc = gram(s, s->m, xr, 0, s->delta, stream);
  if (rc) return rc;
  for (int b = 0; b < s->B; ++b)
    for (int i = 0; i < s->m; ++i)
      FH_CUDA(cudaMemcpyAsync(K_out + ((size_t)b * s->m + i) * s->m, c->gramK + (size_t)b * FH_GRAM_LD * FH_GRAM_LD + (size_t)i * FH_GRAM_LD,
                              (size_t)s->m * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return FH_OK;
}

extern "C" int fh_state_recombine(fh_state* s, int m_new, const double* T, const double* Gz_new, void* stream) {
  if (!s || !T || !Gz_new || m_new < 1 || m_new > s->m) FH_FAIL(FH_EINVAL, "bad fh_state_recombine arguments (rank %d -> %d)", s ? s->m : -1, m_new);
  FH_LAUNCH(fh_k_recombine, dim3((unsigned)fh_nchunks(s->d), s->B), dim3(256), 0, stream, s->Z, s->m, m_new, T, s->d);
  FH_KCHECK();
  FH_CUDA(cudaMemsetAsync(s->Gz, 0, (size_t)s->B * FH_MAXM * FH_MAXM * sizeof(double), (cudaStream_t)stream));
  FH_CUDA(cudaMemsetAsync(s->Gs, 0, (size_t)s->B * FH_MAXM * FH_MAXM * sizeof(double), (cudaStream_t)stream));
  for (int b = 0; b < s->B; ++b)
    for (int i = 0; i < m_new; ++i)
      FH_CUDA(cudaMemcpyAsync(s->Gz + ((size_t)b * FH_MAXM + i) * FH_MAXM, Gz_new + ((size_t)b * m_new + i) * m_new,
                              (size_t)m_new * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  s->m = m_new;
  return refresh_core(s, stream);
}

extern "C" int fh_state_export(fh_state* s, float* pi, float* Z, double* Gz, double* Gs, float* delta, void* stream) {
  if (!s) FH_FAIL(FH_EINVAL, "null state");
  fh_ctx* c = s->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // DCT basis: the internal layout is (c, kj, ki); exported vectors are in the natural (c, ki, kj) order
  auto put = [&](float* dst, const float* src, int nimg) -> int {
    if (s->basis == FH_BASIS_DCT) return transpose_planes(c, src, dst, nimg * c->C, stream);
    FH_CUDA(cudaMemcpyAsync(dst, src, (size_t)nimg * s->d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return FH_OK;
  };
  if (pi && (rc = put(pi, s->pi, s->B))) return rc;
  if (delta && (rc = put(delta, s->delta, s->B))) return rc;
  if (Z)
    for (int b = 0; b < s->B; ++b)
      if (s->m && (rc = put(Z + (size_t)b * s->m * s->d, s->Z + (size_t)b * FH_MAXM * s->d, s->m))) return rc;
  for (int which = 0; which < 2; ++which) {
    double* dst = which ? Gs : Gz;
    const double* src = which ? s->Gs : s->Gz;
    if (!dst) continue;
    for (int b = 0; b < s->B; ++b)
      for (int i = 0; i < s->m; ++i)
        FH_CUDA(cudaMemcpyAsync(dst + ((size_t)b * s->m + i) * s->m, src + ((size_t)b * FH_MAXM + i) * FH_MAXM, (size_t)s->m * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  return FH_OK;
}

// --------------------------------------------------------------------------------------- operators
extern "C" int fh_op_create_inpaint(fh_ctx* c, fh_op** out, const float* mask_dev, int mask_batch) {
  if (!c || !out || !mask_dev || mask_batch < 1) FH_FAIL(FH_EINVAL, "bad fh_op_create_inpaint arguments");
  fh_op* o = (fh_op*)calloc(1, sizeof(fh_op));
  o->ctx = c; o->kind = FH_OP_INPAINT; o->sf = 1; o->mask = mask_dev; o->mask_batch = mask_batch;
  *out = o;
  return FH_OK;
}

extern "C" int fh_op_create_conv(fh_ctx* c, fh_op** out, const float* psf, int ks, int sf) {
  if (!c || !out || !psf || ks < 1 || (ks & 1) == 0 || sf < 1) FH_FAIL(FH_EINVAL, "bad fh_op_create_conv arguments (odd ks required)");
  if (c->H % sf || c->W % sf) FH_FAIL(FH_EINVAL, "image size not divisible by sf");
  if (sf > 1 && ((c->H & (c->H - 1)) || (c->W & (c->W - 1)))) FH_FAIL(FH_EINVAL, "SR needs power-of-two image sizes");
  fh_op* o = (fh_op*)calloc(1, sizeof(fh_op));
  o->ctx = c; o->kind = sf > 1 ? FH_OP_SR : FH_OP_BLUR; o->sf = sf;
  const int cen = ks / 2;                       // utils_sisr.py:37  roll by -int(ks/2)
  std::vector<int> dy, dx; std::vector<float> w;
  int ymin = 1 << 30, ymax = -(1 << 30), xmin = 1 << 30, xmax = -(1 << 30);
  for (int a = 0; a < ks; ++a)
    for (int b = 0; b < ks; ++b) {
      const float v = psf[a * ks + b];
      if (v != 0.f) {
        dy.push_back(a - cen); dx.push_back(b - cen); w.push_back(v);
        ymin = std::min(ymin, a - cen); ymax = std::max(ymax, a - cen);
        xmin = std::min(xmin, b - cen); xmax = std::max(xmax, b - cen);
      }
    }
  if (w.empty()) { free(o); FH_FAIL(FH_EINVAL, "PSF is all zeros"); }
  o->ks = ks;
  FH_CUDA(cudaMalloc((void**)&o->psf_dense, (size_t)ks * ks * sizeof(float)));
  FH_CUDA(cudaMemcpy(o->psf_dense, psf, (size_t)ks * ks * sizeof(float), cudaMemcpyHostToDevice));
  if (ymax - ymin >= c->H || xmax - xmin >= c->W) { free(o); FH_FAIL(FH_EINVAL, "PSF support larger than the image"); }
  o->nt = (int)w.size(); o->ymin = ymin; o->ymax = ymax; o->xmin = xmin; o->xmax = xmax;
  FH_CUDA(cudaMalloc((void**)&o->t_dy, w.size() * sizeof(int)));
  FH_CUDA(cudaMalloc((void**)&o->t_dx, w.size() * sizeof(int)));
  FH_CUDA(cudaMalloc((void**)&o->t_w, w.size() * sizeof(float)));
  FH_CUDA(cudaMemcpy(o->t_dy, dy.data(), w.size() * sizeof(int), cudaMemcpyHostToDevice));
  FH_CUDA(cudaMemcpy(o->t_dx, dx.data(), w.size() * sizeof(int), cudaMemcpyHostToDevice));
  FH_CUDA(cudaMemcpy(o->t_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  {
    // one dense vertical run per PSF column (fh_k_conv_runs): weights in window order for both directions, every run
    // padded to a multiple of 4 with zeros
    std::vector<int4> runs; std::vector<float> wf, wt;
    for (int b = 0; b < ks; ++b) {
      int lo = -1, hi = -1;
      for (int a = 0; a < ks; ++a) if (psf[a * ks + b] != 0.f) { if (lo < 0) lo = a; hi = a; }
      if (lo < 0) continue;
      const int L = hi - lo + 1, Lp = (L + 3) & ~3;
      runs.push_back(make_int4(b - cen, lo - cen, L, (int)wf.size()));
      for (int k = 0; k < Lp; ++k) {
        wf.push_back(k < L ? psf[(hi - k) * ks + b] : 0.f);      // forward: the window starts at the LAST tap of the run
        wt.push_back(k < L ? psf[(lo + k) * ks + b] : 0.f);
      }
    }
    o->nruns = (int)runs.size(); o->nrun_taps = (int)wf.size();
    FH_CUDA(cudaMalloc((void**)&o->runs, runs.size() * sizeof(int4)));
    FH_CUDA(cudaMalloc((void**)&o->run_wf, wf.size() * sizeof(float)));
    FH_CUDA(cudaMalloc((void**)&o->run_wt, wt.size() * sizeof(float)));
    FH_CUDA(cudaMemcpy(o->runs, runs.data(), runs.size() * sizeof(int4), cudaMemcpyHostToDevice));
    FH_CUDA(cudaMemcpy(o->run_wf, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice));
    FH_CUDA(cudaMemcpy(o->run_wt, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
    // float64 copy for the resident float64 solve: runs padded to a multiple of 6 taps (its window rotates in rounds of 6)
    std::vector<int4> runs6; std::vector<double> wf6, wt6;
    for (int b = 0; b < ks; ++b) {
      int lo = -1, hi = -1;
      for (int a = 0; a < ks; ++a) if (psf[a * ks + b] != 0.f) { if (lo < 0) lo = a; hi = a; }
      if (lo < 0) continue;
      const int L = hi - lo + 1, L6 = (L + 5) / 6 * 6;
      runs6.push_back(make_int4(b - cen, lo - cen, L, (int)wf6.size()));
      for (int k = 0; k < L6; ++k) {
        wf6.push_back(k < L ? (double)psf[(hi - k) * ks + b] : 0.0);
        wt6.push_back(k < L ? (double)psf[(lo + k) * ks + b] : 0.0);
      }
    }
    o->nrun_taps6 = (int)wf6.size();
    FH_CUDA(cudaMalloc((void**)&o->runs6, runs6.size() * sizeof(int4)));
    FH_CUDA(cudaMalloc((void**)&o->run_wf6, wf6.size() * sizeof(double)));
    FH_CUDA(cudaMalloc((void**)&o->run_wt6, wt6.size() * sizeof(double)));
    FH_CUDA(cudaMemcpy(o->runs6, runs6.data(), runs6.size() * sizeof(int4), cudaMemcpyHostToDevice));
    FH_CUDA(cudaMemcpy(o->run_wf6, wf6.data(), wf6.size() * sizeof(double), cudaMemcpyHostToDevice));
    FH_CUDA(cudaMemcpy(o->run_wt6, wt6.data(), wt6.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  // rank-1 test on the support box: psf == col (x) row exactly (to 1e-7 of the peak) -> separable path
  o->separable = false;
  if (sf == 1) {
    const int hh = ymax - ymin + 1, ww = xmax - xmin + 1;
    int pa = 0, pb = 0; double peak = 0;
    for (int a = 0; a < hh; ++a) for (int b = 0; b < ww; ++b) {
      const double v = fabs((double)psf[(a + ymin + cen) * ks + (b + xmin + cen)]);
      if (v > peak) { peak = v; pa = a; pb = b; }
    }
    std::vector<double> col(hh), row(ww);
    for (int a = 0; a < hh; ++a) col[a] = psf[(a + ymin + cen) * ks + (pb + xmin + cen)];
    for (int b = 0; b < ww; ++b) row[b] = (double)psf[(pa + ymin + cen) * ks + (b + xmin + cen)] / col[pa];
    double err = 0;
    for (int a = 0; a < hh; ++a) for (int b = 0; b < ww; ++b)
      err = std::max(err, fabs(col[a] * row[b] - (double)psf[(a + ymin + cen) * ks + (b + xmin + cen)]));
    if (err <= 2e-7 * peak) {
      o->separable = true;
      std::vector<float> kc(hh), kr(ww);
      for (int a = 0; a < hh; ++a) kc[a] = (float)col[a];
      for (int b = 0; b < ww; ++b) kr[b] = (float)row[b];
      o->nc = hh; o->nr = ww; o->cc = -ymin; o->cr = -xmin;
      FH_CUDA(cudaMalloc((void**)&o->kc, hh * sizeof(float)));
      FH_CUDA(cudaMalloc((void**)&o->kr, ww * sizeof(float)));
      FH_CUDA(cudaMemcpy(o->kc, kc.data(), hh * sizeof(float), cudaMemcpyHostToDevice));
      FH_CUDA(cudaMemcpy(o->kr, kr.data(), ww * sizeof(float), cudaMemcpyHostToDevice));
    }
  }
  *out = o;
  return FH_OK;
}

extern "C" int fh_op_destroy(fh_op* o) {
  if (!o) return FH_OK;
  cudaFree(o->t_dy); cudaFree(o->t_dx); cudaFree(o->t_w); cudaFree(o->kr); cudaFree(o->kc); cudaFree(o->psf_dense);
  cudaFree(o->runs); cudaFree(o->run_wf); cudaFree(o->run_wt); cudaFree(o->runs6); cudaFree(o->run_wf6); cudaFree(o->run_wt6);
  free(o);
  return FH_OK;
}

static long op_ylen(const fh_op* o) {
  const fh_ctx* c = o->ctx;
  return (long)c->C * (c->H / o->sf) * (c->W / o->sf);
}

// The column-run kernel is used for every batch size: choosing by the number of planes (the tap list is 2 us faster
// per launch at 3 planes: 21 vs 23 us) would make a batch-4 solve differ from four batch-1 solves in the last bits.
static inline bool conv_runs_pays(const fh_ctx* c, int planes) {
#ifdef FH_EMU
  return getenv("FH_EMU_CONV_RUNS") != nullptr;           // CPU emulation: the test picks the kernel, default tap list
#endif
  (void)c; (void)planes;
  return true;
}

extern "C" int fh_op_apply(fh_op* o, const float* in, float* out, int batch, int transpose, void* stream) {
  if (!o || !in || !out || batch < 1) FH_FAIL(FH_EINVAL, "bad fh_op_apply arguments");
  if (o->kind == FH_OP_INPAINT && o->mask_batch != 1 && o->mask_batch != batch)
    FH_FAIL(FH_EINVAL, "mask batch %d does not match the batch %d (one mask for all images, or one per image)", o->mask_batch, batch);
  fh_ctx* c = o->ctx;
  const int planes = batch * c->C;
  if (o->kind == FH_OP_INPAINT) {
    const long hw = (long)c->H * c->W;
    FH_LAUNCH(fh_k_mask_mul, dim3((unsigned)((c->d + 255) / 256), batch), dim3(256), 0, stream, in, o->mask, o->mask_batch, out, c->C, hw);
  } else if (o->kind == FH_OP_BLUR) {
    if (c->H % 32 || c->W % 32) FH_FAIL(FH_EINVAL, "blur needs image sizes divisible by 32");
    dim3 grid(c->W / 32, c->H / 32, planes), block(32, 8);
    if (o->separable) {
      FhSepTaps k{o->kr, o->kc, o->nr, o->nc, o->cr, o->cc};
      const int th = 32 + o->nc - 1, tw = (32 + o->nr - 1) | 1;
      const size_t smem = (size_t)(th * tw + th * 33) * sizeof(float);
      FH_LAUNCH(fh_k_conv_sep, grid, block, smem, stream, in, out, k, c->H, c->W, transpose);
    } else if (conv_runs_pays(c, planes) && c->H % 64 == 0 && o->nruns <= FH_CONV_MAX_RUNS && o->nrun_taps <= FH_CONV_MAX_TAPS && o->nrun_taps <= 2 * o->nt &&
               (size_t)(64 + o->ymax - o->ymin + FH_CONV_PAD_ROWS) * ((32 + o->xmax - o->xmin) | 1) * sizeof(float) <= 40 * 1024) {
      // 64 x 32 tiles: 84 vs 115 us at 48 planes, 23 vs 21 us at 3 planes (latency-bound there; 32-row tiles: 26 us).
      // A PSF whose columns are mostly gaps (more than half of the run taps zero) stays on the tap list.
      FhRuns r{o->runs, o->run_wf, o->run_wt, o->nruns, o->nrun_taps, o->ymin, o->ymax, o->xmin, o->xmax};
      const int th = 64 + (o->ymax - o->ymin) + FH_CONV_PAD_ROWS, tw = (32 + (o->xmax - o->xmin)) | 1;
      const size_t smem = (size_t)th * tw * sizeof(float);
      FH_LAUNCH(fh_k_conv_runs<64>, dim3(c->W / 32, c->H / 64, planes), dim3(32, 8), smem, stream, in, out, r, c->H, c->W, transpose);
    } else {
      FhTaps t{o->t_dy, o->t_dx, o->t_w, o->nt, o->ymin, o->ymax, o->xmin, o->xmax};
      const int th = 32 + (o->ymax - o->ymin), tw = (32 + (o->xmax - o->xmin)) | 1;
      const size_t smem = (size_t)th * tw * sizeof(float);
      FH_LAUNCH(fh_k_conv_taps, grid, block, smem, stream, in, out, t, c->H, c->W, transpose);
    }
  } else {
    const bool fast = c->H == 256 && c->W == 256 && o->sf == 4 && o->ks <= FH_SR_KS && ((o->ks / 2) & 3) == 0;
    if (fast) {
      if (!transpose) FH_LAUNCH(fh_k_sr4_forward, dim3(16, planes), dim3(256), 0, stream, in, out, (const float*)o->psf_dense, o->ks);
      else FH_LAUNCH(fh_k_sr4_transpose, dim3(16, planes), dim3(256), 0, stream, in, out, (const float*)o->psf_dense, o->ks);
    } else {
      FhTaps t{o->t_dy, o->t_dx, o->t_w, o->nt, o->ymin, o->ymax, o->xmin, o->xmax};
      const size_t smem = (size_t)o->nt * 3 * sizeof(float);
      if (!transpose) {
        const int n = (c->H / o->sf) * (c->W / o->sf);
        FH_LAUNCH(fh_k_sr_forward, dim3((n + 255) / 256, planes), dim3(256), smem, stream, in, out, t, c->H, c->W, o->sf);
      } else {
        const int n = c->H * c->W;
        FH_LAUNCH(fh_k_sr_transpose, dim3((n + 255) / 256, planes), dim3(256), smem, stream, in, out, t, c->H, c->W, o->sf);
      }
    }
  }
  FH_KCHECK();
  return FH_OK;
}

// --------------------------------------------------------------------------------------------- CG
// t = A Sigma A^T p   (y-space in, y-space out); for inpainting the two mask multiplies are folded into the
// callers (cg_q applies the outer mask; the inner one is applied here).
static int apply_ASAt(fh_state* s, fh_op* o, const float* p, float* t, void* stream) {
  fh_ctx* c = s->ctx;
  int rc;
  float* img = c->scratch[0];
  if ((rc = fh_op_apply(o, p, img, s->B, 1, stream))) return rc;            // A^T p   (inpaint: mask.*p)
  if (s->basis == FH_BASIS_DCT) {
    if ((rc = to_basis(s, img, c->scratch[1], stream))) return rc;
    if ((rc = cov_dot_basis(s, c->scratch[1], c->scratch[1], stream))) return rc;
    if ((rc = from_basis(s, c->scratch[1], img, stream))) return rc;
  } else {
    if ((rc = cov_dot_basis(s, img, img, stream))) return rc;
  }
  if (o->kind == FH_OP_INPAINT) {
    // outer mask is applied inside fh_k_cg_q
    FH_CUDA(cudaMemcpyAsync(t, img, (size_t)s->B * s->d * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return FH_OK;
  }
  return fh_op_apply(o, img, t, s->B, 0, stream);
}

// Cooperative launch of the resident CG kernel (all CTAs co-resident: the kernel has grid barriers).
static int launch_resident(fh_ctx* c, const FhCgFused& P, FhCgResident R, cudaStream_t st) {
  const dim3 grid((unsigned)P.strips_per_image), block(FH_PASS_THREADS);
  {
    // exchange buffers and a fresh range of sequence numbers (launches on one ctx are stream-ordered)
    const size_t nblk = grid.x, half = (size_t)c->d / 2;
    R.ll_T[0] = c->res_ll; R.ll_T[1] = c->res_ll + half;
    R.ll_dots = c->res_ll + 2 * half;
    R.ll_sum = R.ll_dots + FH_MAXM * nblk;
    R.ll_a = R.ll_sum + FH_MAXM * FH_LL_SPREAD; R.ll_b = R.ll_a + nblk * FH_LL_SPREAD;
    const unsigned need = (unsigned)P.last_allowed + 2u;
    if (c->res_seq > 0xffffffffu - need - 1u) {
      FH_CUDA(cudaMemsetAsync(c->res_ll, 0, c->res_ll_bytes, st));
      c->res_seq = 0;
    }
    R.seq_base = c->res_seq;
    c->res_seq += need;
    R.grid_counter = c->grid_counter + R.image;
    FH_CUDA(cudaMemsetAsync(R.grid_counter, 0, sizeof(unsigned), st));
  }
  const bool big = P.m > FH_RES_MAXM;                      // rows above 32 are re-read from L2 (fh_k_cg_resident<true>)
  const size_t smem = sizeof(FhResSmem) + (size_t)std::min(P.m, FH_RES_MAXM) * FH_RES_SLICE * sizeof(float);
#ifdef FH_EMU
  (void)c; (void)st;
  if (big) emu::launch_coop([&] { fh_k_cg_resident<true>(P, R); }, grid, block, smem);
  else emu::launch_coop([&] { fh_k_cg_resident<false>(P, R); }, grid, block, smem);
  ++fh_g_launches;
#else
  (void)c;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (fh_g_prof_on) fh_prof_begin("fh_k_cg_resident", st);
  if (big) FH_CUDA(cudaLaunchKernelEx(&cfg, fh_k_cg_resident<true>, P, R));
  else FH_CUDA(cudaLaunchKernelEx(&cfg, fh_k_cg_resident<false>, P, R));
  if (fh_g_prof_on) fh_prof_end(st);
  ++fh_g_launches;
#endif
  return FH_OK;
}

// One image at a time in ~18 us per iteration against ~38 us (B = 1) ... 14 us per image (B = 16) for the streaming
// iteration that takes all images per launch: FH_CG_AUTO goes resident up to FH_RES_AUTO_MAX_BATCH images.
#define FH_RES_AUTO_MAX_BATCH 4
static FhR64Gen res64_gen(const fh_op* o, int m, int smem_room);
// `o`: the operator (the float64 kernel also takes the non-separable ones through its gathers; the float32 kernel only
// those that fold into the DCT passes, P.fuse != 0)
static bool resident_applies(const fh_state* s, const FhCgFused& P, const fh_op* o = nullptr) {
  const fh_ctx* c = s->ctx;
  if (!c->res_ok || s->m > FH_MAXM || c->cg_mode == FH_CG_STREAMING) return false;
  const bool f64 = c->cg_mode != FH_CG_RESIDENT_F32 && c->r64_ok;
  if (P.fuse == 0 && !(f64 && o && res64_gen(o, 0, c->r64_smem_max - (int)sizeof(FhRes64Smem)).mode != 0)) return false;
  return c->cg_mode == FH_CG_RESIDENT || c->cg_mode == FH_CG_RESIDENT_F32 || s->B <= FH_RES_AUTO_MAX_BATCH;
}
// which resident kernel: float64 (the parity path; FH_CG_AUTO, FH_CG_RESIDENT) or the float32 one (FH_CG_RESIDENT_F32)
static bool resident_is_f64(const fh_ctx* c) { return c->cg_mode != FH_CG_RESIDENT_F32 && c->r64_ok; }

// Whole solve in one launch per image (fh_cg_resident.cuh): Z stays in shared memory, x/r/p/q in registers.
static int cg_solve_resident(fh_state* s, fh_op* o, FhCgFused P, const float* y, const float* mu, double rtol, int last,
                             int use_stop, float* mat_out, int* niter_host, int* optimal_host, double* resid_host, void* stream,
                             FhCgInfo* info_dev = nullptr) {
  fh_ctx* c = s->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  const int B = s->B;
  const long n = s->d;                                    // mask and blur: measurement space == image space
  const long hw = (long)c->H * c->W;
  float *bv = c->cgv[0], *x = c->cgv[1], *t = c->cgv[5];
  int rc;
  const float* amu = mu;
  if (o->kind != FH_OP_INPAINT) {
    if ((rc = fh_op_apply(o, mu, t, B, 0, stream))) return rc;
    amu = t;
  }
  FH_LAUNCH(fh_k_cg_rhs, dim3((unsigned)((n + 255) / 256), B), dim3(256), 0, stream, y, amu, P.mask, P.mask_batch, hw, bv, x, n);
  FH_KCHECK();
  FH_CUDA(cudaMemsetAsync(c->cg_sc, 0, (size_t)B * sizeof(FhCgScalars), st));
  P.strips_per_image = (int)(s->d / FH_RES_SLICE);
  P.ystrips = P.strips_per_image;
  P.last_allowed = last;
  P.mirror = nullptr;
  float* xout = (o->kind == FH_OP_INPAINT) ? mat_out : x;  // inpainting returns u itself
  for (int b = 0; b < B; ++b) {
    FhCgResident R{};
    R.bvec = bv + (size_t)b * n; R.x_out = xout + (size_t)b * n;
    R.rtol = rtol; R.atol = 0.0; R.use_stop = use_stop; R.image = b;
    R.info_out = info_dev ? info_dev + b : nullptr;
    if ((rc = launch_resident(c, P, R, st))) return rc;
  }
  if (info_dev) {                                          // asynchronous: the outcome stays on the device
    if (o->kind == FH_OP_INPAINT) return FH_OK;
    return fh_op_apply(o, x, mat_out, B, 1, stream);
  }
  FH_CUDA(cudaMemcpyAsync(c->cg_sc_host, c->cg_sc, (size_t)B * sizeof(FhCgScalars), cudaMemcpyDeviceToHost, st));
  FH_CUDA(cudaStreamSynchronize(st));
  for (int b = 0; b < B; ++b) {
    if (!c->cg_sc_host[b].done) FH_FAIL(FH_ESTATE, "resident CG kernel did not finish image %d", b);
    if (niter_host) niter_host[b] = c->cg_sc_host[b].niter;
    if (optimal_host) optimal_host[b] = c->cg_sc_host[b].optimal;
    if (resid_host) resid_host[b] = sqrt(c->cg_sc_host[b].resid2);
  }
  if (o->kind == FH_OP_INPAINT) return FH_OK;
  return fh_op_apply(o, x, mat_out, B, 1, stream);
}

// ---- float64 resident solve (fh_cg_res64.cuh) -------------------------------------------------------------------------
// Shared memory the 2-D PSF gather needs for chunks of `cw` output columns.
static size_t res64_conv_gbuf(const fh_op* o, int cw) {
  const int th = FH_RES_ROWS + (o->ymax - o->ymin) + FH_RES_ROWS, tws = (cw + (o->xmax - o->xmin) + 2) & ~1;
  const size_t bytes = ((size_t)th * tws + (size_t)(FH_PASS_THREADS / cw - 1) * FH_RES_ROWS * cw + ((o->nrun_taps6 + 1) & ~1)) * sizeof(double) +
                       (size_t)o->nruns * sizeof(int4);
  return (bytes + 15) & ~(size_t)15;
}
// The non-separable operators' description for the float64 kernel at rank m (mode 0 when the operator folds into the DCT
// passes or has no gathers here: the caller then falls back to the streaming float32 iteration).  The 2-D PSF gather
// trades shared memory against L2 round trips: every further chunk is a round trip more, and every 6 KB given to the staging
// is a row of Z that has to be re-read from L2 in every iteration -- the chunk width minimising the (measured) sum.
static FhR64Gen res64_gen(const fh_op* o, int m = 0, int smem_room = 180 * 1024) {
  FhR64Gen G{};
  if (!o) return G;
  const fh_ctx* c = o->ctx;
  if (o->kind == FH_OP_BLUR && !o->separable && o->nruns > 0 && o->nruns <= 64 && o->nrun_taps <= 2048) {
    double best = 1e30;
    const char* force = getenv("FH_R64_CW");                 // A/B runs: force the chunk width (256, 128, 64, 32)
    for (int cw = 256; cw >= 32; cw >>= 1) {
      if (force && atoi(force) != cw) continue;
      const size_t bytes = res64_conv_gbuf(o, cw);
      if ((long)bytes > (long)smem_room) continue;
      const int cap = std::min(FH_RES_MAXM, (int)((smem_room - (long)bytes) / (FH_RES_SLICE * 4)));
      // us per iteration relative to one 256-column chunk with every row resident, fitted to measurements at ranks 0 .. 56
      // (FH_R64_CW A/B runs): a second chunk costs 3.0, four chunks 14.3; leaving the all-resident instantiation costs ~1.5
      // and 0.3 per re-read row for the first 8 rows, little beyond.  256 columns win everywhere except for ranks 11 .. 16,
      // where 128 columns keep every row resident (rank 56: 64.3 against 69.0 us with the previous model's choice of 128).
      const int nch = FH_N / cw, nl2 = std::max(0, m - cap);
      const double cost = (nch == 1 ? 0.0 : nch == 2 ? 3.0 : nch == 4 ? 14.3 : 35.0) +
                          (nl2 > 0 ? 1.5 + 0.3 * std::min(nl2, 8) + 0.05 * std::max(0, nl2 - 8) : 0.0);
      if (cost < best) { best = cost; G.cw = cw; G.gbuf_bytes = (int)bytes; }
    }
    if (G.cw) {
      G.mode = 3; G.runs = o->runs6; G.w_fwd = o->run_wf6; G.w_tr = o->run_wt6; G.nruns = o->nruns; G.ntaps = o->nrun_taps6;
      G.ymin = o->ymin; G.ymax = o->ymax; G.xmin = o->xmin; G.xmax = o->xmax;
    }
  } else if (o->kind == FH_OP_SR && c->H == 256 && c->W == 256 && o->sf == 4 && o->ks <= FH_SR_KS && ((o->ks / 2) & 3) == 0) {
    G.mode = 4; G.psf = o->psf_dense; G.ks = o->ks;
    // PSF rows, 10 measurement rows (A^T), 128 results, the image windows of the (at most two) runs of vertically adjacent groups (A)
    G.gbuf_bytes = (int)(((size_t)(FH_SR_KS * FH_R64_SRP + 10 * 64 + 128 + FH_R64_SR_WROWS * (64 + 24)) * sizeof(double) + 15) & ~(size_t)15);
  }
  return G;
}

static int launch_res64(fh_ctx* c, const FhCgFused& P, FhCgRes64 R, cudaStream_t st) {
  const dim3 grid((unsigned)(c->C * FH_R64_CPP)), block(FH_PASS_THREADS);
  const size_t nblk = grid.x, half = (size_t)c->d / 2;
  R.T[0] = c->r64_T[0]; R.T[1] = c->r64_T[1];
  R.ride[0] = c->r64_ride[0]; R.ride[1] = c->r64_ride[1];
  R.ll_dots = c->res_ll + 2 * half;
  R.ll_sum = R.ll_dots + FH_MAXM * nblk;
  R.tw = c->tw64; R.e512 = c->e512;
  {
    const char* e = getenv("FH_R64_ZT");              // 0: never, 1: always, 2 (default): from the previous solve's iteration count
    R.zt_mode = e ? atoi(e) : 2;                     // R.niter_hint: set by the caller from the state
  }
  const unsigned need = (unsigned)P.last_allowed + 3u;
  if (c->res_seq > 0xffffffffu - need - 1u) {
    FH_CUDA(cudaMemsetAsync(c->res_ll, 0, c->res_ll_bytes, st));
    c->res_seq = 0;
  }
  R.seq_base = c->res_seq;
  c->res_seq += need;
  R.grid_counter = c->grid_counter + R.image;
  FH_CUDA(cudaMemsetAsync(R.grid_counter, 0, sizeof(unsigned), st));
  // Shared memory: fixed part, [twiddle tables], [gather staging of a non-separable operator], resident rows of Z.
  // Three instantiations: tables in shared memory while the rank leaves room for them, through L1 above that, and the
  // partially resident one above the capacity.
  const int slice = FH_RES_SLICE * (int)sizeof(float);
  const int room = c->r64_smem_max - (int)sizeof(FhRes64Smem) - R.gen.gbuf_bytes;
  const int cap = std::max(0, std::min(FH_RES_MAXM, room / slice));
  const int cap_tabs = std::max(-1, std::min(FH_RES_MAXM, (room - (int)sizeof(FhRes64Tabs)) / slice));
  if (room < 0) FH_FAIL(FH_EINVAL, "operator staging does not fit the float64 kernel's shared memory");
  const bool tabs = P.m <= cap_tabs;
  const bool big = !tabs && P.m > cap;
  R.mr_cap = cap;
  // Zt mode with float64 rows: twice the slots, while they fit beside the tables (rank <= 15)
  bool allow_wide = true;                            // FH_R64_ZT=3 / 4: modes 2 / 1 with float32 rows at every rank (A/B runs)
  if (R.zt_mode == 3) { R.zt_mode = 2; allow_wide = false; } else if (R.zt_mode == 4) { R.zt_mode = 1; allow_wide = false; }
  R.zt_wide = (allow_wide && R.zt_mode != 0 && !R.apply_only && P.m > 0 && 2 * P.m <= cap_tabs && 2 * P.m <= FH_RES_MAXM) ? 1 : 0;
  const size_t smem = sizeof(FhRes64Smem) + (tabs ? sizeof(FhRes64Tabs) : 0) + (size_t)R.gen.gbuf_bytes +
                      (size_t)(R.zt_wide ? 2 * P.m : std::min(P.m, cap)) * slice;
#ifdef FH_EMU
  (void)st;
  if (tabs) emu::launch_coop([&] { fh_k_cg_res64<false, true>(P, R); }, grid, block, smem);
  else if (big) emu::launch_coop([&] { fh_k_cg_res64<true, false>(P, R); }, grid, block, smem);
  else emu::launch_coop([&] { fh_k_cg_res64<false, false>(P, R); }, grid, block, smem);
  ++fh_g_launches;
#else
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (fh_g_prof_on) fh_prof_begin("fh_k_cg_res64", st);
  if (tabs) FH_CUDA(cudaLaunchKernelEx(&cfg, fh_k_cg_res64<false, true>, P, R));
  else if (big) FH_CUDA(cudaLaunchKernelEx(&cfg, fh_k_cg_res64<true, false>, P, R));
  else FH_CUDA(cudaLaunchKernelEx(&cfg, fh_k_cg_res64<false, false>, P, R));
  if (fh_g_prof_on) fh_prof_end(st);
  ++fh_g_launches;
#endif
  return FH_OK;
}

// out = A in (transpose = 0) or A^T in, float64, for every operator kind (once per solve: right-hand side and result)
static int op_apply64(fh_op* o, const double* in, double* out, int batch, int transpose, void* stream) {
  fh_ctx* c = o->ctx;
  const int planes = batch * c->C;
  if (o->kind == FH_OP_INPAINT) FH_FAIL(FH_EINVAL, "op_apply64: the mask is applied by the caller");
  FhTaps t{o->t_dy, o->t_dx, o->t_w, o->nt, o->ymin, o->ymax, o->xmin, o->xmax};
  if (o->kind == FH_OP_BLUR && o->separable && c->H % 32 == 0 && c->W % 32 == 0 &&
      o->nr <= FH_MAX_SEP_TAPS && o->nc <= FH_MAX_SEP_TAPS &&
      (size_t)((32 + o->nc - 1) * (((32 + o->nr - 1) | 1) + 33)) * sizeof(double) <= 46 * 1024) {
    FhSepTaps k{o->kr, o->kc, o->nr, o->nc, o->cr, o->cc};
    const int th = 32 + o->nc - 1, tw = (32 + o->nr - 1) | 1;
    const size_t smem = (size_t)(th * tw + th * 33) * sizeof(double);
    FH_LAUNCH(fh_k_conv_sep64, dim3(c->W / 32, c->H / 32, planes), dim3(32, 8), smem, stream, in, out, k, c->H, c->W, transpose);
  } else if (o->kind == FH_OP_BLUR) {
    if (c->H % 32 || c->W % 32) FH_FAIL(FH_EINVAL, "blur needs image sizes divisible by 32");
    const int th = 32 + (o->ymax - o->ymin), tw = (32 + (o->xmax - o->xmin)) | 1;
    const size_t smem = (size_t)th * tw * sizeof(double);
    if (smem > 42 * 1024) FH_FAIL(FH_EINVAL, "PSF support too large for the float64 convolution tile");   // + 6 KB of staged taps
    FH_LAUNCH(fh_k_conv_taps64, dim3(c->W / 32, c->H / 32, planes), dim3(32, 8), smem, stream, in, out, t, c->H, c->W, transpose);
  } else if (!transpose) {
    const int n = (c->H / o->sf) * (c->W / o->sf);
    FH_LAUNCH(fh_k_sr_forward64, dim3((n + 7) / 8, planes), dim3(256), 0, stream, in, out, t, c->H, c->W, o->sf);   // a warp per pixel
  } else {
    const int n = c->H * c->W;
    if (o->ymax - o->ymin + 1 > FH_SR_PSF_MAX || o->xmax - o->xmin + 1 > FH_SR_PSF_MAX) FH_FAIL(FH_EINVAL, "PSF support too large for the float64 super-resolution transpose");
    FH_LAUNCH(fh_k_sr_transpose64, dim3((n + 255) / 256, planes), dim3(256), 0, stream, in, out, t, c->H, c->W, o->sf);
  }
  FH_KCHECK();
  return FH_OK;
}

// Whole solve in float64: b = y - A mu (float64 convolution), one cooperative launch per image, mat = u or A^T u (float64).
// y: float64 (B, n); mu: float64 (B, d); mat_out: float64 (B, d).  info_dev != null: nothing is read back.
static int cg_solve_res64(fh_state* s, fh_op* o, FhCgFused P, const double* y, const double* mu, double sigma_y2, double rtol,
                          int last, int use_stop, double* mat_out, int* niter_host, int* optimal_host, double* resid_host,
                          void* stream, FhCgInfo* info_dev) {
  fh_ctx* c = s->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  const int B = s->B;
  const long n = op_ylen(o);                             // measurement space: d, or d / 16 for x4 super-resolution
  const long hw = (long)c->H * c->W;
  double *bv = c->d64[0], *x = c->d64[1], *amu = c->d64[2];
  int rc;
  const double* a = mu;
  if (o->kind != FH_OP_INPAINT) {
    if ((rc = op_apply64(o, mu, amu, B, 0, stream))) return rc;
    a = amu;
  }
  FH_LAUNCH(fh_k_rhs64, dim3((unsigned)((n + 255) / 256), B), dim3(256), 0, stream, y, a, P.mask, P.mask_batch, hw, bv, n);
  FH_KCHECK();
  FH_CUDA(cudaMemsetAsync(c->cg_sc, 0, (size_t)B * sizeof(FhCgScalars), st));
  P.strips_per_image = c->C * FH_R64_CPP;
  P.ystrips = P.strips_per_image;
  P.last_allowed = last;
  P.mirror = nullptr;
  const FhR64Gen gen = res64_gen(o, s->m, c->r64_smem_max - (int)sizeof(FhRes64Smem) - (int)sizeof(FhRes64Tabs));
  double* xout = (o->kind == FH_OP_INPAINT) ? mat_out : x;
  for (int b = 0; b < B; ++b) {
    FhCgRes64 R{};
    R.gen = gen;
    R.bvec = bv + (size_t)b * n; R.x_out = xout + (size_t)b * n;
    R.rtol = rtol; R.atol = 0.0; R.sy2 = sigma_y2; R.use_stop = use_stop; R.image = b; R.niter_hint = s->hint + b;
    R.info_out = info_dev ? info_dev + b : nullptr;
    if ((rc = launch_res64(c, P, R, st))) return rc;
  }
  if (o->kind != FH_OP_INPAINT && (rc = op_apply64(o, x, mat_out, B, 1, stream))) return rc;
  if (info_dev) return FH_OK;
  FH_CUDA(cudaMemcpyAsync(c->cg_sc_host, c->cg_sc, (size_t)B * sizeof(FhCgScalars), cudaMemcpyDeviceToHost, st));
  FH_CUDA(cudaStreamSynchronize(st));
  for (int b = 0; b < B; ++b) {
    if (!c->cg_sc_host[b].done) FH_FAIL(FH_ESTATE, "resident float64 CG kernel did not finish image %d", b);
    if (niter_host) niter_host[b] = c->cg_sc_host[b].niter;
    if (optimal_host) optimal_host[b] = c->cg_sc_host[b].optimal;
    if (resid_host) resid_host[b] = sqrt(c->cg_sc_host[b].resid2);
  }
  return FH_OK;
}

// Sigma v in float64 (denoiser_cov_vector_dot, online_update_bfgs.py:370-374, for a float64 v): one pass of the resident
// float64 kernel without operator -- 2-D DCT, delta (w_hat + Z^T Gs Z delta w_hat), inverse DCT, all float64.  The mechanism
// needs it because mu + Sigma mat cancels to ~1 % of its terms at high sigma (a float32 product costs two digits there).
static int cov_dot_f64(fh_state* s, const double* v, double* out, const int* only_if, void* stream);
extern "C" int fh_cov_dot_f64(fh_state* s, const double* v, double* out, void* stream) { return cov_dot_f64(s, v, out, nullptr, stream); }
extern "C" int fh_cov_dot_f64_if(fh_state* s, const double* v, double* out, const int* flag_dev, void* stream) {
  if (!flag_dev) FH_FAIL(FH_EINVAL, "fh_cov_dot_f64_if needs the device flags");
  return cov_dot_f64(s, v, out, flag_dev, stream);
}
static int cov_dot_f64(fh_state* s, const double* v, double* out, const int* only_if, void* stream) {
  if (!s || !v || !out) FH_FAIL(FH_EINVAL, "bad fh_cov_dot_f64 arguments");
  fh_ctx* c = s->ctx;
  if (s->basis != FH_BASIS_DCT || !c->r64_ok) FH_FAIL(FH_ESTATE, "fh_cov_dot_f64 needs the DCT basis and the resident float64 kernel");
  cudaStream_t st = (cudaStream_t)stream;
  FhCgFused P = fused_params(s, nullptr, 8);
  P.sc = c->cg_sc; P.sy2 = 0.f;
  P.strips_per_image = c->C * FH_R64_CPP;
  P.ystrips = P.strips_per_image;
  P.last_allowed = 1;
  P.mirror = nullptr;
  FH_CUDA(cudaMemsetAsync(c->cg_sc, 0, (size_t)s->B * sizeof(FhCgScalars), st));
  for (int b = 0; b < s->B; ++b) {
    FhCgRes64 R{};
    R.bvec = v + (size_t)b * s->d; R.x_out = out + (size_t)b * s->d;
    R.rtol = 1.0; R.atol = 0.0; R.sy2 = 0.0; R.use_stop = 0; R.image = b; R.apply_only = 1;
    R.only_if = only_if;
    int rc = launch_res64(c, P, R, st);
    if (rc) return rc;
  }
  return FH_OK;
}

// float32 entry points over the float64 solve: widen mu, narrow mat (both exact / one rounding)
static int cg_solve_res64_f32io(fh_state* s, fh_op* o, const FhCgFused& P, const float* y, const float* mu, double sigma_y2,
                                double rtol, int last, int use_stop, float* mat_out, int* niter_host, int* optimal_host,
                                double* resid_host, void* stream, FhCgInfo* info_dev) {
  fh_ctx* c = s->ctx;
  const long tot = (long)s->B * s->d;
  // d64[2] doubles as the widened mu for inpainting (no A mu needed there); blur: A mu overwrites it only after it was read
  double* mu64 = c->d64_out;
  FH_LAUNCH(fh_k_f2d, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, stream, mu, mu64, tot);
  const long ytot = (long)s->B * op_ylen(o);
  FH_LAUNCH(fh_k_f2d, dim3((unsigned)((ytot + 255) / 256)), dim3(256), 0, stream, y, c->d64_y, ytot);
  FH_KCHECK();
  // the result may overwrite the widened mu: mu is consumed (rhs) before any image's solve starts
  int rc = cg_solve_res64(s, o, P, c->d64_y, mu64, sigma_y2, rtol, last, use_stop, c->d64_out, niter_host, optimal_host, resid_host, stream, info_dev);
  if (rc) return rc;
  FH_LAUNCH(fh_k_d2f, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, stream, (const double*)c->d64_out, mat_out, tot);
  FH_KCHECK();
  return FH_OK;
}

// DCT-basis solve: the iteration of fh_cg.cuh.  5 launches per iteration when the operator folds into the DCT passes
// (mask, separable blur), 9 otherwise (p update, A^T, 4 passes, A, q, x/r update).
static int cg_solve_fused(fh_state* s, fh_op* o, const float* y, const float* mu, double sigma_y2, double rtol,
                          int maxiter, int fixed_iters, float* mat_out, int* niter_host, int* optimal_host,
                          double* resid_host, void* stream) {
  fh_ctx* c = s->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  const long n = (long)c->C * (c->H / o->sf) * (c->W / o->sf);
  const int B = s->B;
  const int R = pick_rows(c, B * c->C);
  const int selems = R * FH_N;
  if (n % selems) FH_FAIL(FH_EINVAL, "measurement size %ld is not a multiple of %d", n, selems);
  const int ystrips = (int)(n / selems);
  const long hw = (long)(c->H / o->sf) * (c->W / o->sf);
  FhCgFused P = fused_params(s, o, R);
  P.sc = c->cg_sc; P.sy2 = (float)sigma_y2; P.ystrips = ystrips;
  const int use_stop = fixed_iters > 0 ? 0 : 1;
  const int last = fixed_iters > 0 ? std::min(fixed_iters, maxiter) : maxiter;
  P.last_allowed = last;
  if (resident_applies(s, P, o)) {
    if (resident_is_f64(c)) {
      c->cg_last_mode = FH_CG_RESIDENT;
      return cg_solve_res64_f32io(s, o, P, y, mu, sigma_y2, rtol, last, use_stop, mat_out, niter_host, optimal_host, resid_host, stream, nullptr);
    }
    c->cg_last_mode = FH_CG_RESIDENT_F32;
    return cg_solve_resident(s, o, P, y, mu, rtol, last, use_stop, mat_out, niter_host, optimal_host, resid_host, stream);
  }
  if (c->cg_mode == FH_CG_RESIDENT || c->cg_mode == FH_CG_RESIDENT_F32)
    FH_FAIL(FH_EINVAL, "resident CG solve requested but not applicable (operator does not fold into the DCT passes, or the device cannot hold the cooperative grid)");
  c->cg_last_mode = FH_CG_STREAMING;
  const bool fusedop = P.fuse != 0;
  const dim3 block(FH_PASS_THREADS);
  const dim3 gimg = strip_grid(c, B * P.strips_per_image), gy = strip_grid(c, B * ystrips);
  const size_t zbytes = (size_t)P.zstages * FH_ZROUND_FLOATS * 4;
  float *bv = c->cgv[0], *x = c->cgv[1], *t = c->cgv[5];
  float *img = c->scratch[0], *img2 = c->scratch[1];
  int rc;
  // rhs and x0 = b
  const float* amu = mu;
  if (o->kind != FH_OP_INPAINT) {
    if ((rc = fh_op_apply(o, mu, t, B, 0, stream))) return rc;
    amu = t;
  }
  FH_LAUNCH(fh_k_cg_rhs, dim3((unsigned)((n + 255) / 256), B), dim3(256), 0, stream, y, amu, P.mask, P.mask_batch, hw, bv, x, n);
  FH_KCHECK();
  FH_CUDA(cudaMemsetAsync(c->cg_sc, 0, (size_t)B * sizeof(FhCgScalars), st));
  // one application of the system operator: it == 0 on x (initial residual), it >= 1 on p (updated first) -> q, p.q partials
  auto apply = [&](int it) -> int {
    int rc2 = FH_OK;
    const float* vec = it == 0 ? x : nullptr;
    FH_ROWS_SWITCH(R, {
      if (fusedop) {
        FH_LAUNCH_PDL(fh_k_cg_fwd_rows<RR>, gimg, block, 0, stream, P, (const float*)nullptr, vec, it);
        FH_LAUNCH_PDL(fh_k_cg_fwd_cols<RR>, gimg, block, zbytes, stream, P);
        FH_LAUNCH_PDL(fh_k_cg_inv_cols<RR>, gimg, block, zbytes, stream, P);
        FH_LAUNCH_PDL(fh_k_cg_inv_rows<RR>, gimg, block, 0, stream, P, (float*)nullptr, vec, it);
      } else {
        if (it >= 1) FH_LAUNCH_PDL(fh_k_cg_p<RR>, gy, block, 0, stream, P, it);
        const float* pv = vec ? vec : P.p;
        if ((rc2 = fh_op_apply(o, pv, img, B, 1, stream))) return rc2;
        FH_LAUNCH_PDL(fh_k_cg_fwd_rows<RR>, gimg, block, 0, stream, P, (const float*)img, (const float*)nullptr, 0);
        FH_LAUNCH_PDL(fh_k_cg_fwd_cols<RR>, gimg, block, zbytes, stream, P);
        FH_LAUNCH_PDL(fh_k_cg_inv_cols<RR>, gimg, block, zbytes, stream, P);
        FH_LAUNCH_PDL(fh_k_cg_inv_rows<RR>, gimg, block, 0, stream, P, img2, (const float*)nullptr, 0);
        if ((rc2 = fh_op_apply(o, img2, t, B, 0, stream))) return rc2;
        FH_LAUNCH_PDL(fh_k_cg_q2<RR>, gy, block, 0, stream, P, (const float*)t, vec, it);
      }
    });
    FH_KCHECK();
    return rc2;
  };
  if ((rc = apply(0))) return rc;                                  // q = A_op x0
  FH_ROWS_SWITCH(R, { FH_LAUNCH(fh_k_cg_r0<RR>, gy, block, 0, stream, P, (const float*)bv); });   // r0 = b - q
  FH_LAUNCH(fh_k_cg_r0_scalars, dim3(B), block, 0, stream, P, rtol, 0.0, use_stop);
  FH_KCHECK();
  // The host never synchronises inside the loop.  The kernels report (progress, done) per image into pinned mapped
  // memory; iteration `it` is launched once every running image has executed iteration it - FH_RUNAHEAD, so the GPU
  // always has work queued and at most FH_RUNAHEAD (+1: the stop test of iteration k runs in K1 of k+1) no-op
  // iterations are launched after the last image has stopped.  The reference syncs twice per iteration (cg.py:250,259).
  constexpr int FH_RUNAHEAD = 2;
  for (int b = 0; b < B; ++b) { c->mirror_host[2 * b] = 0; c->mirror_host[2 * b + 1] = 0; }
  P.mirror = c->mirror_dev;
  int it_done = 0;
  for (int it = 1; it <= last; ++it) {
    bool all = false;
    for (long spins = 0;; ++spins) {
      all = true;
      bool ready = true;
      for (int b = 0; b < B; ++b) {
        const bool dn = c->mirror_host[2 * b + 1] != 0;
        all = all && dn;
        ready = ready && (dn || c->mirror_host[2 * b] >= it - FH_RUNAHEAD);
      }
      if (all || ready) break;
      fh_host_relax(spins);                                         // pause; yield the core after a while (no hot spin)
      if (spins > 200000000L) {                                     // ~seconds: something is wrong, fall back to a hard sync
        FH_CUDA(cudaStreamSynchronize(st));
        break;
      }
    }
    if (all) break;
    if ((rc = apply(it))) return rc;                               // (stop test of it-1,) p update, q = A_op p
    FH_ROWS_SWITCH(R, { FH_LAUNCH_PDL(fh_k_cg_xr<RR>, gy, block, 0, stream, P, it); });
    FH_KCHECK();
    it_done = it;
  }
  FH_LAUNCH(fh_k_cg_finish, dim3(B), block, 0, stream, P, it_done + 1);   // records the stop test of the last iteration launched
  FH_KCHECK();
  FH_CUDA(cudaMemcpyAsync(c->cg_sc_host, c->cg_sc, (size_t)B * sizeof(FhCgScalars), cudaMemcpyDeviceToHost, st));
  FH_CUDA(cudaStreamSynchronize(st));
  for (int b = 0; b < B; ++b)
    if (!c->cg_sc_host[b].done) FH_FAIL(FH_ESTATE, "CG loop ended with image %d still running (iteration %d of %d)", b, it_done, last);
  for (int b = 0; b < B; ++b) {
    if (niter_host) niter_host[b] = c->cg_sc_host[b].niter;
    if (optimal_host) optimal_host[b] = c->cg_sc_host[b].optimal;
    if (resid_host) resid_host[b] = sqrt(c->cg_sc_host[b].resid2);
  }
  if (o->kind == FH_OP_INPAINT) {
    FH_CUDA(cudaMemcpyAsync(mat_out, x, (size_t)B * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return FH_OK;
  }
  return fh_op_apply(o, x, mat_out, B, 1, stream);
}

// Fully asynchronous solve: nothing is read back, the call returns as soon as the work is queued on `stream`.  Only
// where the resident kernel applies (the streaming path needs the host to follow the iteration count).
extern "C" int fh_cg_solve_async(fh_state* s, fh_op* o, const float* y, const float* mu, double sigma_y2, double rtol,
                                 int maxiter, int fixed_iters, float* mat_out, fh_cg_info* info_dev, void* stream) {
  if (!s || !o || !y || !mu || !mat_out || !info_dev || maxiter < 1) FH_FAIL(FH_EINVAL, "bad fh_cg_solve_async arguments");
  if (!(rtol > 0) && fixed_iters <= 0) FH_FAIL(FH_EINVAL, "rtol must be positive");
  if (o->kind == FH_OP_INPAINT && o->mask_batch != 1 && o->mask_batch != s->B) FH_FAIL(FH_EINVAL, "mask batch %d does not match the state's batch %d", o->mask_batch, s->B);
  if (s->basis != FH_BASIS_DCT) FH_FAIL(FH_ESTATE, "asynchronous solve needs the DCT basis");
  fh_ctx* c = s->ctx;
  FhCgFused P = fused_params(s, o, 8);
  if (!resident_applies(s, P, o)) FH_FAIL(FH_ESTATE, "asynchronous solve needs the resident kernel (rank %d, operator kind %d)", s->m, o->kind);
  P.sc = c->cg_sc; P.sy2 = (float)sigma_y2;
  const int use_stop = fixed_iters > 0 ? 0 : 1;
  const int last = fixed_iters > 0 ? std::min(fixed_iters, maxiter) : maxiter;
  static_assert(sizeof(FhCgInfo) == sizeof(fh_cg_info), "device record layout");
  if (resident_is_f64(c)) {
    c->cg_last_mode = FH_CG_RESIDENT;
    return cg_solve_res64_f32io(s, o, P, y, mu, sigma_y2, rtol, last, use_stop, mat_out, nullptr, nullptr, nullptr, stream, reinterpret_cast<FhCgInfo*>(info_dev));
  }
  c->cg_last_mode = FH_CG_RESIDENT_F32;
  return cg_solve_resident(s, o, P, y, mu, rtol, last, use_stop, mat_out, nullptr, nullptr, nullptr, stream, reinterpret_cast<FhCgInfo*>(info_dev));
}

// The guidance solve with float64 in and out (the reference's own precision): mu and mat are float64, y float32.
// Available where the float64 resident kernel applies; otherwise FH_ESTATE (use the float32 entry points).
static int solve_f64_common(fh_state* s, fh_op* o, const double* y, const double* mu, double sigma_y2, double rtol, int maxiter,
                            int fixed_iters, double* mat_out, int* niter_host, int* optimal_host, double* resid_host,
                            fh_cg_info* info_dev, void* stream) {
  if (!s || !o || !y || !mu || !mat_out || maxiter < 1) FH_FAIL(FH_EINVAL, "bad fh_cg_solve_f64 arguments");
  if (!(rtol > 0) && fixed_iters <= 0) FH_FAIL(FH_EINVAL, "rtol must be positive");
  if (s->basis != FH_BASIS_DCT) FH_FAIL(FH_ESTATE, "float64 solve needs the DCT basis");
  if (o->kind == FH_OP_INPAINT && o->mask_batch != 1 && o->mask_batch != s->B) FH_FAIL(FH_EINVAL, "mask batch %d does not match the state's batch %d", o->mask_batch, s->B);
  fh_ctx* c = s->ctx;
  FhCgFused P = fused_params(s, o, 8);
  if (!resident_applies(s, P, o) || !resident_is_f64(c))
    FH_FAIL(FH_ESTATE, "float64 solve needs the resident float64 kernel (rank %d, operator kind %d, cg mode %d)", s->m, o->kind, c->cg_mode);
  P.sc = c->cg_sc; P.sy2 = (float)sigma_y2;
  const int use_stop = fixed_iters > 0 ? 0 : 1;
  const int last = fixed_iters > 0 ? std::min(fixed_iters, maxiter) : maxiter;
  c->cg_last_mode = FH_CG_RESIDENT;
  return cg_solve_res64(s, o, P, y, mu, sigma_y2, rtol, last, use_stop, mat_out, niter_host, optimal_host, resid_host, stream,
                        reinterpret_cast<FhCgInfo*>(info_dev));
}
extern "C" int fh_cg_solve_f64(fh_state* s, fh_op* o, const double* y, const double* mu, double sigma_y2, double rtol, int maxiter,
                               int fixed_iters, double* mat_out, int* niter_host, int* optimal_host, double* resid_host, void* stream) {
  return solve_f64_common(s, o, y, mu, sigma_y2, rtol, maxiter, fixed_iters, mat_out, niter_host, optimal_host, resid_host, nullptr, stream);
}
extern "C" int fh_cg_solve_f64_async(fh_state* s, fh_op* o, const double* y, const double* mu, double sigma_y2, double rtol, int maxiter,
                                     int fixed_iters, double* mat_out, fh_cg_info* info_dev, void* stream) {
  if (!info_dev) FH_FAIL(FH_EINVAL, "fh_cg_solve_f64_async needs info_dev");
  return solve_f64_common(s, o, y, mu, sigma_y2, rtol, maxiter, fixed_iters, mat_out, nullptr, nullptr, nullptr, info_dev, stream);
}

extern "C" int fh_cg_solve(fh_state* s, fh_op* o, const float* y, const float* mu, double sigma_y2, double rtol,
                           int maxiter, int fixed_iters, float* mat_out, int* niter_host, int* optimal_host,
                           double* resid_host, void* stream) {
  if (!s || !o || !y || !mu || !mat_out || maxiter < 1) FH_FAIL(FH_EINVAL, "bad fh_cg_solve arguments");
  if (!(rtol > 0) && fixed_iters <= 0) FH_FAIL(FH_EINVAL, "rtol must be positive");
  if (o->kind == FH_OP_INPAINT && o->mask_batch != 1 && o->mask_batch != s->B) FH_FAIL(FH_EINVAL, "mask batch %d does not match the state's batch %d", o->mask_batch, s->B);
  if (s->basis == FH_BASIS_DCT)
    return cg_solve_fused(s, o, y, mu, sigma_y2, rtol, maxiter, fixed_iters, mat_out, niter_host, optimal_host, resid_host, stream);
  fh_ctx* c = s->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  const long n = op_ylen(o);
  const int B = s->B;
  const int nblk = fh_nchunks(n);
  const long hw = (long)(c->H / o->sf) * (c->W / o->sf);
  const float* mask = (o->kind == FH_OP_INPAINT) ? o->mask : nullptr;
  const int mb = o->mask_batch;
  float *bv = c->cgv[0], *x = c->cgv[1], *r = c->cgv[2], *p = c->cgv[3], *q = c->cgv[4], *t = c->cgv[5];
  int rc;
  // rhs and x0 = b
  const float* amu = mu;
  if (o->kind != FH_OP_INPAINT) {
    if ((rc = fh_op_apply(o, mu, t, B, 0, stream))) return rc;
    amu = t;
  }
  FH_LAUNCH(fh_k_cg_rhs, dim3((unsigned)((n + 255) / 256), B), dim3(256), 0, stream, y, amu, mask, mb, hw, bv, x, n);
  FH_KCHECK();
  // r0 = b - A_op x0, p0 = r0
  FH_CUDA(cudaMemsetAsync(c->cg_sc, 0, (size_t)B * sizeof(FhCgScalars), st));
  if ((rc = apply_ASAt(s, o, x, t, stream))) return rc;
  FH_LAUNCH(fh_k_cg_q, dim3(nblk, B), dim3(FH_EW_THREADS), 0, stream, x, t, mask, mb, hw, (float)sigma_y2, q, c->part_a, c->cg_sc, n);
  FH_KCHECK();
  FH_LAUNCH(fh_k_cg_init, dim3(nblk, B), dim3(FH_EW_THREADS), 0, stream, bv, q, r, p, c->part_a, c->part_b, n);
  FH_KCHECK();
  const int use_stop = fixed_iters > 0 ? 0 : 1;
  const int last = fixed_iters > 0 ? std::min(fixed_iters, maxiter) : maxiter;
  FH_LAUNCH(fh_k_cg_init_scalars, dim3((B + 63) / 64), dim3(64), 0, stream, c->cg_sc, c->part_a, c->part_b, nblk, rtol, 0.0, use_stop, B);
  FH_KCHECK();
  int poll = 4;
  int since = 0;
  for (int it = 1; it <= last; ++it) {
    if ((rc = apply_ASAt(s, o, p, t, stream))) return rc;
    FH_LAUNCH(fh_k_cg_q, dim3(nblk, B), dim3(FH_EW_THREADS), 0, stream, p, t, mask, mb, hw, (float)sigma_y2, q, c->part_a, c->cg_sc, n);
    FH_LAUNCH(fh_k_cg_update1, dim3(nblk, B), dim3(FH_EW_THREADS), 0, stream, x, r, p, q, c->part_a, c->part_b, c->cg_sc, it, n);
    FH_LAUNCH(fh_k_cg_update2, dim3(nblk, B), dim3(FH_EW_THREADS), 0, stream, r, p, c->part_b, c->cg_sc, it, last, n);
    FH_KCHECK();
    if (++since >= poll || it == last) {
      since = 0;
      FH_CUDA(cudaMemcpyAsync(c->cg_sc_host, c->cg_sc, (size_t)B * sizeof(FhCgScalars), cudaMemcpyDeviceToHost, st));
      FH_CUDA(cudaStreamSynchronize(st));
      bool all = true;
      for (int b = 0; b < B; ++b) all = all && c->cg_sc_host[b].done;
      if (all) break;
      if (poll < 16) poll *= 2;
    }
  }
  for (int b = 0; b < B; ++b) {
    if (niter_host) niter_host[b] = c->cg_sc_host[b].niter;
    if (optimal_host) optimal_host[b] = c->cg_sc_host[b].optimal;
    if (resid_host) resid_host[b] = sqrt(c->cg_sc_host[b].resid2);
  }
  // mat = u (inpainting) or A^T u
  if (o->kind == FH_OP_INPAINT) {
    FH_CUDA(cudaMemcpyAsync(mat_out, x, (size_t)B * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return FH_OK;
  }
  return fh_op_apply(o, x, mat_out, B, 1, stream);
}

// bench.py's roofline: `reps` back-to-back launches of ONE kernel of the fused iteration at the state's current
// rank, launched exactly as in the product (programmatic dependent launch), timed with CUDA events on `stream` (inputs are whatever the workspace holds; sizes and access pattern are
// those of a real iteration).  which: 0 fwd_rows, 1 fwd_cols, 2 inv_cols, 3 inv_rows, 4 xr.
extern "C" int fh_cg_kernel_time(fh_state* s, fh_op* o, int which, int reps, float* ms_out, void* stream) {
#ifdef FH_EMU
  (void)s; (void)o; (void)which; (void)reps; (void)ms_out; (void)stream;
  FH_FAIL(FH_EINVAL, "not available in the emulation build");
#else
  if (!s || !o || !ms_out || reps < 1 || which < 0 || which > 4) FH_FAIL(FH_EINVAL, "bad fh_cg_kernel_time arguments");
  if (s->basis != FH_BASIS_DCT) FH_FAIL(FH_EINVAL, "fused kernels need the DCT basis");
  fh_ctx* c = s->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  const int R = pick_rows(c, s->B * c->C);
  FhCgFused P = fused_params(s, o, R);
  if (P.fuse == 0) FH_FAIL(FH_EINVAL, "operator does not fold into the DCT passes");
  P.sc = c->cg_sc; P.sy2 = 0.01f; P.last_allowed = 1 << 30;
  // scalars that keep every kernel on its normal path: not done, no stop test, finite alpha / beta
  std::vector<FhCgScalars> sc(s->B);
  for (auto& v : sc) { memset(&v, 0, sizeof(v)); v.rr[0] = v.rr[1] = 1.0; v.stop2 = -1.0; }
  FH_CUDA(cudaMemcpyAsync(c->cg_sc, sc.data(), sc.size() * sizeof(FhCgScalars), cudaMemcpyHostToDevice, st));
  std::vector<double> ones((size_t)s->B * P.ystrips, 1.0);
  FH_CUDA(cudaMemcpyAsync(c->part_a, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  FH_CUDA(cudaMemcpyAsync(c->part_b, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  FH_CUDA(cudaStreamSynchronize(st));
  const dim3 block(FH_PASS_THREADS), gimg = strip_grid(c, s->B * P.strips_per_image);
  const size_t zbytes = (size_t)P.zstages * FH_ZROUND_FLOATS * 4;
  cudaEvent_t e0, e1;
  FH_CUDA(cudaEventCreate(&e0)); FH_CUDA(cudaEventCreate(&e1));
  for (int pass = 0; pass < 2; ++pass) {                 // pass 0 warms up
    if (pass) FH_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < (pass ? reps : 3); ++i) {
      FH_ROWS_SWITCH(R, {
        if (which == 0) FH_LAUNCH_PDL(fh_k_cg_fwd_rows<RR>, gimg, block, 0, stream, P, (const float*)nullptr, (const float*)nullptr, 2);
        else if (which == 1) FH_LAUNCH_PDL(fh_k_cg_fwd_cols<RR>, gimg, block, zbytes, stream, P);
        else if (which == 2) FH_LAUNCH_PDL(fh_k_cg_inv_cols<RR>, gimg, block, zbytes, stream, P);
        else if (which == 3) FH_LAUNCH_PDL(fh_k_cg_inv_rows<RR>, gimg, block, 0, stream, P, (float*)nullptr, (const float*)nullptr, 2);
        else FH_LAUNCH_PDL(fh_k_cg_xr<RR>, gimg, block, 0, stream, P, 2);
      });
    }
    if (pass) FH_CUDA(cudaEventRecord(e1, st));
  }
  FH_KCHECK();
  FH_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  FH_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_out = ms / reps;
  return FH_OK;
#endif
}

// bench.py's roofline for the resident solve: `reps` launches of exactly `iters` CG iterations (+ the initial residual)
// on image 0, at the state's current rank, right-hand side = whatever the last solve left in the workspace; average
// device time per LAUNCH (ms) between two CUDA events on `stream`.
extern "C" int fh_cg_resident_time(fh_state* s, fh_op* o, int iters, int reps, float* ms_out, void* stream) {
#ifdef FH_EMU
  (void)s; (void)o; (void)iters; (void)reps; (void)ms_out; (void)stream;
  FH_FAIL(FH_EINVAL, "not available in the emulation build");
#else
  if (!s || !o || !ms_out || reps < 1 || iters < 1) FH_FAIL(FH_EINVAL, "bad fh_cg_resident_time arguments");
  if (s->basis != FH_BASIS_DCT) FH_FAIL(FH_EINVAL, "resident kernel needs the DCT basis");
  fh_ctx* c = s->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  FhCgFused P = fused_params(s, o, 8);
  const int saved_mode = c->cg_mode;
  c->cg_mode = FH_CG_AUTO;
  const bool ok = resident_applies(s, P, o);
  c->cg_mode = saved_mode;
  if (!ok) FH_FAIL(FH_EINVAL, "resident kernel does not apply to this state / operator");
  P.sc = c->cg_sc; P.sy2 = 0.01f; P.last_allowed = iters;
  P.strips_per_image = (int)(s->d / FH_RES_SLICE); P.ystrips = P.strips_per_image;
  const bool f64 = resident_is_f64(c);                  // the kernel fh_cg_solve would run in the current mode
  FhCgResident R{};
  R.bvec = c->cgv[0]; R.x_out = c->cgv[5]; R.rtol = 1.0; R.atol = 0.0; R.use_stop = 0; R.image = 0;
  R.timing = 1;                         // a converged workspace must not end the run early (the arithmetic is the same)
  FhCgRes64 R64{};
  R64.bvec = c->d64[0]; R64.x_out = c->d64[1]; R64.rtol = 1.0; R64.atol = 0.0; R64.sy2 = 0.01; R64.use_stop = 0; R64.image = 0; R64.niter_hint = s->hint;
  R64.timing = 1;
  R64.gen = res64_gen(o, s->m, c->r64_smem_max - (int)sizeof(FhRes64Smem) - (int)sizeof(FhRes64Tabs));
  // its own right-hand side (the workspace may hold anything, e.g. a converged solve's b), in the measurement space
  {
    const long nb = s->d;
    if (f64) FH_LAUNCH(fh_k_fill_pattern<double>, dim3((unsigned)((nb + 255) / 256)), dim3(256), 0, stream, c->d64[0], nb);
    else FH_LAUNCH(fh_k_fill_pattern<float>, dim3((unsigned)((nb + 255) / 256)), dim3(256), 0, stream, c->cgv[0], nb);
    FH_KCHECK();
  }
  cudaEvent_t e0, e1;
  FH_CUDA(cudaEventCreate(&e0)); FH_CUDA(cudaEventCreate(&e1));
  int rc = FH_OK;
  for (int pass = 0; pass < 2 && rc == FH_OK; ++pass) {              // pass 0 warms up
    if (pass) FH_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < (pass ? reps : 1) && rc == FH_OK; ++i) {
      FH_CUDA(cudaMemsetAsync(c->cg_sc, 0, sizeof(FhCgScalars), st));
      rc = f64 ? launch_res64(c, P, R64, st) : launch_resident(c, P, R, st);
    }
    if (pass) FH_CUDA(cudaEventRecord(e1, st));
  }
  if (rc) return rc;
  FH_KCHECK();
  FH_CUDA(cudaMemcpyAsync(c->cg_sc_host, c->cg_sc, sizeof(FhCgScalars), cudaMemcpyDeviceToHost, st));
  FH_CUDA(cudaEventSynchronize(e1));
  FH_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  FH_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (c->cg_sc_host[0].niter != iters)
    FH_FAIL(FH_ESTATE, "timing run broke off after %d of %d iterations (workspace holds no usable right-hand side)", c->cg_sc_host[0].niter, iters);
  *ms_out = ms / reps;
  return FH_OK;
#endif
}

// ------------------------------------------------------------------------------------ fused mechanism / sampler glue
extern "C" int fh_mech_combine(fh_ctx* c, const double* mu, const double* g, const double* sigv, double sigma2, double threshold,
                               double scale, double* out, int* flag_dev, int batch, void* stream) {
  if (!c || !mu || !g || !sigv || !out || !flag_dev || batch < 1 || batch > c->max_batch || !(sigma2 > 0)) FH_FAIL(FH_EINVAL, "bad fh_mech_combine arguments");
  const int nblk = (int)((c->d + FH_MECH_CHUNK - 1) / FH_MECH_CHUNK);
  // the Gram partials buffer is free between covariance updates: (B, nblk, 3) doubles fit easily
  double* part = c->gram_partial;
  FH_LAUNCH(fh_k_mech_stats, dim3(nblk, batch), dim3(FH_MECH_THREADS), 0, stream, g, sigma2, part, c->d);
  FH_LAUNCH(fh_k_mech_combine, dim3(nblk, batch), dim3(FH_MECH_THREADS), 0, stream, mu, g, sigv, (const double*)part, nblk, sigma2, threshold, scale, out, flag_dev, c->d, (const int*)nullptr);
  FH_KCHECK();
  return FH_OK;
}

// The same in two steps, so that Sigma.mat (fh_cov_dot_f64_if) runs only for the images that take that branch -- like the
// reference, which forms it inside the `if` (conditioning_mechanisms.py:283-286) -- without a host read of the decision.
extern "C" int fh_mech_flag(fh_ctx* c, const double* g, double sigma2, double threshold, int* flag_dev, int batch, void* stream) {
  if (!c || !g || !flag_dev || batch < 1 || batch > c->max_batch || !(sigma2 > 0)) FH_FAIL(FH_EINVAL, "bad fh_mech_flag arguments");
  const int nblk = (int)((c->d + FH_MECH_CHUNK - 1) / FH_MECH_CHUNK);
  double* part = c->gram_partial;
  FH_LAUNCH(fh_k_mech_stats, dim3(nblk, batch), dim3(FH_MECH_THREADS), 0, stream, g, sigma2, part, c->d);
  FH_LAUNCH(fh_k_mech_flag, dim3(batch), dim3(FH_MECH_THREADS), 0, stream, (const double*)part, nblk, threshold, flag_dev);
  FH_KCHECK();
  return FH_OK;
}
extern "C" int fh_mech_apply(fh_ctx* c, const double* mu, const double* g, const double* sigv, double sigma2, double scale,
                             const int* flag_dev, double* out, int batch, void* stream) {
  if (!c || !mu || !g || !sigv || !out || !flag_dev || batch < 1 || batch > c->max_batch || !(sigma2 > 0)) FH_FAIL(FH_EINVAL, "bad fh_mech_apply arguments");
  const int nblk = (int)((c->d + FH_MECH_CHUNK - 1) / FH_MECH_CHUNK);
  FH_LAUNCH(fh_k_mech_combine, dim3(nblk, batch), dim3(FH_MECH_THREADS), 0, stream, mu, g, sigv, (const double*)nullptr, nblk, sigma2, 0.0, scale, out, (int*)nullptr, c->d, flag_dev);
  FH_KCHECK();
  return FH_OK;
}
extern "C" int fh_mech_allclose(fh_ctx* c, const double* a, const double* b, int* notclose_dev, int batch, void* stream) {
  if (!c || !a || !b || !notclose_dev || batch < 1) FH_FAIL(FH_EINVAL, "bad fh_mech_allclose arguments");
  FH_CUDA(cudaMemsetAsync(notclose_dev, 0, (size_t)batch * sizeof(int), (cudaStream_t)stream));
  const int nblk = (int)((c->d + FH_MECH_CHUNK - 1) / FH_MECH_CHUNK);
  FH_LAUNCH(fh_k_mech_allclose, dim3(nblk, batch), dim3(FH_MECH_THREADS), 0, stream, a, b, notclose_dev, c->d);
  FH_KCHECK();
  return FH_OK;
}
// update_time_step fed with the previous evaluation itself: score_prev = (mean_prev - x_prev) / sigma_t^2 (conditioning_mechanisms.py:243)
extern "C" int fh_time_update_from_mean(fh_state* s, const double* x_prev, const double* mean_prev, double sigma_t, double sigma_tnext,
                                        double* mean_out, double* score_out, void* stream) {
  if (!s || !x_prev || !mean_prev || !mean_out || !score_out || !(sigma_t > 0) || !(sigma_tnext > 0)) FH_FAIL(FH_EINVAL, "bad fh_time_update_from_mean arguments");
  if (upd64_applies(s)) return time_update_fused(s, x_prev, mean_prev, nullptr, sigma_t, sigma_tnext, mean_out, score_out, stream);
  fh_ctx* c = s->ctx;
  const long tot = (long)s->B * s->d;
  const dim3 g((unsigned)((tot + 255) / 256)), b(256);
  FH_LAUNCH(fh_k_mech_score, g, b, 0, stream, mean_prev, x_prev, sigma_t * sigma_t, c->cgv[0], c->cgv[1], tot);
  FH_KCHECK();
  int rc = fh_time_update(s, c->cgv[0], c->cgv[1], sigma_t, sigma_tnext, 0, c->cgv[2], c->cgv[3], stream);
  if (rc) return rc;
  FH_LAUNCH(fh_k_tu_finish64, g, b, 0, stream, x_prev, (const float*)c->cgv[3], sigma_tnext * sigma_tnext, mean_out, score_out, tot);
  FH_KCHECK();
  return FH_OK;
}
extern "C" int fh_sampler_euler(const double* x_hat, const double* den, double t_cur, double h, double* d_cur, double* x_prime, long n, void* stream) {
  if (!x_hat || !den || !d_cur || !x_prime || n < 1) FH_FAIL(FH_EINVAL, "bad fh_sampler_euler arguments");
  FH_LAUNCH(fh_k_euler, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, x_hat, den, t_cur, h, d_cur, x_prime, n);
  FH_KCHECK();
  return FH_OK;
}
extern "C" int fh_sampler_heun(const double* x_hat, const double* d_cur, const double* x_prime, const double* den2, double t_next, double h,
                               double* x_next, long n, void* stream) {
  if (!x_hat || !d_cur || !x_prime || !den2 || !x_next || n < 1) FH_FAIL(FH_EINVAL, "bad fh_sampler_heun arguments");
  FH_LAUNCH(fh_k_heun, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, x_hat, d_cur, x_prime, den2, t_next, h, x_next, n);
  FH_KCHECK();
  return FH_OK;
}
