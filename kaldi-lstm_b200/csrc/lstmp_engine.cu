// Host side of the C ABI (include/lstmp_b200.h): owns parameters, momentum buffers, carried
// state and the per-chunk activation record in HBM; decides the work decomposition; launches
// the sm_100a kernels.  No CPU compute path exists here by design.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/lstmp_b200.h"
#include "lstmp_kernels.h"

using namespace lstmp;

namespace lstmp {
cudaError_t launch_gemm_simt(float* C, long long ldc, int M, int N, int K, float alpha, const float* A,
                             long long lda, int tA, const float* B, long long ldb, int tB, float beta,
                             const float* bias, cudaStream_t stream);
#ifdef LSTMP_HAVE_TC_GEMM
cudaError_t launch_gemm_tc(float* C, long long ldc, int M, int N, int K, float alpha, const float* A,
                           long long lda, int tA, const float* B, long long ldb, int tB, float beta,
                           const float* bias, cudaStream_t stream, bool* handled, float* ws, size_t ws_floats,
                           int* nlaunch);
#endif
}  // namespace lstmp

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// Every entry point runs on the handle's device and restores the caller's current device on return (a process may
// hold engines on several GPUs; torch's notion of the current device must not change under it).
struct DeviceGuard {
  int prev = -1, dev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int d) : dev(d) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != d) err = cudaSetDevice(d);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};
struct lstmp_b200_engine {
  int I = 0, C = 0, R = 0, S = 0, Tmax = 0, device = 0, sm_count = 0;
  size_t nparams = 0;
  // arenas, GetParams order (LPS.h:162-189)
  float *params = nullptr, *corr = nullptr, *grads = nullptr;
  size_t off_wx = 0, off_wr = 0, off_bias = 0, off_pi = 0, off_pf = 0, off_po = 0, off_wm = 0;
  // carried state (c and r blocks of prev_nnet_state_, LPS.h:583)
  float *state_c = nullptr, *state_r = nullptr;
  // per-chunk record
  float *gifo = nullptr, *cbuf = nullptr, *hbuf = nullptr, *mbuf = nullptr, *rbuf = nullptr;
  float *dgifo = nullptr, *dr = nullptr, *scratch = nullptr, *small_grads = nullptr;
  unsigned* bar = nullptr;
  unsigned bar_base[kMaxGroupsHost] = {0};
  size_t workspace_bytes = 0;
  Decomp d{};
  FwdParams fp{};
  BwdParams bp{};
  size_t fwd_smem = 0, bwd_smem = 0;
  unsigned bar_base_tc = 0, bar_base_tcb = 0;
  // TMA-fed tcgen05 time loops, forward and backward (lstmp_recurrent_tma.cu)
  bool rec_tma = false;
  FwdTmaParams ftm{};
  BwdTmaParams btm{};
  size_t fwd_tma_smem = 0, bwd_tma_smem = 0;
  uint8_t *rhl = nullptr, *mhl = nullptr, *dghl = nullptr, *drhl = nullptr;
  int T_last = 0;        // frames of the last propagate (0 = none)
  bool have_bwd = false; // a backpropagate record exists for T_last
  unsigned long long launches = 0;
  int gemm_backend = 0;
  bool group_bwd_gemms = true;  // the GEMMs after the backward loop as one group (LSTMP_B200_GROUP_GEMMS=0: one by one)
  HlWorkspace hlws;
  // data-parallel exchange overlapped with the backward pass (lstmp_b200_set_nccl)
  void* nccl_comm = nullptr;
  cudaStream_t nccl_stream = nullptr;
  cudaEvent_t ev_block = nullptr, ev_comm_done = nullptr;
  bool exchange_pending = false;
  long long* dbg_stamps = nullptr;
  // weights-streamed mode: slices do not fit in shared memory -> per-step GEMMs + elementwise kernels
  bool streamed = false;
  float *dm = nullptr, *dc2 = nullptr;
  float* gemm_ws = nullptr;          // split-K partial sums
  size_t gemm_ws_floats = 0;
  // optional per-kernel event timing
  bool timing = false;
  struct Ev { int kind; cudaEvent_t a, b; };
  std::vector<Ev> events;
};

// RAII: brackets one kernel launch with events on its stream when timing is enabled.
struct Timed {
  lstmp_b200_engine* h;
  int kind;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  Timed(lstmp_b200_engine* h_, int kind_, cudaStream_t st_) : h(h_), kind(kind_), st(st_) {
    if (!h->timing) return;
    if (cudaEventCreate(&a) != cudaSuccess) {
      a = nullptr;
      return;
    }
    if (cudaEventCreate(&b) != cudaSuccess) {
      cudaEventDestroy(a);
      a = b = nullptr;
      return;
    }
    cudaEventRecord(a, st);
  }
  ~Timed() {
    if (a && b) {
      cudaEventRecord(b, st);
      h->events.push_back({kind, a, b});
    }
  }
};

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static bool make_decomp(int C, int R, int S, int sm_count, size_t smem_limit, int ngroups, Decomp* out,
                        FwdParams* fp, BwdParams* bp, size_t* fs, size_t* bs) {
  if (ngroups < 1 || ngroups > kMaxGroupsHost || S % ngroups != 0) return false;
  Decomp d;
  d.ngroups = ngroups;
  d.ctas_per_group = sm_count / ngroups;
  if (d.ctas_per_group < 1) return false;
  d.Sg = S / ngroups;
  d.cpc = (C + d.ctas_per_group - 1) / d.ctas_per_group;
  d.rpc = (R + d.ctas_per_group - 1) / d.ctas_per_group;
  long long per = (long long)d.Sg * R;
  long long piece = (per + d.ctas_per_group - 1) / d.ctas_per_group;
  d.piece = (int)((piece + 3) & ~3LL);
  d.fwd_xcap = d.bwd_xcap = 0;
  d.dbg = env_int("LSTMP_B200_DEBUG", 0);
  FwdParams f{};
  BwdParams b{};
  const size_t limit_floats = (smem_limit - (size_t)static_smem_reserve()) / sizeof(float);  // static __shared__ + slack
  size_t fsz = fwd_smem_floats(C, R, d, &f, limit_floats) * sizeof(float);
  size_t bsz = bwd_smem_floats(C, R, d, &b, limit_floats) * sizeof(float);
  if (fsz > smem_limit || bsz > smem_limit) return false;
  *out = d;
  *fp = f;
  *bp = b;
  *fs = fsz;
  *bs = bsz;
  return true;
}

extern "C" int lstmp_b200_abi_version(void) { return 1; }
extern "C" const char* lstmp_b200_last_error(void) { return g_err.c_str(); }
namespace lstmp {
void set_last_error(const char* msg) { g_err = msg ? msg : ""; }  // for the other translation units (lstmp_xent.cu)
}

static int alloc_f(float** p, size_t n, size_t* total) {
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(float));
  if (e != cudaSuccess) return fail(LSTMP_B200_ENOMEM, "cudaMalloc(%zu floats): %s", n, cudaGetErrorString(e));
  e = cudaMemset(*p, 0, n * sizeof(float));
  if (e != cudaSuccess) return fail((int)e, "cudaMemset: %s", cudaGetErrorString(e));
  *total += n * sizeof(float);
  return 0;
}

extern "C" int lstmp_b200_destroy(lstmp_b200_handle_t h) {
  if (!h) return 0;
  DeviceGuard device_guard(h->device);
  float* bufs[] = {h->params, h->corr, h->grads, h->state_c, h->state_r, h->gifo, h->cbuf, h->hbuf,
                   h->mbuf,   h->rbuf, h->dgifo, h->dr,      h->scratch, h->small_grads, h->dm, h->dc2, h->gemm_ws};
  for (float* b : bufs)
    if (b) cudaFree(b);
  if (h->bar) cudaFree(h->bar);
  if (h->ev_block) cudaEventDestroy(h->ev_block);
  if (h->ev_comm_done) cudaEventDestroy(h->ev_comm_done);
  gemm_hl_free(&h->hlws);
  if (h->rhl) cudaFree(h->rhl);  // one allocation holds rhl | mhl | dghl | drhl
  for (auto& e : h->events) {  // timing enabled but never read back
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  h->events.clear();
  if (h->dbg_stamps) {
    cudaDeviceSynchronize();
    std::vector<long long> st(2 + 2 * 1024);
    cudaMemcpy(st.data(), h->dbg_stamps, st.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long long n = st[0];
    fprintf(stderr, "[lstmp_b200 stamps] %lld records (tag: delta cycles from previous)\n", n);
    for (long long i = 0; i < n && i < 1024; ++i)
      fprintf(stderr, "  %4lld %8lld %10lld\n", st[2 + 2 * i], i ? st[3 + 2 * i] - st[1 + 2 * i] : 0LL,
              st[3 + 2 * i] - st[3]);
    cudaFree(h->dbg_stamps);
  }
  delete h;
  return 0;
}

extern "C" int lstmp_b200_create(int I, int C, int R, int S, int Tmax, int device, lstmp_b200_handle_t* out) {
  if (!out) return fail(LSTMP_B200_EINVAL, "out handle is NULL");
  *out = nullptr;
  if (I <= 0 || C <= 0 || R <= 0 || S <= 0 || Tmax <= 0)
    return fail(LSTMP_B200_EINVAL, "dimensions must be positive (I=%d C=%d R=%d S=%d T=%d)", I, C, R, S, Tmax);
  if (I % 4 || C % 4 || R % 4)
    return fail(LSTMP_B200_EINVAL, "input_dim, cell_dim and recur_dim must be multiples of 4 (I=%d C=%d R=%d)", I, C, R);
  if (S > 1024) return fail(LSTMP_B200_EINVAL, "num_stream %d > 1024 not supported", S);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(LSTMP_B200_ENODEV, "no CUDA device: %s (this engine has no CPU path)", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(LSTMP_B200_EINVAL, "device %d out of range (%d devices)", device, ndev);
  DeviceGuard device_guard(device);
  CUDA_TRY(device_guard.err);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(LSTMP_B200_ENODEV, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
                prop.minor);
  int coop = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  if (!coop) return fail(LSTMP_B200_ENODEV, "device %d lacks cooperative launch", device);

  lstmp_b200_engine* h = new (std::nothrow) lstmp_b200_engine();
  if (!h) return fail(LSTMP_B200_ENOMEM, "host allocation failed");
  h->I = I; h->C = C; h->R = R; h->S = S; h->Tmax = Tmax; h->device = device;
  h->sm_count = prop.multiProcessorCount;
  int sm_use = env_int("LSTMP_B200_MAX_CTAS", h->sm_count);
  if (sm_use < 1 || sm_use > h->sm_count) sm_use = h->sm_count;
  size_t smem_limit = prop.sharedMemPerBlockOptin;

  // work decomposition: most stream groups (least all-gather traffic per SM) whose weight
  // slices still fit in shared memory, keeping >= 16 streams per group.
  int forced = env_int("LSTMP_B200_NGROUPS", 0);
  bool ok = false;
  if (forced > 0) {
    ok = make_decomp(C, R, S, sm_use, smem_limit, forced, &h->d, &h->fp, &h->bp, &h->fwd_smem, &h->bwd_smem);
    if (!ok) {
      delete h;
      return fail(LSTMP_B200_ENOMEM, "LSTMP_B200_NGROUPS=%d does not fit (S=%d, smem limit %zu)", forced, S, smem_limit);
    }
  } else {
    for (int g = 8; g >= 1 && !ok; g >>= 1) {
      if (S % g != 0 || (g > 1 && S / g < 16)) continue;
      ok = make_decomp(C, R, S, sm_use, smem_limit, g, &h->d, &h->fp, &h->bp, &h->fwd_smem, &h->bwd_smem);
    }
  }
  if (env_int("LSTMP_B200_FORCE_STREAMED", 0)) ok = false;
  if (!ok) {
    // The slices of this layer do not fit in shared memory across the SMs (e.g. 2048-cell/1024-proj): fall back
    // to the weights-streamed per-timestep path (still CUDA only).
    h->streamed = true;
    h->d = Decomp{};
    h->d.ngroups = 1;
    h->d.ctas_per_group = sm_use;
    h->d.Sg = S;
  } else {
    e = set_kernel_smem_limits(h->fwd_smem, h->bwd_smem);
    if (e != cudaSuccess) {
      delete h;
      return fail((int)e, "cudaFuncSetAttribute(max dynamic smem): %s", cudaGetErrorString(e));
    }
    // default: tcgen05 loops fed by bulk copies in both directions (LSTMP_B200_REC=2); 0 = the FP32 FFMA kernels (also
    // the fallback for num_stream > 64 per group or cell / recurrent dims that are not multiples of 8)
    const int rec = env_int("LSTMP_B200_REC", 2);
    if (rec >= 2) {
      const int kp = env_int("LSTMP_B200_BWD_KP", 4);
      // stream groups: each group gathers only its own streams' activations (half the L2 -> SM traffic per CTA with
      // two groups) at the price of twice the weight slice per CTA; the largest count whose slices fit, >= 16 streams
      const int gforce = env_int("LSTMP_B200_TMA_GROUPS", 0);
      // the backward kernel may use its own group count (LSTMP_B200_TMA_GROUPS_BWD): its shared memory holds two
      // weight slices, so fewer groups (smaller slices) leave it a deeper operand ring
      const int gforce_b = env_int("LSTMP_B200_TMA_GROUPS_BWD", 0);
      for (int G = gforce > 0 ? gforce : 2; G >= 1 && !h->rec_tma; G = (gforce > 0 ? 0 : G - 1)) {
        if (G > 2 || S % G || (G > 1 && gforce <= 0 && S / G < 16)) continue;  // barrier counters for <= 2 groups
        const int Gb = (gforce_b > 0 && gforce_b <= 2 && S % gforce_b == 0) ? gforce_b : G;
        FwdTmaParams f{};
        BwdTmaParams b{};
        size_t fs = 0, bs = 0;
        bool okt = fwd_tma_plan(C, R, S, G, sm_use, smem_limit, &f, &fs);
        // the backward grid is the largest co-resident set of kp-CTA clusters; its shared-memory size depends on the
        // grid through the slice sizes, so plan with the optimistic grid first and again with the real one
        if (okt) okt = bwd_tma_plan(C, R, S, Gb, sm_use, kp, smem_limit, &b, &bs);
        if (okt) okt = tma_set_smem_limits(fs, bs) == cudaSuccess;
        if (okt) {
          const int n = bwd_tma_max_ctas(kp, bs, sm_use);
          okt = n >= kp * Gb && bwd_tma_plan(C, R, S, Gb, n, kp, smem_limit, &b, &bs) &&
                tma_set_smem_limits(fs, bs) == cudaSuccess && bwd_tma_max_ctas(kp, bs, sm_use) >= b.nctas;
        }
        if (okt) {
          h->rec_tma = true;
          f.stagger = b.stagger = env_int("LSTMP_B200_TC_STAGGER", 1);
          h->ftm = f;
          h->btm = b;
          h->fwd_tma_smem = fs;
          h->bwd_tma_smem = bs;
        } else {
          cudaGetLastError();
        }
      }
    }
  }

  // arenas
  size_t off = 0;
  h->off_wx = off; off += (size_t)4 * C * I;
  h->off_wr = off; off += (size_t)4 * C * R;
  h->off_bias = off; off += (size_t)4 * C;
  h->off_pi = off; off += C;
  h->off_pf = off; off += C;
  h->off_po = off; off += C;
  h->off_wm = off; off += (size_t)R * C;
  h->nparams = off;
  size_t ws = 0, dummy = 0;
  const size_t TS = (size_t)Tmax * S, TS1 = (size_t)(Tmax + 1) * S;
  int rc = 0;
  if ((rc = alloc_f(&h->params, h->nparams, &dummy)) || (rc = alloc_f(&h->corr, h->nparams, &dummy)) ||
      (rc = alloc_f(&h->grads, h->nparams, &dummy)) || (rc = alloc_f(&h->state_c, (size_t)S * C, &dummy)) ||
      (rc = alloc_f(&h->state_r, (size_t)S * R, &dummy)) || (rc = alloc_f(&h->gifo, TS * 4 * C, &ws)) ||
      (rc = alloc_f(&h->cbuf, TS1 * C, &ws)) || (rc = alloc_f(&h->hbuf, TS * C, &ws)) ||
      (rc = alloc_f(&h->mbuf, TS * C, &ws)) || (rc = alloc_f(&h->rbuf, TS1 * R, &ws)) ||
      (rc = alloc_f(&h->dgifo, TS * 4 * C, &ws)) || (rc = alloc_f(&h->dr, TS * R, &ws)) ||
      (rc = alloc_f(&h->scratch, h->streamed ? 4 : (size_t)h->d.ngroups * h->d.ctas_per_group * h->d.Sg * R, &ws)) ||
      (rc = alloc_f(&h->dm, h->streamed ? (size_t)S * C : 4, &ws)) ||
      (rc = alloc_f(&h->dc2, h->streamed ? (size_t)2 * S * C : 4, &ws)) ||
      (rc = alloc_f(&h->small_grads, (size_t)kMaxGroupsHost * 7 * C, &ws)) ||
      (rc = alloc_f(&h->gemm_ws, (size_t)4 << 20, &ws))) {
    lstmp_b200_destroy(h);
    return rc;
  }
  e = cudaMalloc((void**)&h->bar, (size_t)(kMaxGroupsHost + 1) * kBarStride * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(h->bar, 0, (size_t)(kMaxGroupsHost + 1) * kBarStride * sizeof(unsigned));
  if (e != cudaSuccess) {
    lstmp_b200_destroy(h);
    return fail((int)e, "barrier counters: %s", cudaGetErrorString(e));
  }
  if (h->rec_tma) {
    // exchange arrays: per group and 64-k chunk one tile image of roundup8(2*Sg) rows x 128 bytes; zero-initialised so
    // that the k tail of the last chunk (never written) contributes nothing
    const size_t tile = (size_t)((2 * h->ftm.Sg + 7) & ~7) * 128, G = (size_t)h->ftm.G;
    const size_t tile_b = (size_t)((2 * h->btm.Sg + 7) & ~7) * 128, Gb = (size_t)h->btm.G;
    const size_t n_r = G * h->ftm.nch_g * tile, n_m = G * h->ftm.nch_p * tile, n_dg = Gb * h->btm.nch_a * tile_b,
                 n_dr = Gb * h->btm.nch_b * tile_b;
    const size_t total = n_r + n_m + n_dg + n_dr;
    e = cudaMalloc((void**)&h->rhl, total);
    if (e == cudaSuccess) e = cudaMemset(h->rhl, 0, total);
    if (e != cudaSuccess) {
      lstmp_b200_destroy(h);
      return fail(LSTMP_B200_ENOMEM, "hi/lo exchange arrays: %s", cudaGetErrorString(e));
    }
    h->mhl = h->rhl + n_r;
    h->dghl = h->mhl + n_m;
    h->drhl = h->dghl + n_dg;
    ws += total;
  }
  h->gemm_ws_floats = env_int("LSTMP_B200_SPLITK", 1) ? ((size_t)4 << 20) : 0;
  h->workspace_bytes = ws;
  h->gemm_backend = 0;
#ifdef LSTMP_HAVE_TC_GEMM
  // 2 (default): bf16 hi/lo tile images + bulk copies (lstmp_gemm_hl.cu); 1: 3xTF32 with loader warps; 0: FP32 SIMT
  h->gemm_backend = env_int("LSTMP_B200_GEMM", 2);
  if (h->gemm_backend < 0 || h->gemm_backend > 2) h->gemm_backend = 2;
  h->group_bwd_gemms = env_int("LSTMP_B200_GROUP_GEMMS", 1) != 0;
#endif
  if (h->d.dbg & 12) {
    if (cudaMalloc((void**)&h->dbg_stamps, (2 + 2 * 1024) * sizeof(long long)) == cudaSuccess)
      cudaMemset(h->dbg_stamps, 0, (2 + 2 * 1024) * sizeof(long long));
    h->d.dbg_stamps = h->dbg_stamps;
  }
  CUDA_TRY(cudaDeviceSynchronize());
  *out = h;
  return 0;
}

#define CHECK_H(h)                                                                                        \
  if (!(h)) return fail(LSTMP_B200_EINVAL, "NULL handle");                                                \
  DeviceGuard device_guard__((h)->device);                                                                \
  if (device_guard__.err != cudaSuccess)                                                                  \
    return fail((int)device_guard__.err, "cudaSetDevice(%d): %s", (h)->device, cudaGetErrorString(device_guard__.err))

extern "C" int lstmp_b200_num_params(lstmp_b200_handle_t h, size_t* n) {
  if (!h || !n) return fail(LSTMP_B200_EINVAL, "NULL argument");
  *n = h->nparams;
  return 0;
}

static int copy2d(float* dst, size_t ldd, const float* src, size_t lds, size_t cols, size_t rows, cudaStream_t st) {
  if (rows == 0 || cols == 0) return 0;
  CUDA_TRY(cudaMemcpy2DAsync(dst, ldd * sizeof(float), src, lds * sizeof(float), cols * sizeof(float), rows,
                             cudaMemcpyDefault, st));
  return 0;
}

extern "C" int lstmp_b200_set_params(lstmp_b200_handle_t h, const float* wx, size_t ld_x, const float* wr,
                                     size_t ld_r, const float* bias, const float* pi, const float* pf,
                                     const float* po, const float* wm, size_t ld_m, void* stream) {
  CHECK_H(h);
  if (!wx || !wr || !bias || !pi || !pf || !po || !wm) return fail(LSTMP_B200_EINVAL, "NULL parameter pointer");
  if (ld_x < (size_t)h->I || ld_r < (size_t)h->R || ld_m < (size_t)h->C) return fail(LSTMP_B200_EINVAL, "stride < columns");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  float* P = h->params;
  if ((rc = copy2d(P + h->off_wx, h->I, wx, ld_x, h->I, (size_t)4 * h->C, st))) return rc;
  if ((rc = copy2d(P + h->off_wr, h->R, wr, ld_r, h->R, (size_t)4 * h->C, st))) return rc;
  if ((rc = copy2d(P + h->off_bias, 4 * h->C, bias, 4 * h->C, (size_t)4 * h->C, 1, st))) return rc;
  if ((rc = copy2d(P + h->off_pi, h->C, pi, h->C, h->C, 1, st))) return rc;
  if ((rc = copy2d(P + h->off_pf, h->C, pf, h->C, h->C, 1, st))) return rc;
  if ((rc = copy2d(P + h->off_po, h->C, po, h->C, h->C, 1, st))) return rc;
  if ((rc = copy2d(P + h->off_wm, h->C, wm, ld_m, h->C, h->R, st))) return rc;
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int lstmp_b200_get_params(lstmp_b200_handle_t h, float* wx, size_t ld_x, float* wr, size_t ld_r,
                                     float* bias, float* pi, float* pf, float* po, float* wm, size_t ld_m,
                                     void* stream) {
  CHECK_H(h);
  if (!wx || !wr || !bias || !pi || !pf || !po || !wm) return fail(LSTMP_B200_EINVAL, "NULL parameter pointer");
  if (ld_x < (size_t)h->I || ld_r < (size_t)h->R || ld_m < (size_t)h->C) return fail(LSTMP_B200_EINVAL, "stride < columns");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const float* P = h->params;
  if ((rc = copy2d(wx, ld_x, P + h->off_wx, h->I, h->I, (size_t)4 * h->C, st))) return rc;
  if ((rc = copy2d(wr, ld_r, P + h->off_wr, h->R, h->R, (size_t)4 * h->C, st))) return rc;
  if ((rc = copy2d(bias, 4 * h->C, P + h->off_bias, 4 * h->C, (size_t)4 * h->C, 1, st))) return rc;
  if ((rc = copy2d(pi, h->C, P + h->off_pi, h->C, h->C, 1, st))) return rc;
  if ((rc = copy2d(pf, h->C, P + h->off_pf, h->C, h->C, 1, st))) return rc;
  if ((rc = copy2d(po, h->C, P + h->off_po, h->C, h->C, 1, st))) return rc;
  if ((rc = copy2d(wm, ld_m, P + h->off_wm, h->C, h->C, h->R, st))) return rc;
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

static float* arena_of(lstmp_b200_handle_t h, int which) {
  return which == 0 ? h->params : which == 1 ? h->corr : which == 2 ? h->grads : nullptr;
}

extern "C" int lstmp_b200_get_flat(lstmp_b200_handle_t h, int which, float* dst, void* stream) {
  CHECK_H(h);
  float* a = arena_of(h, which);
  if (!a || !dst) return fail(LSTMP_B200_EINVAL, "bad arena %d or NULL dst", which);
  CUDA_TRY(cudaMemcpyAsync(dst, a, h->nparams * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
extern "C" int lstmp_b200_set_flat(lstmp_b200_handle_t h, int which, const float* src, void* stream) {
  CHECK_H(h);
  float* a = arena_of(h, which);
  if (!a || !src) return fail(LSTMP_B200_EINVAL, "bad arena %d or NULL src", which);
  CUDA_TRY(cudaMemcpyAsync(a, src, h->nparams * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
extern "C" int lstmp_b200_arena(lstmp_b200_handle_t h, int which, float** dev_ptr, size_t* count) {
  if (!h || !dev_ptr || !count) return fail(LSTMP_B200_EINVAL, "NULL argument");
  float* a = arena_of(h, which);
  if (!a) return fail(LSTMP_B200_EINVAL, "bad arena %d", which);
  *dev_ptr = a;
  *count = h->nparams;
  return 0;
}

extern "C" int lstmp_b200_get_state(lstmp_b200_handle_t h, float* c, size_t ld_c, float* r, size_t ld_r, void* stream) {
  CHECK_H(h);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (c && (rc = copy2d(c, ld_c, h->state_c, h->C, h->C, h->S, st))) return rc;
  if (r && (rc = copy2d(r, ld_r, h->state_r, h->R, h->R, h->S, st))) return rc;
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}
extern "C" int lstmp_b200_set_state(lstmp_b200_handle_t h, const float* c, size_t ld_c, const float* r, size_t ld_r,
                                    void* stream) {
  CHECK_H(h);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (c && (rc = copy2d(h->state_c, h->C, c, ld_c, h->C, h->S, st))) return rc;
  if (r && (rc = copy2d(h->state_r, h->R, r, ld_r, h->R, h->S, st))) return rc;
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int lstmp_b200_clone(lstmp_b200_handle_t src, lstmp_b200_handle_t* out) {
  CHECK_H(src);
  int rc = lstmp_b200_create(src->I, src->C, src->R, src->S, src->Tmax, src->device, out);
  if (rc) return rc;
  lstmp_b200_handle_t d = *out;
  cudaError_t e = cudaMemcpy(d->params, src->params, src->nparams * sizeof(float), cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d->corr, src->corr, src->nparams * sizeof(float), cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(d->state_c, src->state_c, (size_t)src->S * src->C * sizeof(float), cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(d->state_r, src->state_r, (size_t)src->S * src->R * sizeof(float), cudaMemcpyDeviceToDevice);
  if (e != cudaSuccess) {  // do not leak the half-initialised twin
    lstmp_b200_destroy(d);
    *out = nullptr;
    return fail((int)e, "clone: %s", cudaGetErrorString(e));
  }
  return 0;
}

extern "C" int lstmp_b200_reset(lstmp_b200_handle_t h, const int32_t* flags, int n, void* stream) {
  CHECK_H(h);
  if (!flags) return fail(LSTMP_B200_EINVAL, "NULL flags");
  if (n != h->S) return fail(LSTMP_B200_EINVAL, "reset: %d flags for %d streams (LPS.h:214 asserts equality)", n, h->S);
  ResetMask m;
  memset(&m, 0, sizeof m);
  bool any = false;
  for (int s = 0; s < n; ++s)
    if (flags[s] == 1) {
      m.w[s >> 5] |= 1u << (s & 31);
      any = true;
    }
  if (!any) return 0;
  {
    Timed tm(h, 7, (cudaStream_t)stream);
    CUDA_TRY(launch_reset(h->state_c, h->C, h->state_r, h->R, h->S, m, (cudaStream_t)stream));
  }
  h->launches++;
  return 0;
}

static int gemm(lstmp_b200_handle_t h, int kind, float* C, long long ldc, int M, int N, int K, float alpha, const float* A,
                long long lda, int tA, const float* B, long long ldb, int tB, float beta, const float* bias,
                cudaStream_t st, bool reuse_a = false) {
  Timed tm(h, kind, st);
#ifdef LSTMP_HAVE_TC_GEMM
  // (the per-timestep GEMMs of the weights-streamed mode, kinds 1 and 2, re-read the same weights every step: splitting
  // them into tile images 2*T times per chunk costs more than it saves -- they stay on the 3xTF32 kernel)
  if (h->gemm_backend == 2 && !(h->streamed && (kind == 1 || kind == 2))) {
    bool handled = false;
    int nl = 0;
    CUDA_TRY(launch_gemm_hl(&h->hlws, C, ldc, M, N, K, alpha, A, lda, tA, B, ldb, tB, beta, bias, st, &handled, h->gemm_ws,
                            h->gemm_ws_floats, &nl, reuse_a));
    h->launches += nl;
    if (handled) return 0;
  }
  if (h->gemm_backend >= 1) {
    bool handled = false;
    int nl = 1;
    CUDA_TRY(launch_gemm_tc(C, ldc, M, N, K, alpha, A, lda, tA, B, ldb, tB, beta, bias, st, &handled, h->gemm_ws,
                            h->gemm_ws_floats, &nl));
    if (handled) {
      h->launches += nl;
      return 0;
    }
  }
#endif
  CUDA_TRY(launch_gemm_simt(C, ldc, M, N, K, alpha, A, lda, tA, B, ldb, tB, beta, bias, st));
  h->launches++;
  return 0;
}

extern "C" int lstmp_b200_propagate(lstmp_b200_handle_t h, const float* in, size_t ld_in, float* out, size_t ld_out,
                                    int num_rows, void* stream) {
  CHECK_H(h);
  if (!in || !out) return fail(LSTMP_B200_EINVAL, "NULL in/out");
  if (num_rows <= 0 || num_rows % h->S != 0)
    return fail(LSTMP_B200_EINVAL, "propagate: %d rows is not a positive multiple of NumStream=%d (LPS.h:225)", num_rows, h->S);
  const int T = num_rows / h->S;
  if (T > h->Tmax) return fail(LSTMP_B200_EINVAL, "propagate: T=%d exceeds max_frames=%d given at create", T, h->Tmax);
  if (ld_in < (size_t)h->I || ld_out < (size_t)h->R) return fail(LSTMP_B200_EINVAL, "stride < columns");
  cudaStream_t st = (cudaStream_t)stream;
  const int C = h->C, R = h->R, I = h->I, S = h->S;
  int rc;
  // YGIFO[1..T] = in * w_gifo_x^T + bias                                  (LPS.h:246,259)
  if ((rc = gemm(h, 0, h->gifo, 4 * C, num_rows, 4 * C, I, 1.f, in, (long long)ld_in, 0, h->params + h->off_wx, I, 1,
                 0.f, h->params + h->off_bias, st)))
    return rc;
  if (h->streamed) {
    const float* W = h->params;
    CUDA_TRY(cudaMemcpyAsync(h->cbuf, h->state_c, (size_t)S * C * sizeof(float), cudaMemcpyDeviceToDevice, st));  // :231
    CUDA_TRY(cudaMemcpyAsync(h->rbuf, h->state_r, (size_t)S * R * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int t = 0; t < T; ++t) {
      float* gifo_t = h->gifo + (size_t)t * S * 4 * C;
      // gifo(t) += r(t-1) * W_gifo_r^T                                       (LPS.h:275)
      if ((rc = gemm(h, 1, gifo_t, 4 * C, S, 4 * C, R, 1.f, h->rbuf + (size_t)t * S * R, R, 0, W + h->off_wr, R, 1, 1.f,
                     nullptr, st)))
        return rc;
      {
        Timed tm(h, 1, st);
        CUDA_TRY(launch_streamed_fwd_elem(gifo_t, h->cbuf + (size_t)t * S * C, h->cbuf + (size_t)(t + 1) * S * C,
                                          h->hbuf + (size_t)t * S * C, h->mbuf + (size_t)t * S * C, W + h->off_pi,
                                          W + h->off_pf, W + h->off_po, S, C, st));
      }
      h->launches++;
      // r(t) = m(t) * W_r_m^T                                                  (LPS.h:312)
      if ((rc = gemm(h, 1, h->rbuf + (size_t)(t + 1) * S * R, R, S, R, C, 1.f, h->mbuf + (size_t)t * S * C, C, 0,
                     W + h->off_wm, C, 1, 0.f, nullptr, st)))
        return rc;
    }
    CUDA_TRY(cudaMemcpy2DAsync(out, ld_out * sizeof(float), h->rbuf + (size_t)S * R, R * sizeof(float), R * sizeof(float),
                               (size_t)num_rows, cudaMemcpyDeviceToDevice, st));                                  // :328
    CUDA_TRY(cudaMemcpyAsync(h->state_c, h->cbuf + (size_t)T * S * C, (size_t)S * C * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));                                                       // :331
    CUDA_TRY(cudaMemcpyAsync(h->state_r, h->rbuf + (size_t)T * S * R, (size_t)S * R * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
    h->T_last = T;
    h->have_bwd = false;
    return 0;
  }
  if (h->rec_tma) {
    FwdTmaParams q = h->ftm;
    q.I = I; q.C = C; q.R = R; q.S = S; q.T = T;
    q.w_gifo_r = h->params + h->off_wr;
    q.w_r_m = h->params + h->off_wm;
    q.p_i = h->params + h->off_pi;
    q.p_f = h->params + h->off_pf;
    q.p_o = h->params + h->off_po;
    q.gifo = h->gifo; q.cbuf = h->cbuf; q.hbuf = h->hbuf; q.mbuf = h->mbuf; q.rbuf = h->rbuf;
    q.out = out;
    q.ld_out = (long long)ld_out;
    q.state_c = h->state_c;
    q.state_r = h->state_r;
    q.rhl = h->rhl; q.mhl = h->mhl;
    q.bar = h->bar + (size_t)kMaxGroupsHost * kBarStride;
    q.bar_base = h->bar_base_tc;
    q.dbg = h->d.dbg;
    q.dbg_stamps = h->dbg_stamps;
    {
      Timed tm(h, 1, st);
      CUDA_TRY(launch_fwd_tma(q, h->fwd_tma_smem, st));
    }
    h->launches++;
    h->bar_base_tc += (unsigned)(fwd_tma_barriers(T) * q.cpg);
    h->T_last = T;
    h->have_bwd = false;
    return 0;
  }
  FwdParams p = h->fp;
  p.I = I; p.C = C; p.R = R; p.S = S; p.T = T;
  p.d = h->d;
  p.w_gifo_r = h->params + h->off_wr;
  p.w_r_m = h->params + h->off_wm;
  p.p_i = h->params + h->off_pi;
  p.p_f = h->params + h->off_pf;
  p.p_o = h->params + h->off_po;
  p.gifo = h->gifo; p.cbuf = h->cbuf; p.hbuf = h->hbuf; p.mbuf = h->mbuf; p.rbuf = h->rbuf;
  p.out = out;
  p.ld_out = (long long)ld_out;
  p.state_c = h->state_c;
  p.state_r = h->state_r;
  p.bar = h->bar;
  for (int g = 0; g < h->d.ngroups; ++g) p.bar_base[g] = h->bar_base[g];
  {
    Timed tm(h, 1, st);
    CUDA_TRY(launch_fwd(p, h->fwd_smem, st));
  }
  h->launches++;
  for (int g = 0; g < h->d.ngroups; ++g) h->bar_base[g] += (unsigned)(fwd_barriers(T) * h->d.ctas_per_group);
  h->T_last = T;
  h->have_bwd = false;
  return 0;
}

// ---- NCCL, resolved at run time ------------------------------------------------------------
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
static int nccl_sum(float* buf, size_t count, void* comm, cudaStream_t stream) {
  static nccl_allreduce_fn fn = nullptr;
  if (!fn) {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(LSTMP_B200_EUNSUPPORTED, "dlopen(libnccl.so.2): %s", dlerror());
    fn = (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce");
    if (!fn) return fail(LSTMP_B200_EUNSUPPORTED, "ncclAllReduce not found");
  }
  // ncclFloat32 = 7, ncclSum = 0
  int rc = fn(buf, buf, count, 7, 0, comm, stream);
  if (rc != 0) return fail(LSTMP_B200_EUNSUPPORTED, "ncclAllReduce returned %d", rc);
  return 0;
}
// Overlapped exchange (lstmp_b200_set_nccl): the gradient block [off, off + count) of the arena is final on the compute
// stream `st` -- sum it over the ranks on the communication stream while the remaining gradient GEMMs run.
static int exchange_block(lstmp_b200_handle_t h, size_t off, size_t count, cudaStream_t st, bool last) {
  if (!h->nccl_comm) return 0;
  CUDA_TRY(cudaEventRecord(h->ev_block, st));
  CUDA_TRY(cudaStreamWaitEvent(h->nccl_stream, h->ev_block, 0));
  int rc = nccl_sum(h->grads + off, count, h->nccl_comm, h->nccl_stream);
  if (rc) return rc;
  if (last) {
    CUDA_TRY(cudaEventRecord(h->ev_comm_done, h->nccl_stream));
    h->exchange_pending = true;
  }
  return 0;
}

extern "C" int lstmp_b200_backpropagate(lstmp_b200_handle_t h, const float* in, size_t ld_in, const float* out_diff,
                                        size_t ld_od, float* in_diff, size_t ld_id, int num_rows, void* stream) {
  CHECK_H(h);
  if (!in || !out_diff) return fail(LSTMP_B200_EINVAL, "NULL in/out_diff");
  if (h->T_last == 0) return fail(LSTMP_B200_ESTATE, "backpropagate without a preceding propagate");
  if (num_rows != h->T_last * h->S)
    return fail(LSTMP_B200_EINVAL, "backpropagate: %d rows but the last propagate had %d", num_rows, h->T_last * h->S);
  if (ld_in < (size_t)h->I || ld_od < (size_t)h->R || (in_diff && ld_id < (size_t)h->I))
    return fail(LSTMP_B200_EINVAL, "stride < columns");
  cudaStream_t st = (cudaStream_t)stream;
  const int C = h->C, R = h->R, I = h->I, S = h->S, T = h->T_last;
  int rc;
  if (h->streamed) {
    const float* W = h->params;
    CUDA_TRY(cudaMemcpy2DAsync(h->dr, R * sizeof(float), out_diff, ld_od * sizeof(float), R * sizeof(float),
                               (size_t)num_rows, cudaMemcpyDeviceToDevice, st));                                  // :367
    for (int t = T - 1; t >= 0; --t) {
      const bool next = t + 1 < T;
      float* dr_t = h->dr + (size_t)t * S * R;
      // d_r(t) += DGIFO(t+1) * w_gifo_r                                        (LPS.h:391)
      if (next && (rc = gemm(h, 2, dr_t, R, S, R, 4 * C, 1.f, h->dgifo + (size_t)(t + 1) * S * 4 * C, 4 * C, 0,
                             W + h->off_wr, R, 0, 1.f, nullptr, st)))
        return rc;
      // d_m = d_r * w_r_m                                                       (LPS.h:408)
      if ((rc = gemm(h, 2, h->dm, C, S, C, R, 1.f, dr_t, R, 0, W + h->off_wm, C, 0, 0.f, nullptr, st))) return rc;
      float* dc_t = h->dc2 + (size_t)(t & 1) * S * C;
      const float* dc_n = h->dc2 + (size_t)((t + 1) & 1) * S * C;
      {
        Timed tm(h, 2, st);
        CUDA_TRY(launch_streamed_bwd_elem(h->dm, h->gifo + (size_t)t * S * 4 * C,
                                          next ? h->gifo + (size_t)(t + 1) * S * 4 * C : nullptr,
                                          h->cbuf + (size_t)(t + 1) * S * C, h->cbuf + (size_t)t * S * C,
                                          h->hbuf + (size_t)t * S * C,
                                          next ? h->dgifo + (size_t)(t + 1) * S * 4 * C : nullptr, next ? dc_n : nullptr,
                                          h->dgifo + (size_t)t * S * 4 * C, dc_t, W + h->off_pi, W + h->off_pf,
                                          W + h->off_po, S, C, st));
      }
      h->launches++;
    }
    {
      Timed tm(h, 5, st);
      CUDA_TRY(launch_streamed_small_grads(h->dgifo, h->cbuf, h->grads + h->off_bias, num_rows, S, C, st));
    }
    h->launches++;
  } else if (h->rec_tma) {
    BwdTmaParams q = h->btm;
    q.I = I; q.C = C; q.R = R; q.S = S; q.T = T;
    q.w_gifo_r = h->params + h->off_wr;
    q.w_r_m = h->params + h->off_wm;
    q.p_i = h->params + h->off_pi;
    q.p_f = h->params + h->off_pf;
    q.p_o = h->params + h->off_po;
    q.gifo = h->gifo; q.cbuf = h->cbuf; q.hbuf = h->hbuf;
    q.out_diff = out_diff;
    q.ld_od = (long long)ld_od;
    q.dgifo = h->dgifo; q.dr = h->dr;
    // bias / peephole gradients (LPS.h:474-484): straight into the arena, or per-group partials summed below
    q.g_small = q.G == 1 ? h->grads + h->off_bias : h->small_grads;
    q.dghl = h->dghl; q.drhl = h->drhl;
    // its own counter: the forward and backward grids differ in size (clusters)
    q.bar = h->bar + (size_t)kMaxGroupsHost * kBarStride + 128;  // forward: groups at +0, +64
    q.bar_base = h->bar_base_tcb;
    q.dbg = h->d.dbg;
    q.dbg_stamps = h->dbg_stamps;
    {
      Timed tm(h, 2, st);
      CUDA_TRY(launch_bwd_tma(q, h->bwd_tma_smem, st));
    }
    h->launches++;
    h->bar_base_tcb += (unsigned)(bwd_barriers(T) * q.cpg);
    if (q.G > 1) {
      Timed tm(h, 5, st);
      CUDA_TRY(launch_small_grads(h->grads + h->off_bias, h->small_grads, q.G, 7 * C, st));
      h->launches++;
    }
  } else {
  BwdParams p = h->bp;
  p.I = I; p.C = C; p.R = R; p.S = S; p.T = T;
  p.d = h->d;
  p.w_gifo_r = h->params + h->off_wr;
  p.w_r_m = h->params + h->off_wm;
  p.p_i = h->params + h->off_pi;
  p.p_f = h->params + h->off_pf;
  p.p_o = h->params + h->off_po;
  p.gifo = h->gifo; p.cbuf = h->cbuf; p.hbuf = h->hbuf;
  p.out_diff = out_diff;
  p.ld_od = (long long)ld_od;
  p.dgifo = h->dgifo; p.dr = h->dr; p.scratch = h->scratch; p.small_grads = h->small_grads;
  p.bar = h->bar;
  for (int g = 0; g < h->d.ngroups; ++g) p.bar_base[g] = h->bar_base[g];
  {
    Timed tm(h, 2, st);
    CUDA_TRY(launch_bwd(p, h->bwd_smem, st));
  }
  h->launches++;
  for (int g = 0; g < h->d.ngroups; ++g) h->bar_base[g] += (unsigned)(bwd_barriers(T) * h->d.ctas_per_group);
  // bias / peephole gradients: sum the per-group partials                  (LPS.h:474-484)
  {
    Timed tm(h, 5, st);
    CUDA_TRY(launch_small_grads(h->grads + h->off_bias, h->small_grads, h->d.ngroups, 7 * C, st));
  }
  h->launches++;
  }
  // The four contractions that follow the time loop are independent of each other:
  //   in_diff     = DGIFO[1..T]   * w_gifo_x        (LPS.h:457)
  //   G(w_gifo_x) = DGIFO[1..T]^T * in              (LPS.h:468)
  //   G(w_gifo_r) = DGIFO[1..T]^T * R[0..T-1]       (LPS.h:471)
  //   G(w_r_m)    = DR[1..T]^T    * M[1..T]         (LPS.h:486)
  // The tile-image back end runs them as ONE group: every operand split in one launch (DGIFO^T once for both of its
  // products), every product in one persistent launch whose work items share the 148 CTAs, one split-K reduce.
  bool grouped = false;
#ifdef LSTMP_HAVE_TC_GEMM
  if (h->gemm_backend == 2 && h->group_bwd_gemms) {
    Timed tm(h, 4, st);
    HlGemmDesc d[4];
    int n = 0;
    if (in_diff)
      d[n++] = HlGemmDesc{in_diff, (long long)ld_id, num_rows, I, 4 * C, 1.f, h->dgifo, 4 * C, 0,
                          h->params + h->off_wx, I, 0, 0.f, nullptr};
    d[n++] = HlGemmDesc{h->grads + h->off_wx, I, 4 * C, I, num_rows, 1.f, h->dgifo, 4 * C, 1, in, (long long)ld_in, 0, 0.f,
                        nullptr};
    d[n++] = HlGemmDesc{h->grads + h->off_wr, R, 4 * C, R, num_rows, 1.f, h->dgifo, 4 * C, 1, h->rbuf, R, 0, 0.f, nullptr};
    d[n++] = HlGemmDesc{h->grads + h->off_wm, C, R, C, num_rows, 1.f, h->dr, R, 1, h->mbuf, C, 0, 0.f, nullptr};
    int nl = 0;
    CUDA_TRY(launch_gemm_hl_group(&h->hlws, d, n, st, &grouped, h->gemm_ws, h->gemm_ws_floats, &nl));
    h->launches += nl;
  }
#endif
  if (!grouped) {
    if (in_diff && (rc = gemm(h, 3, in_diff, (long long)ld_id, num_rows, I, 4 * C, 1.f, h->dgifo, 4 * C, 0,
                              h->params + h->off_wx, I, 0, 0.f, nullptr, st)))
      return rc;
    if ((rc = gemm(h, 4, h->grads + h->off_wx, I, 4 * C, I, num_rows, 1.f, h->dgifo, 4 * C, 1, in, (long long)ld_in, 0,
                   0.f, nullptr, st)))
      return rc;
  }
  if ((rc = exchange_block(h, h->off_wx, (size_t)4 * C * I, st, false))) return rc;
  if (!grouped) {
    if ((rc = gemm(h, 4, h->grads + h->off_wr, R, 4 * C, R, num_rows, 1.f, h->dgifo, 4 * C, 1, h->rbuf, R, 0, 0.f,
                   nullptr, st, /*reuse_a: DGIFO^T was split for the previous GEMM*/ true)))
      return rc;
  }
  // w_gifo_r | bias | peepholes are contiguous in the arena; bias / peepholes were written before the GEMMs
  if ((rc = exchange_block(h, h->off_wr, h->off_wm - h->off_wr, st, false))) return rc;
  if (!grouped) {
    if ((rc = gemm(h, 4, h->grads + h->off_wm, C, R, C, num_rows, 1.f, h->dr, R, 1, h->mbuf, C, 0, 0.f, nullptr, st)))
      return rc;
  }
  if ((rc = exchange_block(h, h->off_wm, (size_t)R * C, st, true))) return rc;
  h->have_bwd = true;
  return 0;
}

static int wait_exchange(lstmp_b200_handle_t h, cudaStream_t st) {
  if (h->exchange_pending) {
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_comm_done, 0));
    h->exchange_pending = false;
  }
  return 0;
}

extern "C" int lstmp_b200_update(lstmp_b200_handle_t h, float learn_rate, float momentum, void* stream) {
  CHECK_H(h);
  if (int rc = wait_exchange(h, (cudaStream_t)stream)) return rc;
  {
    Timed tm(h, 6, (cudaStream_t)stream);
    CUDA_TRY(launch_update(h->params, h->corr, h->grads, h->nparams, learn_rate, momentum, 0.f, (cudaStream_t)stream));
  }
  h->launches++;
  return 0;
}

extern "C" int lstmp_b200_update_clipped(lstmp_b200_handle_t h, float learn_rate, float momentum, float max_grad,
                                         void* stream) {
  CHECK_H(h);
  if (int rc = wait_exchange(h, (cudaStream_t)stream)) return rc;
  {
    Timed tm(h, 6, (cudaStream_t)stream);
    CUDA_TRY(launch_update(h->params, h->corr, h->grads, h->nparams, learn_rate, momentum, max_grad > 0.f ? max_grad : 0.f,
                           (cudaStream_t)stream));
  }
  h->launches++;
  return 0;
}

extern "C" int lstmp_b200_allreduce_grads_nccl(lstmp_b200_handle_t h, void* comm, void* stream) {
  CHECK_H(h);
  if (!comm) return fail(LSTMP_B200_EINVAL, "NULL ncclComm_t");
  return nccl_sum(h->grads, h->nparams, comm, (cudaStream_t)stream);
}

extern "C" int lstmp_b200_set_nccl(lstmp_b200_handle_t h, void* comm, void* comm_stream) {
  CHECK_H(h);
  if (comm && !h->ev_block) {
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_block, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_comm_done, cudaEventDisableTiming));
  }
  h->nccl_comm = comm;
  h->nccl_stream = (cudaStream_t)comm_stream;
  h->exchange_pending = false;
  return 0;
}

extern "C" int lstmp_b200_get_info(lstmp_b200_handle_t h, lstmp_b200_info_t* info) {
  if (!h || !info) return fail(LSTMP_B200_EINVAL, "NULL argument");
  memset(info, 0, sizeof *info);
  info->input_dim = h->I; info->cell_dim = h->C; info->recur_dim = h->R; info->num_stream = h->S;
  info->max_frames = h->Tmax; info->sm_count = h->sm_count;
  info->ngroups = h->d.ngroups; info->ctas_per_group = h->d.ctas_per_group; info->streams_per_group = h->d.Sg;
  info->cells_per_cta = h->d.cpc; info->rcols_per_cta = h->d.rpc;
  info->fwd_smem_bytes = h->rec_tma ? h->fwd_tma_smem : h->fwd_smem;
  info->bwd_smem_bytes = h->rec_tma ? h->bwd_tma_smem : h->bwd_smem;
  info->workspace_bytes = h->workspace_bytes;
  info->kernel_launches = h->launches;
  info->gemm_backend = h->gemm_backend;
  info->weights_streamed = h->streamed ? 1 : 0;
  info->fwd_tensor_core = h->rec_tma ? 2 : 0;
  info->bwd_tensor_core = h->rec_tma ? 2 : 0;
  if (h->rec_tma) {
    info->ngroups = h->ftm.G;
    info->ctas_per_group = h->ftm.cpg;
    info->streams_per_group = h->ftm.Sg;
    info->cells_per_cta = h->ftm.cpc;
    info->rcols_per_cta = h->ftm.rpc;
    info->bwd_ctas = h->btm.nctas;
    info->bwd_cluster = h->btm.kp;
  }
  return 0;
}

extern "C" int lstmp_b200_timing_enable(lstmp_b200_handle_t h, int on) {
  if (!h) return fail(LSTMP_B200_EINVAL, "NULL handle");
  h->timing = on != 0;
  return 0;
}
extern "C" int lstmp_b200_timing_read(lstmp_b200_handle_t h, lstmp_b200_timing_t* out) {
  CHECK_H(h);
  if (!out) return fail(LSTMP_B200_EINVAL, "NULL out");
  memset(out, 0, sizeof *out);
  CUDA_TRY(cudaDeviceSynchronize());
  for (auto& e : h->events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess && e.kind >= 0 && e.kind < LSTMP_B200_TIMING_KINDS) {
      out->ms[e.kind] += ms;
      out->count[e.kind]++;
    }
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  h->events.clear();
  return 0;
}

extern "C" int lstmp_b200_get_record(lstmp_b200_handle_t h, int backward, float* dst, size_t ld, void* stream) {
  CHECK_H(h);
  if (!dst) return fail(LSTMP_B200_EINVAL, "NULL dst");
  if (h->T_last == 0) return fail(LSTMP_B200_ESTATE, "no propagate record");
  if (backward && !h->have_bwd) return fail(LSTMP_B200_ESTATE, "no backpropagate record");
  const int C = h->C, R = h->R, S = h->S, T = h->T_last;
  const size_t W = (size_t)7 * C + R, rows = (size_t)T * S;
  if (ld < W) return fail(LSTMP_B200_EINVAL, "ld_dst < 7C+R");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (!backward) {
    if ((rc = copy2d(dst, ld, h->gifo, 4 * C, 4 * C, rows, st))) return rc;
    if ((rc = copy2d(dst + 4 * C, ld, h->cbuf + (size_t)S * C, C, C, rows, st))) return rc;
    if ((rc = copy2d(dst + 5 * C, ld, h->hbuf, C, C, rows, st))) return rc;
    if ((rc = copy2d(dst + 6 * C, ld, h->mbuf, C, C, rows, st))) return rc;
    if ((rc = copy2d(dst + 7 * C, ld, h->rbuf + (size_t)S * R, R, R, rows, st))) return rc;
  } else {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, dst);
    bool on_device = (e == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged));
    if (e != cudaSuccess) cudaGetLastError();
    if (on_device) {
      CUDA_TRY(cudaMemset2DAsync(dst + 4 * C, ld * sizeof(float), 0, (size_t)3 * C * sizeof(float), rows, st));
    } else {
      CUDA_TRY(cudaStreamSynchronize(st));
      for (size_t r = 0; r < rows; ++r) memset(dst + r * ld + 4 * C, 0, (size_t)3 * C * sizeof(float));
    }
    if ((rc = copy2d(dst, ld, h->dgifo, 4 * C, 4 * C, rows, st))) return rc;
    if ((rc = copy2d(dst + 7 * C, ld, h->dr, R, R, rows, st))) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int lstmp_b200_debug_gemm(int backend, float* C, size_t ldc, int M, int N, int K, float alpha,
                                     const float* A, size_t lda, int tA, const float* B, size_t ldb, int tB, float beta,
                                     const float* bias, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (backend == 0) {
    CUDA_TRY(launch_gemm_simt(C, (long long)ldc, M, N, K, alpha, A, (long long)lda, tA, B, (long long)ldb, tB, beta,
                              bias, st));
    return 0;
  }
#ifdef LSTMP_HAVE_TC_GEMM
  if (backend == 1) {
    bool handled = false;
    int handled_n = 0;
    static float* dbg_ws = nullptr;   // test hook only: 16 MB split-K workspace, allocated once
    const size_t dbg_ws_floats = (size_t)4 << 20;
    if (!dbg_ws && cudaMalloc((void**)&dbg_ws, dbg_ws_floats * sizeof(float)) != cudaSuccess) dbg_ws = nullptr;
    CUDA_TRY(launch_gemm_tc(C, (long long)ldc, M, N, K, alpha, A, (long long)lda, tA, B, (long long)ldb, tB, beta, bias,
                            st, &handled, dbg_ws, dbg_ws ? dbg_ws_floats : 0, &handled_n));
    if (!handled) return fail(LSTMP_B200_EUNSUPPORTED, "tcgen05 GEMM does not handle this shape/alignment");
    return 0;
  }
  if (backend == 2) {
    bool handled = false;
    int nl = 0;
    static HlWorkspace dbg_hl;          // test hook only
    static float* dbg_ws2 = nullptr;
    const size_t dbg_ws_floats = (size_t)4 << 20;
    if (!dbg_ws2 && cudaMalloc((void**)&dbg_ws2, dbg_ws_floats * sizeof(float)) != cudaSuccess) dbg_ws2 = nullptr;
    CUDA_TRY(launch_gemm_hl(&dbg_hl, C, (long long)ldc, M, N, K, alpha, A, (long long)lda, tA, B, (long long)ldb, tB, beta,
                            bias, st, &handled, dbg_ws2, dbg_ws2 ? dbg_ws_floats : 0, &nl, false));
    if (!handled) return fail(LSTMP_B200_EUNSUPPORTED, "bf16 hi/lo GEMM does not handle this alignment");
    return 0;
  }
#endif
  return fail(LSTMP_B200_EUNSUPPORTED, "unknown GEMM backend %d", backend);
}

extern "C" int lstmp_b200_debug_gemm_group(int n, const lstmp_b200_gemm_desc* d, void* stream) {
#ifdef LSTMP_HAVE_TC_GEMM
  if (!d || n < 1 || n > 4) return fail(LSTMP_B200_EINVAL, "gemm group: 1..4 products");
  HlGemmDesc g[4];
  for (int i = 0; i < n; ++i)
    g[i] = HlGemmDesc{d[i].C, (long long)d[i].ldc, d[i].M, d[i].N, d[i].K, d[i].alpha, d[i].A, (long long)d[i].lda, d[i].tA,
                      d[i].B, (long long)d[i].ldb, d[i].tB, d[i].beta, d[i].bias};
  bool handled = false;
  int nl = 0;
  static HlWorkspace dbg_hl;          // test hook only
  static float* dbg_ws = nullptr;
  const size_t dbg_ws_floats = (size_t)4 << 20;
  if (!dbg_ws && cudaMalloc((void**)&dbg_ws, dbg_ws_floats * sizeof(float)) != cudaSuccess) dbg_ws = nullptr;
  CUDA_TRY(launch_gemm_hl_group(&dbg_hl, g, n, (cudaStream_t)stream, &handled, dbg_ws, dbg_ws ? dbg_ws_floats : 0, &nl));
  if (!handled) return fail(LSTMP_B200_EUNSUPPORTED, "bf16 hi/lo GEMM group does not handle this alignment");
  return nl;
#else
  (void)n; (void)d; (void)stream;
  return fail(LSTMP_B200_EUNSUPPORTED, "built without the tensor-core GEMMs");
#endif
}
