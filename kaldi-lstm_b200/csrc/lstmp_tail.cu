// The output tail of the acoustic model: AffineTransform 512 -> 16624 + Softmax +
// Xent::EvalMasked, forward and backward, as ONE component behind the C ABI (lstmp_b200_tail_*), so that
// BASELINE.json configs[3] (LstmProjectedStreams 800/512 + AffineTransform 512->16624 + Softmax,
// /root/reference README.md:25-28, google/nnet.proto:4-5; run by the trainer at
// google/nnetbin/bd-nnet-train-lstm-streams.cc:215-228) runs end to end on the device.
//
// Semantics, [upstream] Kaldi nnet1 of the reference's vintage (not vendored in the reference tree; the test
// suite checks them against a numpy restatement):
//   AffineTransform::PropagateFnc      out = in * W^T + b                       (W = linearity_ [num_pdf x input_dim])
//   Softmax::PropagateFnc              y = softmax per row (subtract the row max, exp, scale by 1/sum)
//   Xent::EvalMasked                   diff = mask * (y - t), loss / entropy / correct / frames (nnet-loss.cc:76-164)
//   Softmax::BackpropagateFnc          in_diff = out_diff  (the soft-max derivative is folded into the xent diff)
//   AffineTransform::BackpropagateFnc  in_diff = diff * W
//   AffineTransform::Update            W_corr = diff^T * in + momentum * W_corr ; b_corr = colsum(diff) + momentum * b_corr
//                                      W -= lr * W_corr ; b -= lr * b_corr        (learn-rate coefficients 1, no L1/L2)
// Here: logits = tcgen05 3xTF32 GEMM with the bias fused; softmax + xent fused in one kernel that reads a row of logits
// from HBM once into shared memory and writes diff IN PLACE over the logits (the dense posteriors are materialised only
// on request); in_diff and W_grad are two more GEMMs off that diff; the bias gradient is a column-sum kernel.  The
// fresh gradients live in one flat arena [W | b] (what a data-parallel caller all-reduces, section 8e); the momentum
// step is the same fused update kernel as the LSTM layers'.  No CPU path.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "../../include/lstmp_b200.h"
#include "lstmp_kernels.h"

namespace lstmp {
void set_last_error(const char* msg);
cudaError_t launch_gemm_simt(float* C, long long ldc, int M, int N, int K, float alpha, const float* A, long long lda,
                             int tA, const float* B, long long ldb, int tB, float beta, const float* bias,
                             cudaStream_t stream);
#ifdef LSTMP_HAVE_TC_GEMM
cudaError_t launch_gemm_tc(float* C, long long ldc, int M, int N, int K, float alpha, const float* A, long long lda,
                           int tA, const float* B, long long ldb, int tB, float beta, const float* bias,
                           cudaStream_t stream, bool* handled, float* ws, size_t ws_floats, int* nlaunch);
#endif

// g[c] = sum over rows of diff[r][c] in a fixed order (AddRowSumMat).  32 columns x 8 row lanes per CTA: a warp reads
// 128 contiguous bytes of a row, lane j of a column sums rows j, j+8, ... and the 8 partials are added in order.
__global__ void __launch_bounds__(256) colsum_kernel(float* __restrict__ g, const float* __restrict__ diff, long long ld,
                                                     int rows, int cols) {
  __shared__ float part[8][33];
  const int cx = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float a0 = 0.f, a1 = 0.f;
  if (c < cols) {
    int r = j;
    for (; r + 8 < rows; r += 16) {
      a0 += diff[(long long)r * ld + c];
      a1 += diff[(long long)(r + 8) * ld + c];
    }
    if (r < rows) a0 += diff[(long long)r * ld + c];
  }
  part[j][cx] = a0 + a1;
  __syncthreads();
  if (j == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][cx];
    g[c] = t;
  }
}
}  // namespace lstmp

using namespace lstmp;

namespace {
int tfail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  lstmp::set_last_error(buf);
  return code;
}
#define T_TRY(expr)                                                                                       \
  do {                                                                                                    \
    cudaError_t e__ = (expr);                                                                             \
    if (e__ != cudaSuccess)                                                                               \
      return tfail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
}  // namespace

struct lstmp_b200_tail {
  int I = 0, P = 0, max_rows = 0, device = 0;
  size_t nparams = 0;
  float *params = nullptr, *corr = nullptr, *grads = nullptr;  // [W (P x I) | b (P)]
  float* diff = nullptr;                                       // [max_rows x P]: logits, then diff in place
  float* gemm_ws = nullptr;
  size_t gemm_ws_floats = 0;
  lstmp_b200_xent_handle_t xent = nullptr;
  HlWorkspace hlws;
  int gemm_backend = 2;
  int rows_last = 0;
  unsigned long long launches = 0;
};

static int tail_gemm(lstmp_b200_tail* h, float* C, long long ldc, int M, int N, int K, const float* A, long long lda,
                     int tA, const float* B, long long ldb, int tB, const float* bias, cudaStream_t st) {
#ifdef LSTMP_HAVE_TC_GEMM
  bool handled = false;
  int nl = 1;
  if (h->gemm_backend == 2) {
    T_TRY(launch_gemm_hl(&h->hlws, C, ldc, M, N, K, 1.f, A, lda, tA, B, ldb, tB, 0.f, bias, st, &handled, h->gemm_ws,
                         h->gemm_ws_floats, &nl, false));
    h->launches += nl;
    if (handled) return 0;
  }
  T_TRY(launch_gemm_tc(C, ldc, M, N, K, 1.f, A, lda, tA, B, ldb, tB, 0.f, bias, st, &handled, h->gemm_ws,
                       h->gemm_ws_floats, &nl));
  if (handled) {
    h->launches += nl;
    return 0;
  }
#endif
  T_TRY(launch_gemm_simt(C, ldc, M, N, K, 1.f, A, lda, tA, B, ldb, tB, 0.f, bias, st));
  h->launches++;
  return 0;
}

extern "C" int lstmp_b200_tail_destroy(lstmp_b200_tail_handle_t h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  float* bufs[] = {h->params, h->corr, h->grads, h->diff, h->gemm_ws};
  for (float* b : bufs)
    if (b) cudaFree(b);
  lstmp_b200_xent_destroy(h->xent);
  gemm_hl_free(&h->hlws);
  delete h;
  return 0;
}

extern "C" int lstmp_b200_tail_create(int input_dim, int num_pdf, int max_frames, int device,
                                      lstmp_b200_tail_handle_t* out) {
  if (!out) return tfail(LSTMP_B200_EINVAL, "out handle is NULL");
  *out = nullptr;
  if (input_dim <= 0 || num_pdf <= 0 || max_frames <= 0)
    return tfail(LSTMP_B200_EINVAL, "tail: dimensions must be positive (I=%d P=%d frames=%d)", input_dim, num_pdf, max_frames);
  if (input_dim % 4 || num_pdf % 4)
    return tfail(LSTMP_B200_EINVAL, "tail: input_dim and num_pdf must be multiples of 4 (I=%d P=%d)", input_dim, num_pdf);
  if ((size_t)num_pdf * sizeof(float) > 200 * 1024)
    return tfail(LSTMP_B200_EUNSUPPORTED, "tail: a row of %d logits does not fit in shared memory", num_pdf);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return tfail(LSTMP_B200_ENODEV, "no CUDA device: %s (this engine has no CPU path)", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return tfail(LSTMP_B200_EINVAL, "device %d out of range", device);
  T_TRY(cudaSetDevice(device));
  lstmp_b200_tail* h = new (std::nothrow) lstmp_b200_tail();
  if (!h) return tfail(LSTMP_B200_ENOMEM, "host allocation failed");
  h->I = input_dim; h->P = num_pdf; h->max_rows = max_frames; h->device = device;
  h->nparams = (size_t)num_pdf * input_dim + num_pdf;
  h->gemm_ws_floats = (size_t)4 << 20;
  {
    const char* v = getenv("LSTMP_B200_GEMM");
    if (v && *v) h->gemm_backend = atoi(v);
  }
  bool ok = cudaMalloc((void**)&h->params, h->nparams * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void**)&h->corr, h->nparams * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void**)&h->grads, h->nparams * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void**)&h->diff, (size_t)max_frames * num_pdf * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void**)&h->gemm_ws, h->gemm_ws_floats * sizeof(float)) == cudaSuccess &&
            cudaMemset(h->params, 0, h->nparams * sizeof(float)) == cudaSuccess &&
            cudaMemset(h->corr, 0, h->nparams * sizeof(float)) == cudaSuccess &&
            cudaMemset(h->grads, 0, h->nparams * sizeof(float)) == cudaSuccess;
  if (!ok) {
    e = cudaGetLastError();
    lstmp_b200_tail_destroy(h);
    return tfail(LSTMP_B200_ENOMEM, "tail: device allocation failed: %s", cudaGetErrorString(e));
  }
  int rc = lstmp_b200_xent_create(max_frames, device, &h->xent);
  if (rc) {
    lstmp_b200_tail_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

extern "C" int lstmp_b200_tail_arena(lstmp_b200_tail_handle_t h, int which, float** dev_ptr, size_t* count) {
  if (!h || !dev_ptr || !count) return tfail(LSTMP_B200_EINVAL, "NULL argument");
  float* a = which == 0 ? h->params : which == 1 ? h->corr : which == 2 ? h->grads : nullptr;
  if (!a) return tfail(LSTMP_B200_EINVAL, "bad arena %d", which);
  *dev_ptr = a;
  *count = h->nparams;
  return 0;
}

extern "C" int lstmp_b200_tail_set_flat(lstmp_b200_tail_handle_t h, int which, const float* src, void* stream) {
  float* a;
  size_t n;
  int rc = lstmp_b200_tail_arena(h, which, &a, &n);
  if (rc) return rc;
  if (!src) return tfail(LSTMP_B200_EINVAL, "NULL src");
  T_TRY(cudaSetDevice(h->device));
  T_TRY(cudaMemcpyAsync(a, src, n * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  T_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
extern "C" int lstmp_b200_tail_get_flat(lstmp_b200_tail_handle_t h, int which, float* dst, void* stream) {
  float* a;
  size_t n;
  int rc = lstmp_b200_tail_arena(h, which, &a, &n);
  if (rc) return rc;
  if (!dst) return tfail(LSTMP_B200_EINVAL, "NULL dst");
  T_TRY(cudaSetDevice(h->device));
  T_TRY(cudaMemcpyAsync(dst, a, n * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  T_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

extern "C" int lstmp_b200_tail_propagate_eval(lstmp_b200_tail_handle_t h, const float* in, size_t ld_in, int num_frames,
                                              const float* frame_mask_host, const int32_t* post_row_ptr_host,
                                              const int32_t* post_pdf_host, const float* post_weight_host,
                                              float* post_out, size_t ld_post, void* stream) {
  if (!h || !in || !frame_mask_host || !post_row_ptr_host) return tfail(LSTMP_B200_EINVAL, "NULL argument");
  if (num_frames <= 0 || num_frames > h->max_rows)
    return tfail(LSTMP_B200_EINVAL, "tail: %d frames, handle was created for at most %d", num_frames, h->max_rows);
  if (ld_in < (size_t)h->I) return tfail(LSTMP_B200_EINVAL, "stride < columns");
  T_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  // logits = in * W^T + b                                                (AffineTransform::PropagateFnc)
  int rc = tail_gemm(h, h->diff, h->P, num_frames, h->P, h->I, in, (long long)ld_in, 0, h->params, h->I, 1,
                     h->params + (size_t)h->P * h->I, st);
  if (rc) return rc;
  // softmax + masked xent, diff written in place over the logits         (Softmax::PropagateFnc + Xent::EvalMasked)
  rc = lstmp_b200_xent_eval_masked_logits(h->xent, frame_mask_host, h->diff, h->P, num_frames, h->P, post_row_ptr_host,
                                          post_pdf_host, post_weight_host, post_out, ld_post, h->diff, h->P, stream);
  if (rc) return rc;
  h->rows_last = num_frames;
  return 0;
}

extern "C" int lstmp_b200_tail_backpropagate(lstmp_b200_tail_handle_t h, const float* in, size_t ld_in, float* in_diff,
                                             size_t ld_id, int num_frames, void* stream) {
  if (!h || !in) return tfail(LSTMP_B200_EINVAL, "NULL argument");
  if (h->rows_last == 0) return tfail(LSTMP_B200_ESTATE, "tail: backpropagate without a preceding propagate_eval");
  if (num_frames != h->rows_last)
    return tfail(LSTMP_B200_EINVAL, "tail: backpropagate of %d frames but the last propagate_eval had %d", num_frames,
                 h->rows_last);
  if (ld_in < (size_t)h->I || (in_diff && ld_id < (size_t)h->I)) return tfail(LSTMP_B200_EINVAL, "stride < columns");
  T_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // in_diff = diff * W                                                     (AffineTransform::BackpropagateFnc)
  if (in_diff && (rc = tail_gemm(h, in_diff, (long long)ld_id, num_frames, h->I, h->P, h->diff, h->P, 0, h->params, h->I, 0,
                                 nullptr, st)))
    return rc;
  // G(W) = diff^T * in ; G(b) = column sums of diff                        (AffineTransform::Update, gradient part)
  if ((rc = tail_gemm(h, h->grads, h->I, h->P, h->I, num_frames, h->diff, h->P, 1, in, (long long)ld_in, 0, nullptr, st)))
    return rc;
  colsum_kernel<<<(h->P + 31) / 32, 256, 0, st>>>(h->grads + (size_t)h->P * h->I, h->diff, h->P, num_frames, h->P);
  T_TRY(cudaGetLastError());
  h->launches++;
  return 0;
}

extern "C" int lstmp_b200_tail_update(lstmp_b200_tail_handle_t h, float learn_rate, float momentum, void* stream) {
  if (!h) return tfail(LSTMP_B200_EINVAL, "NULL handle");
  T_TRY(cudaSetDevice(h->device));
  T_TRY(launch_update(h->params, h->corr, h->grads, h->nparams, learn_rate, momentum, 0.f, (cudaStream_t)stream));
  h->launches++;
  return 0;
}

typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
extern "C" int lstmp_b200_tail_allreduce_grads_nccl(lstmp_b200_tail_handle_t h, void* comm, void* stream) {
  if (!h || !comm) return tfail(LSTMP_B200_EINVAL, "NULL argument");
  T_TRY(cudaSetDevice(h->device));
  static nccl_allreduce_fn fn = nullptr;
  if (!fn) {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return tfail(LSTMP_B200_EUNSUPPORTED, "dlopen(libnccl.so.2): %s", dlerror());
    fn = (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce");
    if (!fn) return tfail(LSTMP_B200_EUNSUPPORTED, "ncclAllReduce not found");
  }
  int rc = fn(h->grads, h->grads, h->nparams, 7 /* ncclFloat32 */, 0 /* ncclSum */, comm, (cudaStream_t)stream);
  if (rc != 0) return tfail(LSTMP_B200_EUNSUPPORTED, "ncclAllReduce returned %d", rc);
  return 0;
}

extern "C" int lstmp_b200_tail_get_diff(lstmp_b200_tail_handle_t h, float* dst, size_t ld, void* stream) {
  if (!h || !dst) return tfail(LSTMP_B200_EINVAL, "NULL argument");
  if (h->rows_last == 0) return tfail(LSTMP_B200_ESTATE, "tail: no propagate_eval yet");
  if (ld < (size_t)h->P) return tfail(LSTMP_B200_EINVAL, "stride < columns");
  T_TRY(cudaSetDevice(h->device));
  T_TRY(cudaMemcpy2DAsync(dst, ld * sizeof(float), h->diff, (size_t)h->P * sizeof(float), (size_t)h->P * sizeof(float),
                          (size_t)h->rows_last, cudaMemcpyDefault, (cudaStream_t)stream));
  T_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

extern "C" int lstmp_b200_tail_get_stats(lstmp_b200_tail_handle_t h, lstmp_b200_xent_stats_t* out, void* stream) {
  if (!h || !out) return tfail(LSTMP_B200_EINVAL, "NULL argument");
  int rc = lstmp_b200_xent_get_stats(h->xent, out, stream);
  if (rc) return rc;
  out->kernel_launches += h->launches;
  return 0;
}
extern "C" int lstmp_b200_tail_reset_stats(lstmp_b200_tail_handle_t h, void* stream) {
  if (!h) return tfail(LSTMP_B200_EINVAL, "NULL handle");
  return lstmp_b200_xent_reset_stats(h->xent, stream);
}
