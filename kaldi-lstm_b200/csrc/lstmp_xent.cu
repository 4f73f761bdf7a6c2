// Masked cross-entropy with SPARSE targets for sm_100a  (SURVEY.md section 8(f) rank 1).
//
// Replaces Xent::EvalMasked (reference google/nnet/nnet-loss.cc:76-164), which builds a dense
// [frames x num_pdf] fp32 target matrix on the host and copies it to the device every chunk (340 MB per chunk at
// 5120 x 16624), runs ~12 dense elementwise passes over it, and does two D2H argmax copies plus two blocking Sum()s.
// Here the posterior stays sparse (CSR: row_ptr / pdf / weight, a few bytes per frame), ONE kernel makes one pass
// over net_out (read y, write diff = mask*(y - t), first-max argmax on the fly), the handful of target entries of a
// row are applied by one thread, and the per-row statistics are reduced in a fixed order by a second tiny kernel
// into device-resident double accumulators that are only read back by lstmp_b200_xent_get_stats (Report()).
// HBM-bound: algorithmic bytes = 8 * frames * num_pdf (+ ~20 B per frame).
//
// Numerics follow the reference's CPU matrix path: t*log(y) and t*log(t + 1e-20) in fp32 per entry, summed in
// double; argmax = first column holding the row maximum (cu-matrix.cc:1333-1343); duplicate pdf-ids in one frame
// accumulate (nnet-loss.cc:93).  Where the reference's DENSE formulation yields NaN (softmax underflow y == 0 at a
// column with t == 0: 0 * log(0)), the sparse formulation contributes 0.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>

#include "../../include/lstmp_b200.h"

namespace lstmp {
void set_last_error(const char* msg);  // lstmp_engine.cu

namespace xent {
constexpr int kThreads = 256;

// Thread 0 of a row's CTA: applies the frame's handful of sparse target entries to diff and computes the row statistics.
// yr: the row's softmax outputs (global memory, or shared memory in the fused softmax kernel).
__device__ __forceinline__ void apply_targets(int r, const float* yr, float m, int num_pdf,
                                              const int* __restrict__ row_ptr, const int* __restrict__ pdf,
                                              const float* __restrict__ weight, float* dr, int best_id,
                                              float* __restrict__ row_xent, float* __restrict__ row_ent,
                                              int* __restrict__ row_correct) {
  // sparse targets of this frame: entries [beg, end); duplicates of a pdf accumulate      (:85-96)
  const int beg = row_ptr[r], end = row_ptr[r + 1];
  float xe = 0.f, en = 0.f;
  float tmax = -1e21f;  // maximum over the LISTED columns and its first column
  int tmax_id = -1;
  int first_free = 0;   // smallest column NOT in the list (holds an implicit 0 of the dense target row)
  for (int e = beg; e < end; ++e) {
    const int p = pdf[e];
    bool seen = false;
    for (int q = beg; q < e; ++q) seen |= (pdf[q] == p);
    if (seen) continue;
    float t = 0.f;
    for (int q = e; q < end; ++q)
      if (pdf[q] == p) t += weight[q];
    dr[p] = (yr[p] - t) * m;                          // diff = (y - t) * mask, same rounding     :103-106
    xe += (logf(yr[p]) * t) * m;                      // t * log(y), masked                     :123-128
    en += (logf(t + 1e-20f) * t) * m;                 // t * log(t + 1e-20), masked             :130-136
    if (t > tmax || (t == tmax && p < tmax_id)) {
      tmax = t;
      tmax_id = p;
    }
  }
  // smallest column not listed (lists are tiny; columns are tested in increasing order)
  for (bool again = true; again;) {
    again = false;
    for (int e = beg; e < end; ++e)
      if (pdf[e] == first_free) {
        ++first_free;
        again = true;
      }
  }
  // argmax of the dense target row: first column holding max(listed values, implicit zeros)
  int tgt_id;
  if (first_free >= num_pdf) {
    tgt_id = tmax_id;  // every column is listed
  } else if (tmax > 0.f) {
    tgt_id = tmax_id;
  } else if (tmax == 0.f) {
    tgt_id = tmax_id < first_free ? tmax_id : first_free;
  } else {
    tgt_id = first_free;  // all listed values negative (or the list is empty): the first implicit zero
  }
  row_xent[r] = xe;
  row_ent[r] = en;
  row_correct[r] = (m == 1.0f && tgt_id == best_id) ? 1 : 0;                              // :117-121
}

// One CTA per frame (row).
__global__ void __launch_bounds__(kThreads) xent_rows_kernel(const float* __restrict__ y, long long ld_y, int num_pdf,
                                                             const int* __restrict__ row_ptr,
                                                             const int* __restrict__ pdf,
                                                             const float* __restrict__ weight,
                                                             const float* __restrict__ mask, float* __restrict__ diff,
                                                             long long ld_d, float* __restrict__ row_xent,
                                                             float* __restrict__ row_ent,
                                                             int* __restrict__ row_correct, int vec) {
  const int r = blockIdx.x, tid = threadIdx.x;
  const float m = mask[r];
  const float* yr = y + (size_t)r * ld_y;
  float* dr = diff + (size_t)r * ld_d;

  // dense pass: diff = mask * y, running first-max argmax of y                          (nnet-loss.cc:103-111)
  float best = -1e21f;
  int best_id = -1;
  const int n4 = vec ? (num_pdf >> 2) : 0;  // 128-bit part (rows 16-byte aligned, checked on the host)
  for (int c = tid; c < n4; c += kThreads) {
    const float4 v = reinterpret_cast<const float4*>(yr)[c];
    reinterpret_cast<float4*>(dr)[c] = make_float4(v.x * m, v.y * m, v.z * m, v.w * m);
    if (best < v.x) { best = v.x; best_id = 4 * c; }       // columns in increasing order: keeps the first maximum
    if (best < v.y) { best = v.y; best_id = 4 * c + 1; }
    if (best < v.z) { best = v.z; best_id = 4 * c + 2; }
    if (best < v.w) { best = v.w; best_id = 4 * c + 3; }
  }
  for (int c = 4 * n4 + tid; c < num_pdf; c += kThreads) {
    const float v = yr[c];
    dr[c] = v * m;
    if (best < v) {  // columns visited in increasing order per thread: keeps the first maximum
      best = v;
      best_id = c;
    }
  }
  // block arg-max with "smaller column wins ties" = first maximum of the row
  __shared__ float s_val[kThreads / 32];
  __shared__ int s_id[kThreads / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_down_sync(0xffffffffu, best, off);
    const int oi = __shfl_down_sync(0xffffffffu, best_id, off);
    if (oi >= 0 && (ov > best || (ov == best && (best_id < 0 || oi < best_id)))) {
      best = ov;
      best_id = oi;
    }
  }
  if ((tid & 31) == 0) {
    s_val[tid >> 5] = best;
    s_id[tid >> 5] = best_id;
  }
  __syncthreads();  // also: every dr[c] of this row is written before thread 0 touches the target columns
  if (tid == 0) {
    for (int w = 1; w < kThreads / 32; ++w) {
      const float ov = s_val[w];
      const int oi = s_id[w];
      if (oi >= 0 && (ov > best || (ov == best && (best_id < 0 || oi < best_id)))) {
        best = ov;
        best_id = oi;
      }
    }
    apply_targets(r, yr, m, num_pdf, row_ptr, pdf, weight, dr, best_id, row_xent, row_ent, row_correct);
  }
}


// Block-wide arg-max reduction with "smaller column wins ties" (= first maximum of the row); result valid in thread 0.
__device__ __forceinline__ void block_first_max(float& best, int& best_id) {
  __shared__ float s_val[kThreads / 32];
  __shared__ int s_id[kThreads / 32];
  const int tid = threadIdx.x;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_down_sync(0xffffffffu, best, off);
    const int oi = __shfl_down_sync(0xffffffffu, best_id, off);
    if (oi >= 0 && (ov > best || (ov == best && (best_id < 0 || oi < best_id)))) {
      best = ov;
      best_id = oi;
    }
  }
  if ((tid & 31) == 0) {
    s_val[tid >> 5] = best;
    s_id[tid >> 5] = best_id;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kThreads / 32; ++w) {
      const float ov = s_val[w];
      const int oi = s_id[w];
      if (oi >= 0 && (ov > best || (ov == best && (best_id < 0 || oi < best_id)))) {
        best = ov;
        best_id = oi;
      }
    }
  }
}
__device__ __forceinline__ float block_sum_fixed(float v) {  // fixed-order tree: deterministic
  __shared__ float s_sum[kThreads / 32];
  __shared__ float s_tot;
  const int tid = threadIdx.x;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((tid & 31) == 0) s_sum[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += s_sum[w];
    s_tot = t;
  }
  __syncthreads();
  return s_tot;
}
__device__ __forceinline__ float block_max(float v) {
  __shared__ float s_mx[kThreads / 32];
  __shared__ float s_m;
  const int tid = threadIdx.x;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, off));
  if ((tid & 31) == 0) s_mx[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    float t = s_mx[0];
    for (int w = 1; w < kThreads / 32; ++w) t = fmaxf(t, s_mx[w]);
    s_m = t;
  }
  __syncthreads();
  return s_m;
}

// Softmax + masked cross-entropy of one frame per CTA, fused (the Softmax component +
// Xent::EvalMasked of the 512 -> 16624 tail).  The row of logits is read from HBM ONCE (128-bit loads) into shared
// memory; max, exp / sum and the normalisation run there (Kaldi's ApplySoftMaxPerRow: subtract the row maximum, Exp,
// scale by 1/sum); then diff = mask * (y - t) is written (128-bit stores), optionally the posteriors y too, with the
// same first-maximum arg-max and sparse-target logic as xent_rows_kernel.  diff may alias the logits (in place).
__global__ void __launch_bounds__(kThreads) softmax_xent_rows_kernel(
    const float* __restrict__ a, long long ld_a, int num_pdf, const int* __restrict__ row_ptr,
    const int* __restrict__ pdf, const float* __restrict__ weight, const float* __restrict__ mask, float* post,
    long long ld_p, float* diff, long long ld_d, float* __restrict__ row_xent, float* __restrict__ row_ent,
    int* __restrict__ row_correct, int vec) {
  extern __shared__ __align__(16) float srow[];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float m = mask[r];
  const float* ar = a + (size_t)r * ld_a;
  float* dr = diff + (size_t)r * ld_d;
  float* pr = post ? post + (size_t)r * ld_p : nullptr;
  const int n4 = vec ? (num_pdf >> 2) : 0;  // 128-bit part (rows 16-byte aligned, checked on the host)
  float mx = -INFINITY;
  for (int c = tid; c < n4; c += kThreads) {
    const float4 v = reinterpret_cast<const float4*>(ar)[c];
    reinterpret_cast<float4*>(srow)[c] = v;
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int c = 4 * n4 + tid; c < num_pdf; c += kThreads) {
    const float v = ar[c];
    srow[c] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_max(mx);  // (contains the __syncthreads that publishes srow)
  float sum = 0.f;
  for (int c = tid; c < num_pdf; c += kThreads) {
    const float e = expf(srow[c] - mx);
    srow[c] = e;
    sum += e;
  }
  sum = block_sum_fixed(sum);
  const float inv = 1.0f / sum;
  float best = -1e21f;
  int best_id = -1;
  for (int c = tid; c < n4; c += kThreads) {
    float4 y = reinterpret_cast<float4*>(srow)[c];
    y.x *= inv; y.y *= inv; y.z *= inv; y.w *= inv;
    reinterpret_cast<float4*>(srow)[c] = y;
    if (pr) reinterpret_cast<float4*>(pr)[c] = y;
    reinterpret_cast<float4*>(dr)[c] = make_float4(y.x * m, y.y * m, y.z * m, y.w * m);
    if (best < y.x) { best = y.x; best_id = 4 * c; }       // columns in increasing order: keeps the first maximum
    if (best < y.y) { best = y.y; best_id = 4 * c + 1; }
    if (best < y.z) { best = y.z; best_id = 4 * c + 2; }
    if (best < y.w) { best = y.w; best_id = 4 * c + 3; }
  }
  for (int c = 4 * n4 + tid; c < num_pdf; c += kThreads) {
    const float y = srow[c] * inv;
    srow[c] = y;
    if (pr) pr[c] = y;
    dr[c] = y * m;
    if (best < y) {
      best = y;
      best_id = c;
    }
  }
  block_first_max(best, best_id);  // (its __syncthreads also orders every dr[c] / srow[c] before thread 0 goes on)
  if (tid == 0)
    apply_targets(r, srow, m, num_pdf, row_ptr, pdf, weight, dr, best_id, row_xent, row_ent, row_correct);
}

// Fixed-order reduction of the per-row statistics into the accumulators (deterministic, no atomics).
// acc[0] += -sum xent, acc[1] += -sum ent ; cnt[0] += correct, cnt[1] += (int) sum(mask)     :138-142
__global__ void __launch_bounds__(kThreads) xent_reduce_kernel(int rows, const float* __restrict__ row_xent,
                                                               const float* __restrict__ row_ent,
                                                               const int* __restrict__ row_correct,
                                                               const float* __restrict__ mask, double* acc,
                                                               long long* cnt) {
  __shared__ double s[4][kThreads];
  const int tid = threadIdx.x;
  double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
  for (int r = tid; r < rows; r += kThreads) {
    a += (double)row_xent[r];
    b += (double)row_ent[r];
    c += (double)row_correct[r];
    d += (double)mask[r];
  }
  s[0][tid] = a;
  s[1][tid] = b;
  s[2][tid] = c;
  s[3][tid] = d;
  __syncthreads();
  for (int off = kThreads / 2; off > 0; off >>= 1) {
    if (tid < off) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k][tid] += s[k][tid + off];
    }
    __syncthreads();
  }
  if (tid == 0) {
    acc[0] -= s[0][0];
    acc[1] -= s[1][0];
    cnt[0] += (long long)(s[2][0] + 0.5);
    cnt[1] += (long long)(int)(float)s[3][0];  // int32 valid_frames = (int32)frame_mask_.Sum()
  }
}
}  // namespace xent
}  // namespace lstmp

using namespace lstmp;

struct lstmp_b200_xent {
  int device = 0, max_rows = 0;
  size_t cap_entries = 0;
  // device
  int *row_ptr = nullptr, *pdf = nullptr, *row_correct = nullptr;
  float *weight = nullptr, *mask = nullptr, *row_xent = nullptr, *row_ent = nullptr;
  double* acc = nullptr;     // [0] loss_, [1] entropy_
  long long* cnt = nullptr;  // [0] correct_, [1] frames_
  // pinned host staging (the posterior and the mask are host objects in Kaldi)
  int *h_row_ptr = nullptr, *h_pdf = nullptr;
  float *h_weight = nullptr, *h_mask = nullptr;
  cudaEvent_t staged = nullptr;  // the staging buffers may be rewritten once this event has completed
  unsigned long long launches = 0;
};

static int xfail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  char buf[512];
  snprintf(buf, sizeof buf, fmt, a, b, c);
  lstmp::set_last_error(buf);
  return code;
}
#define XCUDA(expr)                                                                              \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) return xfail((int)e__, "%s (xent, line %lld)", cudaGetErrorString(e__), __LINE__); \
  } while (0)

static int ensure_entries(lstmp_b200_xent* h, size_t n) {
  if (n <= h->cap_entries) return 0;
  size_t cap = h->cap_entries ? h->cap_entries : 1024;
  while (cap < n) cap *= 2;
  XCUDA(cudaDeviceSynchronize());  // earlier launches may still read the old buffers
  if (h->pdf) cudaFree(h->pdf);
  if (h->weight) cudaFree(h->weight);
  if (h->h_pdf) cudaFreeHost(h->h_pdf);
  if (h->h_weight) cudaFreeHost(h->h_weight);
  h->pdf = nullptr; h->weight = nullptr; h->h_pdf = nullptr; h->h_weight = nullptr;
  h->cap_entries = 0;
  XCUDA(cudaMalloc((void**)&h->pdf, cap * sizeof(int)));
  XCUDA(cudaMalloc((void**)&h->weight, cap * sizeof(float)));
  XCUDA(cudaMallocHost((void**)&h->h_pdf, cap * sizeof(int)));
  XCUDA(cudaMallocHost((void**)&h->h_weight, cap * sizeof(float)));
  h->cap_entries = cap;
  return 0;
}

extern "C" int lstmp_b200_xent_destroy(lstmp_b200_xent_handle_t h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  void* dev[] = {h->row_ptr, h->pdf, h->row_correct, h->weight, h->mask, h->row_xent, h->row_ent, h->acc, h->cnt};
  for (void* p : dev)
    if (p) cudaFree(p);
  void* host[] = {h->h_row_ptr, h->h_pdf, h->h_weight, h->h_mask};
  for (void* p : host)
    if (p) cudaFreeHost(p);
  if (h->staged) cudaEventDestroy(h->staged);
  delete h;
  return 0;
}

extern "C" int lstmp_b200_xent_create(int max_frames, int device, lstmp_b200_xent_handle_t* out) {
  if (!out) return xfail(LSTMP_B200_EINVAL, "xent_create: out handle is NULL");
  *out = nullptr;
  if (max_frames <= 0) return xfail(LSTMP_B200_EINVAL, "xent_create: max_frames must be positive%s (got %lld)", "", max_frames);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return xfail(LSTMP_B200_ENODEV, "no CUDA device: %s (this engine has no CPU path)", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return xfail(LSTMP_B200_EINVAL, "xent_create: device out of range%s (%lld)", "", device);
  XCUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  XCUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return xfail(LSTMP_B200_ENODEV, "device is not sm_100%s; this library contains sm_100a code only");
  lstmp_b200_xent* h = new (std::nothrow) lstmp_b200_xent();
  if (!h) return xfail(LSTMP_B200_ENOMEM, "host allocation failed");
  h->device = device;
  h->max_rows = max_frames;
  const size_t R = (size_t)max_frames;
  cudaError_t rc = cudaSuccess;
  auto ok = [&](cudaError_t x) {
    if (rc == cudaSuccess) rc = x;
  };
  ok(cudaMalloc((void**)&h->row_ptr, (R + 1) * sizeof(int)));
  ok(cudaMalloc((void**)&h->row_correct, R * sizeof(int)));
  ok(cudaMalloc((void**)&h->mask, R * sizeof(float)));
  ok(cudaMalloc((void**)&h->row_xent, R * sizeof(float)));
  ok(cudaMalloc((void**)&h->row_ent, R * sizeof(float)));
  ok(cudaMalloc((void**)&h->acc, 2 * sizeof(double)));
  ok(cudaMalloc((void**)&h->cnt, 2 * sizeof(long long)));
  ok(cudaMallocHost((void**)&h->h_row_ptr, (R + 1) * sizeof(int)));
  ok(cudaMallocHost((void**)&h->h_mask, R * sizeof(float)));
  ok(cudaEventCreateWithFlags(&h->staged, cudaEventDisableTiming));
  if (rc == cudaSuccess) rc = cudaMemset(h->acc, 0, 2 * sizeof(double));
  if (rc == cudaSuccess) rc = cudaMemset(h->cnt, 0, 2 * sizeof(long long));
  if (rc != cudaSuccess) {
    lstmp_b200_xent_destroy(h);
    return xfail(LSTMP_B200_ENOMEM, "xent_create: %s", cudaGetErrorString(rc));
  }
  int r2 = ensure_entries(h, (size_t)max_frames);
  if (r2) {
    lstmp_b200_xent_destroy(h);
    return r2;
  }
  XCUDA(cudaDeviceSynchronize());
  *out = h;
  return 0;
}

static int eval_common(lstmp_b200_xent_handle_t h, const float* frame_mask_host, const float* net_out, size_t ld_out,
                       int num_frames, int num_pdf, const int32_t* post_row_ptr_host, const int32_t* post_pdf_host,
                       const float* post_weight_host, float* diff, size_t ld_diff, void* stream, bool logits,
                       float* post_out, size_t ld_post) {
  if (!h) return xfail(LSTMP_B200_EINVAL, "xent_eval_masked: NULL handle");
  if (!frame_mask_host || !net_out || !post_row_ptr_host || !diff)
    return xfail(LSTMP_B200_EINVAL, "xent_eval_masked: NULL argument");
  if (num_frames <= 0 || num_frames > h->max_rows)
    return xfail(LSTMP_B200_EINVAL, "xent_eval_masked: %s%lld frames, handle was created for at most %lld", "", num_frames,
                 h->max_rows);
  if (num_pdf <= 0 || ld_out < (size_t)num_pdf || ld_diff < (size_t)num_pdf)
    return xfail(LSTMP_B200_EINVAL, "xent_eval_masked: bad num_pdf / leading dimension%s (num_pdf %lld)", "", num_pdf);
  if (!logits && (const float*)diff == net_out)  // the target columns of net_out are read after diff has been written
    return xfail(LSTMP_B200_EINVAL, "xent_eval_masked: diff must not alias net_out (use the _logits entry point in place)");
  if (post_out && ld_post < (size_t)num_pdf) return xfail(LSTMP_B200_EINVAL, "xent_eval_masked: bad posterior stride");
  if (post_row_ptr_host[0] != 0) return xfail(LSTMP_B200_EINVAL, "xent_eval_masked: post_row_ptr[0] must be 0");
  const long long nnz = post_row_ptr_host[num_frames];
  for (int t = 0; t < num_frames; ++t)
    if (post_row_ptr_host[t + 1] < post_row_ptr_host[t])
      return xfail(LSTMP_B200_EINVAL, "xent_eval_masked: post_row_ptr not monotone%s at frame %lld", "", t);
  if (nnz > 0 && (!post_pdf_host || !post_weight_host)) return xfail(LSTMP_B200_EINVAL, "xent_eval_masked: NULL posterior");
  for (long long e = 0; e < nnz; ++e)
    if (post_pdf_host[e] >= num_pdf || post_pdf_host[e] < 0)  // KALDI_ERR, nnet-loss.cc:88-91
      return xfail(LSTMP_B200_EINVAL,
                   "Posterior pdf-id out of NN-output dimension%s: nn-outputs %lld, posterior pdf-id %lld", "", num_pdf,
                   post_pdf_host[e]);
  XCUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  XCUDA(cudaEventSynchronize(h->staged));  // previous call's H2D copies have consumed the staging buffers
  int rc = ensure_entries(h, (size_t)(nnz > 0 ? nnz : 1));
  if (rc) return rc;
  memcpy(h->h_row_ptr, post_row_ptr_host, ((size_t)num_frames + 1) * sizeof(int));
  memcpy(h->h_mask, frame_mask_host, (size_t)num_frames * sizeof(float));
  if (nnz > 0) {
    memcpy(h->h_pdf, post_pdf_host, (size_t)nnz * sizeof(int));
    memcpy(h->h_weight, post_weight_host, (size_t)nnz * sizeof(float));
  }
  XCUDA(cudaMemcpyAsync(h->row_ptr, h->h_row_ptr, ((size_t)num_frames + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  XCUDA(cudaMemcpyAsync(h->mask, h->h_mask, (size_t)num_frames * sizeof(float), cudaMemcpyHostToDevice, st));
  if (nnz > 0) {
    XCUDA(cudaMemcpyAsync(h->pdf, h->h_pdf, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    XCUDA(cudaMemcpyAsync(h->weight, h->h_weight, (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  XCUDA(cudaEventRecord(h->staged, st));
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int vec = (al16(net_out) && al16(diff) && !(ld_out & 3) && !(ld_diff & 3) &&
                   (!post_out || (al16(post_out) && !(ld_post & 3)))) ? 1 : 0;
  if (logits) {
    const size_t smem = (size_t)num_pdf * sizeof(float);
    if (smem > 200 * 1024)
      return xfail(LSTMP_B200_EUNSUPPORTED, "fused softmax: a row of %s%lld logits does not fit in shared memory", "", num_pdf);
    static size_t cur[64] = {0};
    if (smem > cur[h->device & 63]) {
      XCUDA(cudaFuncSetAttribute((const void*)xent::softmax_xent_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
      cur[h->device & 63] = smem;
    }
    xent::softmax_xent_rows_kernel<<<num_frames, xent::kThreads, smem, st>>>(
        net_out, (long long)ld_out, num_pdf, h->row_ptr, h->pdf, h->weight, h->mask, post_out, (long long)ld_post, diff,
        (long long)ld_diff, h->row_xent, h->row_ent, h->row_correct, vec);
  } else {
    xent::xent_rows_kernel<<<num_frames, xent::kThreads, 0, st>>>(net_out, (long long)ld_out, num_pdf, h->row_ptr, h->pdf,
                                                                  h->weight, h->mask, diff, (long long)ld_diff,
                                                                  h->row_xent, h->row_ent, h->row_correct, vec);
  }
  XCUDA(cudaGetLastError());
  xent::xent_reduce_kernel<<<1, xent::kThreads, 0, st>>>(num_frames, h->row_xent, h->row_ent, h->row_correct, h->mask,
                                                         h->acc, h->cnt);
  XCUDA(cudaGetLastError());
  h->launches += 2;
  return 0;
}

extern "C" int lstmp_b200_xent_eval_masked(lstmp_b200_xent_handle_t h, const float* frame_mask_host,
                                           const float* net_out, size_t ld_out, int num_frames, int num_pdf,
                                           const int32_t* post_row_ptr_host, const int32_t* post_pdf_host,
                                           const float* post_weight_host, float* diff, size_t ld_diff, void* stream) {
  return eval_common(h, frame_mask_host, net_out, ld_out, num_frames, num_pdf, post_row_ptr_host, post_pdf_host,
                     post_weight_host, diff, ld_diff, stream, false, nullptr, 0);
}

extern "C" int lstmp_b200_xent_eval_masked_logits(lstmp_b200_xent_handle_t h, const float* frame_mask_host,
                                                  const float* logits, size_t ld_logits, int num_frames, int num_pdf,
                                                  const int32_t* post_row_ptr_host, const int32_t* post_pdf_host,
                                                  const float* post_weight_host, float* post_out, size_t ld_post,
                                                  float* diff, size_t ld_diff, void* stream) {
  return eval_common(h, frame_mask_host, logits, ld_logits, num_frames, num_pdf, post_row_ptr_host, post_pdf_host,
                     post_weight_host, diff, ld_diff, stream, true, post_out, ld_post);
}

extern "C" int lstmp_b200_xent_get_stats(lstmp_b200_xent_handle_t h, lstmp_b200_xent_stats_t* out, void* stream) {
  if (!h || !out) return xfail(LSTMP_B200_EINVAL, "xent_get_stats: NULL argument");
  XCUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  double acc[2];
  long long cnt[2];
  XCUDA(cudaMemcpyAsync(acc, h->acc, sizeof acc, cudaMemcpyDeviceToHost, st));
  XCUDA(cudaMemcpyAsync(cnt, h->cnt, sizeof cnt, cudaMemcpyDeviceToHost, st));
  XCUDA(cudaStreamSynchronize(st));
  out->loss = acc[0];
  out->entropy = acc[1];
  out->correct = cnt[0];
  out->frames = cnt[1];
  out->kernel_launches = h->launches;
  return 0;
}

extern "C" int lstmp_b200_xent_reset_stats(lstmp_b200_xent_handle_t h, void* stream) {
  if (!h) return xfail(LSTMP_B200_EINVAL, "xent_reset_stats: NULL handle");
  XCUDA(cudaSetDevice(h->device));
  XCUDA(cudaMemsetAsync(h->acc, 0, 2 * sizeof(double), (cudaStream_t)stream));
  XCUDA(cudaMemsetAsync(h->cnt, 0, 2 * sizeof(long long), (cudaStream_t)stream));
  return 0;
}
