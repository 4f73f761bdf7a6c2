// Device-side multi-stream chunk assembly and the TimeShift row gather.
//
// The reference trainer (google/nnetbin/bd-nnet-train-lstm-streams.cc:128-212) fills every BPTT chunk on the host --
// a serial row-by-row CopyFromVec with the targets-delay shift and last-frame padding (:187-206) --, copies the
// [T*S x D] chunk to the device (:212, CuMatrix(feat)) and runs the feature transform (AddShift + Rescale,
// google/feature_transform.nnet.txt:2-5) as two more kernels, every chunk.  Here an utterance crosses PCIe ONCE, when
// a stream takes it (:152-170): pinned staging -> async H2D on an own copy stream into one of the stream's two device
// slots (double-buffered: the gather of the previous chunk may still be reading the other one).  Per chunk only
// 3*S ints go H2D (curt, lent, slot), and ONE kernel gathers row (t, s) = transform(utt_s[min(curt_s + t + delay,
// lent_s - 1)]) straight into the time-major chunk matrix the network reads.  Bit-exact with the host loop (same
// fp32 add, then multiply; copies otherwise).  The host bookkeeping (keys / targets / curt / lent / new_utt_flags /
// frame_mask) stays on the host, in kaldi/b200-stream-dispatch.h, exactly as in the reference.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/lstmp_b200.h"
#include "lstmp_common.cuh"

namespace lstmp {
void set_last_error(const char* msg);
}

namespace {
int dfail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  lstmp::set_last_error(buf);
  return code;
}
#define D_TRY(expr)                                                                                       \
  do {                                                                                                    \
    cudaError_t e__ = (expr);                                                                             \
    if (e__ != cudaSuccess)                                                                               \
      return dfail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

constexpr int kMetaRing = 64;     // pinned per-chunk descriptors in flight
constexpr int kStageRing = 8;     // pinned utterance staging buffers in flight

// meta: [0,S) curt, [S,2S) lent, [2S,3S) slot of each stream.  One thread per 16-byte unit of the chunk matrix.
__global__ void __launch_bounds__(256) assemble_kernel(float* __restrict__ feat, long long ld, int S, int T, int D4,
                                                       int delay, const float* __restrict__ pool,
                                                       long long slot_floats, const int* __restrict__ meta,
                                                       const float4* __restrict__ shift,
                                                       const float4* __restrict__ scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)T * S * D4;
  if (idx >= total) return;
  const int q = (int)(idx % D4);
  const int row = (int)(idx / D4);
  const int s = row % S, t = row / S;
  const int L = meta[S + s];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (L > 0) {
    const int cur = meta[s] + t;                                   // curt at this t (TRAIN.cc:204 increments every t)
    const int f = (cur + delay < L) ? cur + delay : L - 1;         // TRAIN.cc:198-202
    const float4* src = reinterpret_cast<const float4*>(pool + ((long long)(2 * s + meta[2 * S + s])) * slot_floats +
                                                        (long long)f * (4 * D4));
    v = __ldg(src + q);
    if (shift) {                                                   // <AddShift>: out = in + shift
      const float4 a = shift[q];
      v.x = __fadd_rn(v.x, a.x); v.y = __fadd_rn(v.y, a.y); v.z = __fadd_rn(v.z, a.z); v.w = __fadd_rn(v.w, a.w);
    }
    if (scale) {                                                   // <Rescale>: out = out * scale
      const float4 a = scale[q];
      v.x = __fmul_rn(v.x, a.x); v.y = __fmul_rn(v.y, a.y); v.z = __fmul_rn(v.z, a.z); v.w = __fmul_rn(v.w, a.w);
    }
  }
  *reinterpret_cast<float4*>(feat + (long long)row * ld + 4 * q) = v;
}

// out row dst = in row clamp(dst + shift, 0, rows - 1)   (standard/nnet/nnet-time-shift.h:42-51)
__global__ void __launch_bounds__(256) time_shift_kernel(float* __restrict__ out, long long ldo,
                                                         const float* __restrict__ in, long long ldi, int rows,
                                                         int cols, int shift) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int c = (int)(idx % cols), dst = (int)(idx / cols);
  int src = dst + shift;
  src = src < 0 ? 0 : src;
  src = src > rows - 1 ? rows - 1 : src;
  out[(long long)dst * ldo + c] = in[(long long)src * ldi + c];
}
}  // namespace

struct lstmp_b200_dispatch {
  int S = 0, T = 0, delay = 0, D = 0, max_frames = 0, device = 0;
  float* pool = nullptr;            // [S][2 slots][max_frames x D]
  float *shift = nullptr, *scale = nullptr;
  bool have_shift = false, have_scale = false;
  int* meta_dev = nullptr;          // [kMetaRing][3*S]
  int* meta_host = nullptr;         // pinned, same shape
  cudaEvent_t meta_ev[kMetaRing] = {nullptr};
  float* stage = nullptr;           // pinned [kStageRing][max_frames x D]
  cudaEvent_t stage_ev[kStageRing] = {nullptr};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t loaded = nullptr;     // all loads issued so far have landed (copy stream)
  cudaEvent_t asm_ev[kMetaRing] = {nullptr};  // assemble n has run (compute stream)
  std::vector<int> slot;            // active slot of each stream
  std::vector<long long> slot_last_read[2];   // last assemble that read slot k of stream s (-1: never)
  long long n_asm = 0, n_load = 0;
  unsigned long long launches = 0, h2d_bytes = 0;
  bool pending_loads = false;
};

extern "C" int lstmp_b200_dispatch_destroy(lstmp_b200_dispatch_handle_t h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (h->pool) cudaFree(h->pool);
  if (h->shift) cudaFree(h->shift);
  if (h->scale) cudaFree(h->scale);
  if (h->meta_dev) cudaFree(h->meta_dev);
  if (h->meta_host) cudaFreeHost(h->meta_host);
  if (h->stage) cudaFreeHost(h->stage);
  for (int i = 0; i < kMetaRing; ++i) {
    if (h->meta_ev[i]) cudaEventDestroy(h->meta_ev[i]);
    if (h->asm_ev[i]) cudaEventDestroy(h->asm_ev[i]);
  }
  for (int i = 0; i < kStageRing; ++i)
    if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
  if (h->loaded) cudaEventDestroy(h->loaded);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return 0;
}

extern "C" int lstmp_b200_dispatch_create(int num_stream, int batch_size, int targets_delay, int feat_dim,
                                          int max_utt_frames, int device, lstmp_b200_dispatch_handle_t* out) {
  if (!out) return dfail(LSTMP_B200_EINVAL, "out handle is NULL");
  *out = nullptr;
  if (num_stream <= 0 || batch_size <= 0 || targets_delay < 0 || feat_dim <= 0 || max_utt_frames <= 0)
    return dfail(LSTMP_B200_EINVAL, "dispatch: bad shape (S=%d T=%d delay=%d D=%d max_frames=%d)", num_stream, batch_size,
                 targets_delay, feat_dim, max_utt_frames);
  if (feat_dim % 4) return dfail(LSTMP_B200_EINVAL, "dispatch: feat_dim %d must be a multiple of 4", feat_dim);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return dfail(LSTMP_B200_ENODEV, "no CUDA device: %s (the chunk assembly has no CPU path)", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return dfail(LSTMP_B200_EINVAL, "device %d out of range", device);
  D_TRY(cudaSetDevice(device));
  lstmp_b200_dispatch* h = new (std::nothrow) lstmp_b200_dispatch();
  if (!h) return dfail(LSTMP_B200_ENOMEM, "host allocation failed");
  h->S = num_stream; h->T = batch_size; h->delay = targets_delay; h->D = feat_dim; h->max_frames = max_utt_frames;
  h->device = device;
  h->slot.assign(num_stream, 1);  // the first load of a stream flips to slot 0
  h->slot_last_read[0].assign(num_stream, -1);
  h->slot_last_read[1].assign(num_stream, -1);
  const size_t slot_floats = (size_t)max_utt_frames * feat_dim;
  bool ok = cudaMalloc((void**)&h->pool, (size_t)2 * num_stream * slot_floats * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void**)&h->shift, feat_dim * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void**)&h->scale, feat_dim * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void**)&h->meta_dev, (size_t)kMetaRing * 3 * num_stream * sizeof(int)) == cudaSuccess &&
            cudaMallocHost((void**)&h->meta_host, (size_t)kMetaRing * 3 * num_stream * sizeof(int)) == cudaSuccess &&
            cudaMallocHost((void**)&h->stage, (size_t)kStageRing * slot_floats * sizeof(float)) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->loaded, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < kMetaRing; ++i)
    ok = cudaEventCreateWithFlags(&h->meta_ev[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->asm_ev[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < kStageRing; ++i)
    ok = cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    e = cudaGetLastError();
    lstmp_b200_dispatch_destroy(h);
    return dfail(LSTMP_B200_ENOMEM, "dispatch: allocation failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return 0;
}

extern "C" int lstmp_b200_dispatch_set_transform(lstmp_b200_dispatch_handle_t h, const float* shift,
                                                 const float* scale) {
  if (!h) return dfail(LSTMP_B200_EINVAL, "NULL handle");
  D_TRY(cudaSetDevice(h->device));
  D_TRY(cudaDeviceSynchronize());
  h->have_shift = shift != nullptr;
  h->have_scale = scale != nullptr;
  if (shift) D_TRY(cudaMemcpy(h->shift, shift, h->D * sizeof(float), cudaMemcpyHostToDevice));
  if (scale) D_TRY(cudaMemcpy(h->scale, scale, h->D * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int lstmp_b200_dispatch_load_utt(lstmp_b200_dispatch_handle_t h, int stream, const float* feats,
                                            size_t ld, int num_frames) {
  if (!h || !feats) return dfail(LSTMP_B200_EINVAL, "NULL argument");
  if (stream < 0 || stream >= h->S) return dfail(LSTMP_B200_EINVAL, "dispatch: stream %d out of range", stream);
  if (num_frames <= 0 || num_frames > h->max_frames)
    return dfail(LSTMP_B200_EINVAL, "dispatch: utterance of %d frames (max_utt_frames=%d)", num_frames, h->max_frames);
  if (ld < (size_t)h->D) return dfail(LSTMP_B200_EINVAL, "stride < columns");
  D_TRY(cudaSetDevice(h->device));
  const size_t slot_floats = (size_t)h->max_frames * h->D;
  const int sg = (int)(h->n_load % kStageRing);
  if (h->n_load >= kStageRing) D_TRY(cudaEventSynchronize(h->stage_ev[sg]));  // staging buffer free again
  float* st = h->stage + (size_t)sg * slot_floats;
  for (int f = 0; f < num_frames; ++f) memcpy(st + (size_t)f * h->D, feats + (size_t)f * ld, h->D * sizeof(float));
  const int next = h->slot[stream] ^ 1;
  // the target slot was last read by an earlier assemble (the utterance before the current one): order the copy
  // after it.  The ring entry holds that assemble's event or a later one -- waiting for a later one is only stronger.
  const long long lr = h->slot_last_read[next][stream];
  if (lr >= 0) D_TRY(cudaStreamWaitEvent(h->copy_stream, h->asm_ev[lr % kMetaRing], 0));
  float* dst = h->pool + ((size_t)(2 * stream + next)) * slot_floats;
  D_TRY(cudaMemcpyAsync(dst, st, (size_t)num_frames * h->D * sizeof(float), cudaMemcpyHostToDevice, h->copy_stream));
  D_TRY(cudaEventRecord(h->stage_ev[sg], h->copy_stream));
  h->slot[stream] = next;
  h->pending_loads = true;
  h->n_load++;
  h->h2d_bytes += (unsigned long long)num_frames * h->D * sizeof(float);
  return 0;
}

extern "C" int lstmp_b200_dispatch_assemble(lstmp_b200_dispatch_handle_t h, const int32_t* curt, const int32_t* lent,
                                            float* feat, size_t ld_feat, void* stream) {
  if (!h || !curt || !lent || !feat) return dfail(LSTMP_B200_EINVAL, "NULL argument");
  if (ld_feat < (size_t)h->D || (ld_feat % 4)) return dfail(LSTMP_B200_EINVAL, "feat stride must be >= D and a multiple of 4");
  D_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int S = h->S;
  const int mi = (int)(h->n_asm % kMetaRing);
  if (h->n_asm >= kMetaRing) D_TRY(cudaEventSynchronize(h->meta_ev[mi]));  // descriptor slot free again
  int* mh = h->meta_host + (size_t)mi * 3 * S;
  for (int s = 0; s < S; ++s) {
    if (lent[s] < 0 || lent[s] > h->max_frames || curt[s] < 0)
      return dfail(LSTMP_B200_EINVAL, "dispatch: stream %d has curt=%d lent=%d", s, curt[s], lent[s]);
    mh[s] = curt[s];
    mh[S + s] = lent[s];
    mh[2 * S + s] = h->slot[s];
    if (lent[s] > 0) h->slot_last_read[h->slot[s]][s] = h->n_asm;
  }
  int* md = h->meta_dev + (size_t)mi * 3 * S;
  D_TRY(cudaMemcpyAsync(md, mh, (size_t)3 * S * sizeof(int), cudaMemcpyHostToDevice, st));
  D_TRY(cudaEventRecord(h->meta_ev[mi], st));
  if (h->pending_loads) {  // utterances taken since the last chunk must have landed before the gather reads them
    D_TRY(cudaEventRecord(h->loaded, h->copy_stream));
    D_TRY(cudaStreamWaitEvent(st, h->loaded, 0));
    h->pending_loads = false;
  }
  const int D4 = h->D / 4;
  const long long total = (long long)h->T * S * D4;
  assemble_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      feat, (long long)ld_feat, S, h->T, D4, h->delay, h->pool, (long long)h->max_frames * h->D, md,
      h->have_shift ? reinterpret_cast<const float4*>(h->shift) : nullptr,
      h->have_scale ? reinterpret_cast<const float4*>(h->scale) : nullptr);
  D_TRY(cudaGetLastError());
  D_TRY(cudaEventRecord(h->asm_ev[mi], st));
  h->n_asm++;
  h->launches++;
  h->h2d_bytes += (unsigned long long)3 * S * sizeof(int);
  return 0;
}

extern "C" int lstmp_b200_dispatch_get_stats(lstmp_b200_dispatch_handle_t h, lstmp_b200_dispatch_stats_t* out) {
  if (!h || !out) return dfail(LSTMP_B200_EINVAL, "NULL argument");
  out->kernel_launches = h->launches;
  out->h2d_bytes = h->h2d_bytes;
  out->utterances_loaded = (unsigned long long)h->n_load;
  out->chunks_assembled = (unsigned long long)h->n_asm;
  return 0;
}

extern "C" int lstmp_b200_time_shift(const float* in, size_t ld_in, float* out, size_t ld_out, int num_rows, int num_cols,
                                     int shift, void* stream) {
  if (!in || !out) return dfail(LSTMP_B200_EINVAL, "NULL in/out");
  if (num_rows < 0 || num_cols < 0 || ld_in < (size_t)num_cols || ld_out < (size_t)num_cols)
    return dfail(LSTMP_B200_EINVAL, "time_shift: bad shape");
  if (in == out) return dfail(LSTMP_B200_EINVAL, "time_shift: in-place is not supported (rows are gathered)");
  if (num_rows == 0 || num_cols == 0) return 0;
  const long long total = (long long)num_rows * num_cols;
  time_shift_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      out, (long long)ld_out, in, (long long)ld_in, num_rows, num_cols, shift);
  D_TRY(cudaGetLastError());
  return 0;
}
