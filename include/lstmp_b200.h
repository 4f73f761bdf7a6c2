/*
 * lstmp_b200.h -- C ABI of the B200-native LstmProjectedStreams engine (sm_100a).
 *
 * This is the drop-in boundary for ONE path of dophist/kaldi-lstm: the multi-stream
 * projected LSTM layer's forward / truncated-BPTT backward / SGD update.  Each entry
 * point replaces what the reference's component does through CuMatrix calls
 * (citations are relative to the reference tree, google/nnet/bd-nnet-lstm-projected-streams.h
 * = "LPS.h"):
 *
 *   lstmp_b200_create              ctor + InitData/ReadData buffer allocation   LPS.h:27-33,76-97,119-130
 *   lstmp_b200_set_params/get_...  ReadData / WriteData / GetParams            LPS.h:101-189
 *   lstmp_b200_reset               Reset(std::vector<int>&)                    LPS.h:212-220
 *   lstmp_b200_propagate           PropagateFnc(in, out)                       LPS.h:222-332
 *   lstmp_b200_backpropagate       BackpropagateFnc(in, out, out_diff, in_diff) LPS.h:334-499
 *   lstmp_b200_update              momentum accumulation + Update()            LPS.h:465-487,501-512
 *
 * and, one row of SURVEY.md section 8(f) at a time, the callers and data formats either side of that layer:
 *
 *   lstmp_b200_xent_*              Xent::EvalMasked + Report (sparse targets)  google/nnet/nnet-loss.cc:76-164,293-307
 *   lstmp_b200_tail_*              AffineTransform + Softmax + EvalMasked      google/nnet.proto:4-5, TRAIN.cc:215-228
 *   lstmp_b200_dispatch_*          multi-stream chunk assembly + CMVN           TRAIN.cc:146-212, feature_transform.nnet.txt
 *   lstmp_b200_time_shift, lstmp_b200_update_clipped     the standard/ version's TimeShift and gradient clip
 *   lstmp_b200_allreduce_grads_nccl, lstmp_b200_set_nccl stream-sharded data parallelism (new relative to the reference)
 *
 * The kernels fused behind these calls also replace google/cudamatrix/bd-cu-kernels.cu
 * (cudaF_add_mat_diag_vec / cudaF_add_mat_dot_mat, bd-cu-kernels-ansi.h:9-22) and the
 * CuMatrixBase methods listed in SURVEY.md section 8a (a6-a10).
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success, a negative
 *     LSTMP_B200_E* code on a usage error, or a positive cudaError_t value on a CUDA
 *     failure.  lstmp_b200_last_error() gives a message for the calling thread.
 *   - matrices are row-major fp32 with a leading dimension (row stride) in floats,
 *     exactly Kaldi's CuMatrixBase {data_, num_cols_, num_rows_, stride_} layout
 *     (google/cudamatrix/cu-matrix.h:479-489).  `in`, `out`, `out_diff`, `in_diff` are
 *     DEVICE pointers owned by the caller (Nnet's propagate/backpropagate buffers).
 *   - frame rows are time-major: row = t*S + s (bd-nnet-train-lstm-streams.cc:187-206).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is
 *     what Kaldi's CuDevice uses).
 *   - there is NO CPU fallback: every entry point needs a CUDA device of compute
 *     capability 10.x and fails loudly otherwise.
 */
#ifndef LSTMP_B200_H_
#define LSTMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lstmp_b200_engine* lstmp_b200_handle_t;

enum {
  LSTMP_B200_OK = 0,
  LSTMP_B200_EINVAL = -1,      /* bad argument / shape (KALDI_ASSERT in the reference) */
  LSTMP_B200_ENODEV = -2,      /* no sm_100 device */
  LSTMP_B200_ENOMEM = -3,      /* shape does not fit on-chip / device memory */
  LSTMP_B200_ESTATE = -4,      /* call order violated (e.g. backpropagate without propagate) */
  LSTMP_B200_EUNSUPPORTED = -5 /* NCCL entry point without NCCL at run time, etc. */
};

/* ABI version of this header (bumped on incompatible change). */
int lstmp_b200_abi_version(void);
const char* lstmp_b200_last_error(void);

/* input_dim I, cell_dim C (<CellDim>), recur_dim R (= OutputDim, LPS.h:30), num_stream S
 * (<NumStream>), max_frames = largest T (BPTT chunk length) a later call will use.
 * Requires I % 4 == 0, C % 4 == 0, R % 4 == 0 (128-bit vector / TMA bulk granularity).
 * Parameters start at zero; gradient/momentum buffers and the carried state start at zero
 * (LPS.h:76,89-97).  When the layer's weight slices fit in shared memory across the SMs (800/512 does) the
 * persistent per-chunk kernels are used; otherwise (e.g. 2048/1024) the engine falls back to a weights-streamed
 * per-timestep CUDA path with the same results (lstmp_b200_info_t.weights_streamed). */
int lstmp_b200_create(int input_dim, int cell_dim, int recur_dim, int num_stream, int max_frames, int device,
                      lstmp_b200_handle_t* out);
int lstmp_b200_destroy(lstmp_b200_handle_t h);
/* Deep copy: parameters, momentum buffers and carried state (Component::Copy, LPS.h:38). */
int lstmp_b200_clone(lstmp_b200_handle_t src, lstmp_b200_handle_t* out);

/* NumParams() (LPS.h:152-160). */
int lstmp_b200_num_params(lstmp_b200_handle_t h, size_t* n);

/* Parameters as the seven tensors of LPS.h:590-613.  Pointers may be host or device
 * (cudaMemcpyDefault); ld* are row strides in floats.
 *   w_gifo_x [4C x I], w_gifo_r [4C x R], bias [4C], peephole_{i,f,o}_c [C], w_r_m [R x C]
 * Gate order along 4C is g,i,f,o (LPS.h:234-243). */
int lstmp_b200_set_params(lstmp_b200_handle_t h, const float* w_gifo_x, size_t ld_x, const float* w_gifo_r,
                          size_t ld_r, const float* bias, const float* peephole_i_c, const float* peephole_f_c,
                          const float* peephole_o_c, const float* w_r_m, size_t ld_m, void* stream);
int lstmp_b200_get_params(lstmp_b200_handle_t h, float* w_gifo_x, size_t ld_x, float* w_gifo_r, size_t ld_r,
                          float* bias, float* peephole_i_c, float* peephole_f_c, float* peephole_o_c,
                          float* w_r_m, size_t ld_m, void* stream);
/* Flat vector in GetParams() order (LPS.h:162-189); host or device pointer.
 * which: 0 = parameters, 1 = momentum-accumulated gradients (*_corr_), 2 = fresh gradient of the
 * last backpropagate (before momentum / all-reduce). */
int lstmp_b200_get_flat(lstmp_b200_handle_t h, int which, float* dst, void* stream);
int lstmp_b200_set_flat(lstmp_b200_handle_t h, int which, const float* src, void* stream);
/* Device address + length of the same three arenas, e.g. to all-reduce the fresh gradient with the
 * caller's own communicator between backpropagate and update. */
int lstmp_b200_arena(lstmp_b200_handle_t h, int which, float** dev_ptr, size_t* count);

/* Carried state prev_nnet_state_ (LPS.h:583): only the c and r column blocks are ever read
 * (LPS.h:275-281), so only those are exposed.  c [S x C], r [S x R]; host or device pointers. */
int lstmp_b200_get_state(lstmp_b200_handle_t h, float* c, size_t ld_c, float* r, size_t ld_r, void* stream);
int lstmp_b200_set_state(lstmp_b200_handle_t h, const float* c, size_t ld_c, const float* r, size_t ld_r,
                         void* stream);

/* Reset (LPS.h:212-220): host_flags[s] == 1 zeroes stream s's carried state.  n must equal S. */
int lstmp_b200_reset(lstmp_b200_handle_t h, const int32_t* host_flags, int n, void* stream);

/* PropagateFnc (LPS.h:222-332).  in [num_rows x I], out [num_rows x R]; num_rows = T*S with
 * T <= max_frames, else LSTMP_B200_EINVAL (the reference asserts, LPS.h:225). */
int lstmp_b200_propagate(lstmp_b200_handle_t h, const float* in, size_t ld_in, float* out, size_t ld_out,
                         int num_rows, void* stream);

/* BackpropagateFnc (LPS.h:334-499) for the chunk of the last propagate.  in_diff may be NULL
 * (first trainable layer).  Leaves the fresh gradient (sum over all T*S rows, LPS.h:468-487
 * with beta = 0) in arena 2; momentum is applied in lstmp_b200_update so that a data-parallel
 * caller can all-reduce the fresh gradient first. */
int lstmp_b200_backpropagate(lstmp_b200_handle_t h, const float* in, size_t ld_in, const float* out_diff,
                             size_t ld_od, float* in_diff, size_t ld_id, int num_rows, void* stream);

/* corr = G + momentum*corr (LPS.h:465-487), then param -= learn_rate*corr (LPS.h:501-512). */
int lstmp_b200_update(lstmp_b200_handle_t h, float learn_rate, float momentum, void* stream);

/* Sum the fresh gradient arena over all ranks of `nccl_comm` (an ncclComm_t) in place, on
 * `stream`.  NCCL is resolved at run time with dlopen("libnccl.so.2"); returns
 * LSTMP_B200_EUNSUPPORTED when it cannot be loaded. */
int lstmp_b200_allreduce_grads_nccl(lstmp_b200_handle_t h, void* nccl_comm, void* stream);
/* Overlapped exchange: with a communicator set, lstmp_b200_backpropagate itself sums each gradient block over the ranks
 * on `comm_stream` as soon as the kernel that produced it has been enqueued -- w_gifo_x after its GEMM, then
 * w_gifo_r | bias | peepholes, then w_r_m -- so that the all-reduce of a block runs under the following gradient GEMMs
 * (and under the backward pass of the layers below); lstmp_b200_update[_clipped] waits for the last block.  Every rank
 * must call backpropagate on its engines in the same order.  nccl_comm == NULL switches it off. */
int lstmp_b200_set_nccl(lstmp_b200_handle_t h, void* nccl_comm, void* comm_stream);

/* Introspection for benchmarks and tests. */
typedef struct {
  int input_dim, cell_dim, recur_dim, num_stream, max_frames;
  int sm_count;
  int ngroups, ctas_per_group, streams_per_group, cells_per_cta, rcols_per_cta;
  size_t fwd_smem_bytes, bwd_smem_bytes;
  size_t workspace_bytes;          /* activations + scratch owned by the engine */
  unsigned long long kernel_launches; /* kernels launched by this handle so far */
  int gemm_backend;                /* 0 = fp32 SIMT, 1 = tcgen05 3xTF32 */
  int weights_streamed;            /* 1: weight slices do not fit in shared memory; per-timestep GEMM path */
  int fwd_tensor_core;             /* forward time loop: 2 TMA-fed tcgen05 (bf16 hi/lo split, num_stream <= 64),
                                      1 tcgen05 with register loaders (round 1), 0 FP32 FFMA */
  int bwd_tensor_core;             /* backward time loop: 2 TMA-fed tcgen05 in clusters, 0 FP32 FFMA */
  int bwd_ctas, bwd_cluster;       /* grid and cluster size of the backward tcgen05 kernel (0 when not used) */
} lstmp_b200_info_t;
int lstmp_b200_get_info(lstmp_b200_handle_t h, lstmp_b200_info_t* info);

/* Per-kernel device timing (CUDA events recorded on the launch stream around every kernel this
 * handle launches while enabled).  kinds: 0 input GEMM (+bias), 1 forward recurrent kernel,
 * 2 backward recurrent kernel, 3 in_diff GEMM, 4 weight-gradient GEMMs, 5 bias/peephole gradient
 * gather, 6 update, 7 reset.  lstmp_b200_timing_read synchronises, returns the accumulated
 * milliseconds / launch counts since the last read and clears them. */
#define LSTMP_B200_TIMING_KINDS 8
typedef struct {
  double ms[LSTMP_B200_TIMING_KINDS];
  unsigned long long count[LSTMP_B200_TIMING_KINDS];
} lstmp_b200_timing_t;
int lstmp_b200_timing_enable(lstmp_b200_handle_t h, int on);
int lstmp_b200_timing_read(lstmp_b200_handle_t h, lstmp_b200_timing_t* out);

/* Debug/test access to the per-chunk activation record in the reference's propagate_buf_ /
 * backpropagate_buf_ column layout [g|i|f|o|c|h|m (C each)|r (R)] for frames 1..T
 * (LPS.h:234-241,355-362).  dst is a HOST or device buffer of T*S rows x (7C+R).  In the backward
 * record only the g,i,f,o and r blocks are materialised (d_c/d_h/d_m never leave the SM); the
 * others are written as zeros. */
int lstmp_b200_get_record(lstmp_b200_handle_t h, int backward, float* dst, size_t ld_dst, void* stream);

/* Test hook: run one of the engine's GEMM kernels on caller device buffers.
 *   C[M x N] = alpha * op(A) * op(B) + beta * C (+ bias[n]);  tA/tB: 0 = as stored, 1 = transposed
 *   (op(A) is M x K, op(B) is K x N; row-major with leading dimensions, as CuMatrixBase::AddMatMat,
 *   cu-matrix.cc:909-945).  backend: 0 = fp32 SIMT, 1 = tcgen05 3xTF32, 2 = tcgen05 on bf16 hi/lo tile images (the
 *   engine's default; returns LSTMP_B200_EUNSUPPORTED if the shape/alignment is not handled by that kernel). */
int lstmp_b200_debug_gemm(int backend, float* C, size_t ldc, int M, int N, int K, float alpha, const float* A,
                          size_t lda, int tA, const float* B, size_t ldb, int tB, float beta, const float* bias,
                          void* stream);

/* Test hook for the grouped launch the engine uses after a layer's backward time loop (in_diff, G(w_gifo_x),
 * G(w_gifo_r), G(w_r_m): LPS.h:457, 468, 471, 486 -- four AddMatMat calls in the reference): up to 4 independent
 * products, same argument meaning as lstmp_b200_debug_gemm, run as one operand-split launch + one persistent tcgen05
 * launch + one split-K reduce.  Returns the number of kernel launches (>= 2) or a negative LSTMP_B200_E* code. */
typedef struct lstmp_b200_gemm_desc {
  float* C;
  size_t ldc;
  int M, N, K;
  float alpha;
  const float* A;
  size_t lda;
  int tA;
  const float* B;
  size_t ldb;
  int tB;
  float beta;
  const float* bias;
} lstmp_b200_gemm_desc;
int lstmp_b200_debug_gemm_group(int n, const lstmp_b200_gemm_desc* d, void* stream);

/* Update with the element-wise gradient clipping of the single-stream `standard/` component
 * (standard/nnet/nnet-lstm-projected.h:469-493): corr = G + momentum*corr (the AddMatMat beta of :438-460); every
 * element of corr clamped to [-max_grad, +max_grad] IN PLACE (ClipGradMat / ClipGradVec, :469-478, max_grad = 50 at
 * :482); param -= learn_rate * corr.  max_grad <= 0 means no clipping (== lstmp_b200_update). */
int lstmp_b200_update_clipped(lstmp_b200_handle_t h, float learn_rate, float momentum, float max_grad, void* stream);

/* ---- Multi-stream chunk assembly on the device -------------------------------------
 * Replaces the host-side batch fill of the trainer (google/nnetbin/bd-nnet-train-lstm-streams.cc:187-206, feature
 * part), the per-chunk H2D copy of the [T*S x D] chunk (:212) and the AddShift + Rescale feature transform
 * (google/feature_transform.nnet.txt:2-5).  An utterance goes host -> device ONCE, when a stream takes it
 * (load_utt, TRAIN.cc:152-170: pinned staging + async copy on the dispatcher's own stream, two device slots per
 * stream); per chunk `assemble` copies 3*S ints and runs one gather kernel:
 *   feat[t*S + s] = transform(utt_s[min(curt[s] + t + targets_delay, lent[s] - 1)])   (zeros when lent[s] == 0)
 * with transform(x) = (x + shift) * scale, bit-identical to the host loop.  curt / lent / targets / frame_mask /
 * new_utt_flags stay with the caller (kaldi/b200-stream-dispatch.h mirrors the reference loop).  feat_dim % 4 == 0. */
typedef struct lstmp_b200_dispatch* lstmp_b200_dispatch_handle_t;
int lstmp_b200_dispatch_create(int num_stream, int batch_size, int targets_delay, int feat_dim, int max_utt_frames,
                               int device, lstmp_b200_dispatch_handle_t* out);
int lstmp_b200_dispatch_destroy(lstmp_b200_dispatch_handle_t h);
/* shift / scale: HOST vectors of feat_dim floats, either may be NULL (component absent). */
int lstmp_b200_dispatch_set_transform(lstmp_b200_dispatch_handle_t h, const float* shift, const float* scale);
/* feats: HOST matrix [num_frames x feat_dim], row stride ld; the copy is asynchronous, feats may be freed on return. */
int lstmp_b200_dispatch_load_utt(lstmp_b200_dispatch_handle_t h, int stream, const float* feats, size_t ld,
                                 int num_frames);
/* curt, lent: HOST int32[num_stream] as BEFORE this chunk's fill (the caller advances curt by batch_size afterwards,
 * TRAIN.cc:204); feat: DEVICE [batch_size*num_stream x feat_dim], row stride ld_feat (a multiple of 4). */
int lstmp_b200_dispatch_assemble(lstmp_b200_dispatch_handle_t h, const int32_t* curt, const int32_t* lent, float* feat,
                                 size_t ld_feat, void* stream);
typedef struct {
  unsigned long long kernel_launches, h2d_bytes, utterances_loaded, chunks_assembled;
} lstmp_b200_dispatch_stats_t;
int lstmp_b200_dispatch_get_stats(lstmp_b200_dispatch_handle_t h, lstmp_b200_dispatch_stats_t* out);

/* TimeShift::PropagateFnc (standard/nnet/nnet-time-shift.h:42-51): out row dst = in row clamp(dst + shift, 0,
 * num_rows - 1); device matrices, out must not alias in.  (Its BackpropagateFnc is empty in the reference, :53-56.) */
int lstmp_b200_time_shift(const float* in, size_t ld_in, float* out, size_t ld_out, int num_rows, int num_cols, int shift,
                          void* stream);

/* ---- Masked cross-entropy with sparse targets --------------------------------
 * Replaces Xent::EvalMasked (google/nnet/nnet-loss.cc:76-164) and the accumulators / Report() of class Xent
 * (nnet-loss.h:33-75, nnet-loss.cc:293-307).  net_out = the softmax outputs [num_frames x num_pdf] (device, row
 * stride ld_out), diff [num_frames x num_pdf] (device, fully written) = frame_mask * (net_out - target).  The Kaldi
 * `Posterior` (host std::vector<std::vector<std::pair<int32,BaseFloat>>>) is passed flattened as CSR:
 * post_row_ptr[num_frames + 1], post_pdf[nnz], post_weight[nnz] (host; duplicates of a pdf within a frame accumulate,
 * nnet-loss.cc:93); frame_mask is the host Vector of 0/1 floats (nnet-loss.cc:76, :98-100).  A pdf-id outside
 * [0, num_pdf) is the reference's KALDI_ERR (:88-91) -> LSTMP_B200_EINVAL.  Statistics accumulate on the device
 * (loss_, entropy_ in double; correct_, frames_) and are read back only by lstmp_b200_xent_get_stats.
 * The call is asynchronous on `stream`; the host arrays may be reused as soon as it returns. */
typedef struct lstmp_b200_xent* lstmp_b200_xent_handle_t;
typedef struct {
  double loss;      /* loss_    : -sum mask * t * log(y)                 (nnet-loss.cc:123-128,138) */
  double entropy;   /* entropy_ : -sum mask * t * log(t + 1e-20)         (:130-136,139) */
  long long correct;/* correct_ : valid frames whose arg-max matches the target's      (:108-121,140) */
  long long frames; /* frames_  : sum over calls of (int32) sum(frame_mask)            (:141-142) */
  unsigned long long kernel_launches;
} lstmp_b200_xent_stats_t;
int lstmp_b200_xent_create(int max_frames, int device, lstmp_b200_xent_handle_t* out);
int lstmp_b200_xent_destroy(lstmp_b200_xent_handle_t h);
int lstmp_b200_xent_eval_masked(lstmp_b200_xent_handle_t h, const float* frame_mask_host, const float* net_out,
                                size_t ld_out, int num_frames, int num_pdf, const int32_t* post_row_ptr_host,
                                const int32_t* post_pdf_host, const float* post_weight_host, float* diff,
                                size_t ld_diff, void* stream);
int lstmp_b200_xent_get_stats(lstmp_b200_xent_handle_t h, lstmp_b200_xent_stats_t* out, void* stream);
int lstmp_b200_xent_reset_stats(lstmp_b200_xent_handle_t h, void* stream);

/* Softmax + EvalMasked fused: `logits` are the PRE-softmax activations (the output of
 * the last AffineTransform); one kernel reads each row once, applies Kaldi's ApplySoftMaxPerRow (subtract the row
 * maximum, exp, scale by 1/sum) in shared memory and writes diff = frame_mask * (softmax - target); post_out (nullable)
 * receives the soft-max outputs themselves.  diff may alias logits (in place).  Same statistics, posterior format and
 * error behaviour as lstmp_b200_xent_eval_masked.  num_pdf * 4 bytes must fit in shared memory (<= 200 KB). */
int lstmp_b200_xent_eval_masked_logits(lstmp_b200_xent_handle_t h, const float* frame_mask_host, const float* logits,
                                       size_t ld_logits, int num_frames, int num_pdf, const int32_t* post_row_ptr_host,
                                       const int32_t* post_pdf_host, const float* post_weight_host, float* post_out,
                                       size_t ld_post, float* diff, size_t ld_diff, void* stream);

/* ---- Output tail: AffineTransform input_dim -> num_pdf + Softmax + Xent::EvalMasked --
 * One component for the last two layers of the reference network (google/nnet.proto:4-5: <AffineTransform> 16624 512,
 * <Softmax> 16624 16624) and the objective the trainer evaluates on them
 * (google/nnetbin/bd-nnet-train-lstm-streams.cc:215-228): with it BASELINE.json configs[3] runs end to end.
 * Parameters / momentum / fresh gradients are flat arenas [linearity_ (num_pdf x input_dim, row-major) | bias_ (num_pdf)]
 * (`which`: 0 params, 1 momentum-accumulated corr, 2 fresh gradients -- the one a data-parallel caller all-reduces).
 *   propagate_eval: logits = in * W^T + b ; y = softmax(logits) ; diff = mask * (y - t) kept inside the component;
 *                   statistics accumulate as in lstmp_b200_xent_*; post_out (nullable, device) receives y.
 *   backpropagate:  in_diff = diff * W (nullable) ; G(W) = diff^T * in ; G(b) = column sums of diff.
 *   update:         corr = G + momentum * corr ; param -= learn_rate * corr   ([upstream] AffineTransform::Update with
 *                   learn-rate coefficients 1 and no L1/L2 penalty, as in the reference's nnet.proto). */
typedef struct lstmp_b200_tail* lstmp_b200_tail_handle_t;
int lstmp_b200_tail_create(int input_dim, int num_pdf, int max_frames, int device, lstmp_b200_tail_handle_t* out);
int lstmp_b200_tail_destroy(lstmp_b200_tail_handle_t h);
int lstmp_b200_tail_arena(lstmp_b200_tail_handle_t h, int which, float** dev_ptr, size_t* count);
int lstmp_b200_tail_set_flat(lstmp_b200_tail_handle_t h, int which, const float* src, void* stream);
int lstmp_b200_tail_get_flat(lstmp_b200_tail_handle_t h, int which, float* dst, void* stream);
int lstmp_b200_tail_propagate_eval(lstmp_b200_tail_handle_t h, const float* in, size_t ld_in, int num_frames,
                                   const float* frame_mask_host, const int32_t* post_row_ptr_host,
                                   const int32_t* post_pdf_host, const float* post_weight_host, float* post_out,
                                   size_t ld_post, void* stream);
int lstmp_b200_tail_backpropagate(lstmp_b200_tail_handle_t h, const float* in, size_t ld_in, float* in_diff, size_t ld_id,
                                  int num_frames, void* stream);
int lstmp_b200_tail_update(lstmp_b200_tail_handle_t h, float learn_rate, float momentum, void* stream);
int lstmp_b200_tail_allreduce_grads_nccl(lstmp_b200_tail_handle_t h, void* nccl_comm, void* stream);
/* test / debug: the diff of the last propagate_eval, [num_frames x num_pdf] to a host or device buffer */
int lstmp_b200_tail_get_diff(lstmp_b200_tail_handle_t h, float* dst, size_t ld_dst, void* stream);
int lstmp_b200_tail_get_stats(lstmp_b200_tail_handle_t h, lstmp_b200_xent_stats_t* out, void* stream);
int lstmp_b200_tail_reset_stats(lstmp_b200_tail_handle_t h, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSTMP_B200_H_ */
