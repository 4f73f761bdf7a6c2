// oracle/ref_build/ref_driver_google.cc -- TEST INFRASTRUCTURE (translation unit 1 of oracle/_ref/libkaldi_lstm_ref.so).
//
// Compiles the reference's OWN google/ sources where they lie under /root/reference:
//   google/nnet/bd-nnet-lstm-projected-streams.h   (#include, unmodified)          -> LstmProjectedStreams
//   google/nnet/nnet-loss.h                        (#include, unmodified)          -> class Xent
//   google/nnet/nnet-loss.cc: Xent::EvalMasked, Xent::Report                       (bodies extracted at build time)
//   google/matrix/kaldi-matrix.cc, google/cudamatrix/cu-matrix.cc                  (bodies extracted at build time)
// against the CPU-computing Kaldi surface in shim/kaldi-ref-shim.h, and exports a small C ABI so that the tests
// can run the REFERENCE ITSELF next to the restated oracle (oracle/lstmp_streams_oracle.c, oracle/xent_oracle.py).
// `private` is re-defined around the reference headers ONLY so that the tests can set the parameters and read
// the `*_corr_` / `propagate_buf_` members; no reference code is changed by it.
#include "kaldi-ref-shim.h"

#define private public
#define protected public
#include "nnet/bd-nnet-lstm-projected-streams.h"
#include "nnet/nnet-loss.h"
#undef private
#undef protected

namespace kaldi {
int g_ref_verbose = -1;
ref_sgemm_fn g_ref_sgemm = NULL;

// plain triple loop in the cblas contract; used when no external sgemm was installed
void ref_builtin_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float *A, int lda, const float *B,
                       int ldb, float beta, float *C, int ldc) {
  for (int i = 0; i < M; i++) {
    float *c = C + (size_t)i * ldc;
    if (beta == 0.0f) {
      for (int j = 0; j < N; j++) c[j] = 0.0f;
    } else if (beta != 1.0f) {
      for (int j = 0; j < N; j++) c[j] *= beta;
    }
    if (!transB) {
      for (int k = 0; k < K; k++) {
        float a = alpha * (transA ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k]);
        const float *b = B + (size_t)k * ldb;
        for (int j = 0; j < N; j++) c[j] += a * b[j];
      }
    } else {
      for (int j = 0; j < N; j++) {
        const float *b = B + (size_t)j * ldb;
        float s = 0.0f;
        if (!transA) {
          const float *a = A + (size_t)i * lda;
          for (int k = 0; k < K; k++) s += a[k] * b[k];
        } else {
          for (int k = 0; k < K; k++) s += A[(size_t)k * lda + i] * b[k];
        }
        c[j] += alpha * s;
      }
    }
  }
}

namespace nnet1 {
#include "loss_ops.inc"
}  // namespace nnet1
}  // namespace kaldi

using namespace kaldi;
using namespace kaldi::nnet1;

namespace {
struct RefLstm {
  LstmProjectedStreams comp;
  CuMatrix<BaseFloat> in, out, out_diff, in_diff;
  NnetTrainOptions opts;
  RefLstm(int I, int R) : comp(I, R) {}
};

void copy_in(CuMatrix<BaseFloat> *m, const float *src, int rows, int cols, int ld) {
  m->Resize(rows, cols, kUndefined);
  for (int r = 0; r < rows; r++) std::memcpy(m->Data() + (size_t)r * m->Stride(), src + (size_t)r * ld, sizeof(float) * cols);
}
void copy_out(const CuMatrixBase<BaseFloat> &m, float *dst, int ld) {
  for (int r = 0; r < m.NumRows(); r++)
    std::memcpy(dst + (size_t)r * ld, m.Data() + (size_t)r * m.Stride(), sizeof(float) * m.NumCols());
}
void mat_to_flat(const CuMatrixBase<BaseFloat> &m, float *&p) {
  copy_out(m, p, m.NumCols());
  p += (size_t)m.NumRows() * m.NumCols();
}
void vec_to_flat(const CuVectorBase<BaseFloat> &v, float *&p) {
  std::memcpy(p, v.Data(), sizeof(float) * v.Dim());
  p += v.Dim();
}
void flat_to_mat(CuMatrixBase<BaseFloat> &m, const float *&p) {
  for (int r = 0; r < m.NumRows(); r++) std::memcpy(m.Data() + (size_t)r * m.Stride(), p + (size_t)r * m.NumCols(), sizeof(float) * m.NumCols());
  p += (size_t)m.NumRows() * m.NumCols();
}
void flat_to_vec(CuVectorBase<BaseFloat> &v, const float *&p) {
  std::memcpy(v.Data(), p, sizeof(float) * v.Dim());
  p += v.Dim();
}
thread_local std::string g_last_error;
}  // namespace

#define REF_TRY try {
#define REF_CATCH(rc)                 \
  }                                   \
  catch (const std::exception &e) {   \
    g_last_error = e.what();          \
    return rc;                        \
  }

extern "C" {
const char *lstmp_ref_last_error() { return g_last_error.c_str(); }
void lstmp_ref_set_sgemm(void *fn) { g_ref_sgemm = (ref_sgemm_fn)fn; }
const char *lstmp_ref_sources() {
  return "google/nnet/bd-nnet-lstm-projected-streams.h google/nnet/nnet-loss.h google/nnet/nnet-loss.cc "
         "google/matrix/kaldi-matrix.cc google/cudamatrix/cu-matrix.cc";
}

// InitData with the reference's own config tokens (LPS.h:55-99), then the caller sets the parameters
void *lstmp_ref_create(int I, int C, int R, int S) {
  REF_TRY
  RefLstm *h = new RefLstm(I, R);
  std::ostringstream cfg;
  cfg << "<CellDim> " << C << " <NumStream> " << S << " <ParamScale> 0.01 ";
  std::istringstream is(cfg.str());
  h->comp.InitData(is);
  return h;
  REF_CATCH(NULL)
}
void lstmp_ref_destroy(void *hv) { delete (RefLstm *)hv; }
long lstmp_ref_num_params(void *hv) { return ((RefLstm *)hv)->comp.NumParams(); }
// GetParams itself (LPS.h:162-189)
int lstmp_ref_get_params(void *hv, float *dst) {
  REF_TRY
  Vector<BaseFloat> v;
  ((RefLstm *)hv)->comp.GetParams(&v);
  std::memcpy(dst, v.Data(), sizeof(float) * v.Dim());
  return 0;
  REF_CATCH(-1)
}
void lstmp_ref_set_params(void *hv, const float *p) {
  LstmProjectedStreams &c = ((RefLstm *)hv)->comp;
  flat_to_mat(c.w_gifo_x_, p);
  flat_to_mat(c.w_gifo_r_, p);
  flat_to_vec(c.bias_, p);
  flat_to_vec(c.peephole_i_c_, p);
  flat_to_vec(c.peephole_f_c_, p);
  flat_to_vec(c.peephole_o_c_, p);
  flat_to_mat(c.w_r_m_, p);
}
// the momentum-accumulated gradients `*_corr_`, in GetParams order
void lstmp_ref_get_corr(void *hv, float *p) {
  LstmProjectedStreams &c = ((RefLstm *)hv)->comp;
  mat_to_flat(c.w_gifo_x_corr_, p);
  mat_to_flat(c.w_gifo_r_corr_, p);
  vec_to_flat(c.bias_corr_, p);
  vec_to_flat(c.peephole_i_c_corr_, p);
  vec_to_flat(c.peephole_f_c_corr_, p);
  vec_to_flat(c.peephole_o_c_corr_, p);
  mat_to_flat(c.w_r_m_corr_, p);
}
void lstmp_ref_set_corr(void *hv, const float *p) {
  LstmProjectedStreams &c = ((RefLstm *)hv)->comp;
  flat_to_mat(c.w_gifo_x_corr_, p);
  flat_to_mat(c.w_gifo_r_corr_, p);
  flat_to_vec(c.bias_corr_, p);
  flat_to_vec(c.peephole_i_c_corr_, p);
  flat_to_vec(c.peephole_f_c_corr_, p);
  flat_to_vec(c.peephole_o_c_corr_, p);
  flat_to_mat(c.w_r_m_corr_, p);
}
void lstmp_ref_get_state(void *hv, float *dst) {
  LstmProjectedStreams &c = ((RefLstm *)hv)->comp;
  copy_out(c.prev_nnet_state_, dst, c.prev_nnet_state_.NumCols());
}
void lstmp_ref_set_state(void *hv, const float *src) {
  LstmProjectedStreams &c = ((RefLstm *)hv)->comp;
  const float *p = src;
  flat_to_mat(c.prev_nnet_state_, p);
}
// rows of propagate_buf_ / backpropagate_buf_ after the last call ((T+2)*S, LPS.h:230, 352)
int lstmp_ref_buf_rows(void *hv, int which) {
  LstmProjectedStreams &c = ((RefLstm *)hv)->comp;
  return which ? c.backpropagate_buf_.NumRows() : c.propagate_buf_.NumRows();
}
void lstmp_ref_get_buf(void *hv, int which, float *dst) {
  LstmProjectedStreams &c = ((RefLstm *)hv)->comp;
  const CuMatrix<BaseFloat> &m = which ? c.backpropagate_buf_ : c.propagate_buf_;
  copy_out(m, dst, m.NumCols());
}
int lstmp_ref_reset(void *hv, const int *flags, int n) {
  REF_TRY
  std::vector<int> f(flags, flags + n);
  ((RefLstm *)hv)->comp.Reset(f);
  return 0;
  REF_CATCH(-1)
}
// Component::Propagate -> PropagateFnc (LPS.h:222-332)
int lstmp_ref_propagate(void *hv, const float *in, int ld_in, float *out, int ld_out, int rows) {
  REF_TRY
  RefLstm *h = (RefLstm *)hv;
  copy_in(&h->in, in, rows, h->comp.InputDim(), ld_in);
  h->comp.Propagate(h->in, &h->out);
  copy_out(h->out, out, ld_out);
  return 0;
  REF_CATCH(-1)
}
// Component::Backpropagate -> BackpropagateFnc (LPS.h:334-499); momentum travels in opts_ as in the trainer
int lstmp_ref_backpropagate(void *hv, const float *in, int ld_in, const float *out_diff, int ld_od, float *in_diff,
                            int ld_id, int rows, float momentum) {
  REF_TRY
  RefLstm *h = (RefLstm *)hv;
  copy_in(&h->in, in, rows, h->comp.InputDim(), ld_in);
  copy_in(&h->out_diff, out_diff, rows, h->comp.OutputDim(), ld_od);
  h->opts.momentum = momentum;
  h->comp.SetTrainOptions(h->opts);
  h->comp.Backpropagate(h->in, h->out, h->out_diff, &h->in_diff);
  if (in_diff) copy_out(h->in_diff, in_diff, ld_id);
  return 0;
  REF_CATCH(-1)
}
// Update (LPS.h:501-512)
int lstmp_ref_update(void *hv, float lr) {
  REF_TRY
  RefLstm *h = (RefLstm *)hv;
  h->opts.learn_rate = lr;
  h->comp.SetTrainOptions(h->opts);
  h->comp.Update(h->in, h->out_diff);
  return 0;
  REF_CATCH(-1)
}
// WriteData / ReadData round trip through the reference's own serialiser (LPS.h:101-160)
long lstmp_ref_write(void *hv, int binary, char *dst, long cap) {
  REF_TRY
  std::ostringstream os;
  ((RefLstm *)hv)->comp.WriteData(os, binary != 0);
  std::string s = os.str();
  if ((long)s.size() <= cap) std::memcpy(dst, s.data(), s.size());
  return (long)s.size();
  REF_CATCH(-1)
}
int lstmp_ref_read(void *hv, int binary, const char *src, long n) {
  REF_TRY
  std::istringstream is(std::string(src, (size_t)n));
  ((RefLstm *)hv)->comp.ReadData(is, binary != 0);
  return 0;
  REF_CATCH(-1)
}

// ---- Xent::EvalMasked / Report (google/nnet/nnet-loss.cc:76-164, 293-307) ---------------------------------
struct RefXent {
  Xent x;
  CuMatrix<BaseFloat> net_out, diff;
};
void *xent_ref_create() { return new RefXent(); }
void xent_ref_destroy(void *hv) { delete (RefXent *)hv; }
// posterior as CSR: row_ptr[frames+1], pdf[nnz], weight[nnz]
int xent_ref_eval_masked(void *hv, const float *mask, const float *net_out, int ld, int frames, int num_pdf,
                         const int *row_ptr, const int *pdf, const float *weight, float *diff, int ld_diff) {
  REF_TRY
  RefXent *h = (RefXent *)hv;
  Vector<BaseFloat> m(frames);
  for (int i = 0; i < frames; i++) m(i) = mask[i];
  copy_in(&h->net_out, net_out, frames, num_pdf, ld);
  Posterior post(frames);
  for (int t = 0; t < frames; t++)
    for (int k = row_ptr[t]; k < row_ptr[t + 1]; k++) post[t].push_back(std::make_pair((int32)pdf[k], (BaseFloat)weight[k]));
  h->x.EvalMasked(m, h->net_out, post, &h->diff);
  copy_out(h->diff, diff, ld_diff);
  return 0;
  REF_CATCH(-1)
}
// frames_, correct_, loss_, entropy_ (nnet-loss.h:60-63)
void xent_ref_get_stats(void *hv, double *out4) {
  RefXent *h = (RefXent *)hv;
  out4[0] = h->x.frames_;
  out4[1] = h->x.correct_;
  out4[2] = h->x.loss_;
  out4[3] = h->x.entropy_;
}
long xent_ref_report(void *hv, char *dst, long cap) {
  REF_TRY
  std::string s = ((RefXent *)hv)->x.Report();
  if ((long)s.size() < cap) std::memcpy(dst, s.c_str(), s.size() + 1);
  return (long)s.size();
  REF_CATCH(-1)
}
}  // extern "C"
