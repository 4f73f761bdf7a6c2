// oracle/ref_build/ref_driver_standard.cc -- TEST INFRASTRUCTURE (translation unit 2 of oracle/_ref/libkaldi_lstm_ref.so).
//
// The reference's standard/ version, compiled where it lies under /root/reference:
//   standard/nnet/nnet-lstm-projected.h   (#include, unmodified) -> LstmProjected (single utterance, S = 1; its
//                                          Update clips the gradients element-wise at 50, :480-493)
//   standard/nnet/nnet-time-shift.h       (#include, unmodified) -> TimeShift (:42-51)
// Own translation unit because standard/ and google/ both own a directory called nnet/.
#include "kaldi-ref-shim.h"

#define private public
#define protected public
#include "nnet/nnet-lstm-projected.h"
#include "nnet/nnet-time-shift.h"
#undef private
#undef protected

using namespace kaldi;
using namespace kaldi::nnet1;

namespace {
struct RefStd {
  LstmProjected comp;
  CuMatrix<BaseFloat> in, out, out_diff, in_diff;
  NnetTrainOptions opts;
  RefStd(int I, int R) : comp(I, R) {}
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
thread_local std::string g_err;
}  // namespace

#define REF_TRY try {
#define REF_CATCH(rc)               \
  }                                 \
  catch (const std::exception &e) { \
    g_err = e.what();               \
    return rc;                      \
  }

extern "C" {
const char *lstmp_std_ref_last_error() { return g_err.c_str(); }
void *lstmp_std_ref_create(int I, int C, int R) {
  REF_TRY
  RefStd *h = new RefStd(I, R);
  std::ostringstream cfg;
  cfg << "<CellDim> " << C << " <ParamScale> 0.01 ";
  std::istringstream is(cfg.str());
  h->comp.InitData(is);
  return h;
  REF_CATCH(NULL)
}
void lstmp_std_ref_destroy(void *hv) { delete (RefStd *)hv; }
long lstmp_std_ref_num_params(void *hv) { return ((RefStd *)hv)->comp.NumParams(); }
int lstmp_std_ref_get_params(void *hv, float *dst) {
  REF_TRY
  Vector<BaseFloat> v;
  ((RefStd *)hv)->comp.GetParams(&v);
  std::memcpy(dst, v.Data(), sizeof(float) * v.Dim());
  return 0;
  REF_CATCH(-1)
}
void lstmp_std_ref_set_params(void *hv, const float *p) {
  LstmProjected &c = ((RefStd *)hv)->comp;
  flat_to_mat(c.w_gifo_x_, p);
  flat_to_mat(c.w_gifo_r_, p);
  flat_to_vec(c.bias_, p);
  flat_to_vec(c.peephole_i_c_, p);
  flat_to_vec(c.peephole_f_c_, p);
  flat_to_vec(c.peephole_o_c_, p);
  flat_to_mat(c.w_r_m_, p);
}
void lstmp_std_ref_get_corr(void *hv, float *p) {
  LstmProjected &c = ((RefStd *)hv)->comp;
  mat_to_flat(c.w_gifo_x_corr_, p);
  mat_to_flat(c.w_gifo_r_corr_, p);
  vec_to_flat(c.bias_corr_, p);
  vec_to_flat(c.peephole_i_c_corr_, p);
  vec_to_flat(c.peephole_f_c_corr_, p);
  vec_to_flat(c.peephole_o_c_corr_, p);
  mat_to_flat(c.w_r_m_corr_, p);
}
void lstmp_std_ref_set_corr(void *hv, const float *p) {
  LstmProjected &c = ((RefStd *)hv)->comp;
  flat_to_mat(c.w_gifo_x_corr_, p);
  flat_to_mat(c.w_gifo_r_corr_, p);
  flat_to_vec(c.bias_corr_, p);
  flat_to_vec(c.peephole_i_c_corr_, p);
  flat_to_vec(c.peephole_f_c_corr_, p);
  flat_to_vec(c.peephole_o_c_corr_, p);
  flat_to_mat(c.w_r_m_corr_, p);
}
int lstmp_std_ref_propagate(void *hv, const float *in, int ld_in, float *out, int ld_out, int rows) {
  REF_TRY
  RefStd *h = (RefStd *)hv;
  copy_in(&h->in, in, rows, h->comp.InputDim(), ld_in);
  h->comp.Propagate(h->in, &h->out);
  copy_out(h->out, out, ld_out);
  return 0;
  REF_CATCH(-1)
}
int lstmp_std_ref_backpropagate(void *hv, const float *in, int ld_in, const float *out_diff, int ld_od, float *in_diff,
                                int ld_id, int rows, float momentum) {
  REF_TRY
  RefStd *h = (RefStd *)hv;
  copy_in(&h->in, in, rows, h->comp.InputDim(), ld_in);
  copy_in(&h->out_diff, out_diff, rows, h->comp.OutputDim(), ld_od);
  h->opts.momentum = momentum;
  h->comp.SetTrainOptions(h->opts);
  h->comp.Backpropagate(h->in, h->out, h->out_diff, &h->in_diff);
  if (in_diff) copy_out(h->in_diff, in_diff, ld_id);
  return 0;
  REF_CATCH(-1)
}
// Update incl. the element-wise gradient clip (standard/nnet/nnet-lstm-projected.h:480-511)
int lstmp_std_ref_update(void *hv, float lr) {
  REF_TRY
  RefStd *h = (RefStd *)hv;
  h->opts.learn_rate = lr;
  h->comp.SetTrainOptions(h->opts);
  h->comp.Update(h->in, h->out_diff);
  return 0;
  REF_CATCH(-1)
}

// TimeShift::PropagateFnc (standard/nnet/nnet-time-shift.h:42-51)
int timeshift_ref_propagate(int shift, const float *in, int ld_in, float *out, int ld_out, int rows, int dim) {
  REF_TRY
  TimeShift ts(dim, dim);
  std::ostringstream cfg;
  cfg << "<Shift> " << shift << " ";
  std::istringstream is(cfg.str());
  ts.InitData(is);
  CuMatrix<BaseFloat> min, mout;
  copy_in(&min, in, rows, dim, ld_in);
  ts.Propagate(min, &mout);
  copy_out(mout, out, ld_out);
  return 0;
  REF_CATCH(-1)
}
}  // extern "C"
