// kaldi/b200-lstm-projected-streams.h
//
// Drop-in replacement for the reference's kaldi::nnet1::LstmProjectedStreams
// (google/nnet/bd-nnet-lstm-projected-streams.h): the same Component surface -- PropagateFnc /
// BackpropagateFnc / Update / Reset / InitData / ReadData / WriteData / NumParams / GetParams /
// Copy / GetType -- with the arithmetic moved behind the C ABI of include/lstmp_b200.h
// (hand-written sm_100a kernels).  Host side stays C++ against Kaldi's CuMatrix types; no CPU
// fallback, no multi-backend dispatch.
//
// Build against a real Kaldi tree with -DHAVE_KALDI, or against
// compat/kaldi-compat.h (test scaffolding) as the tests here do.
#ifndef B200_KALDI_NNET_LSTM_PROJECTED_STREAMS_H_
#define B200_KALDI_NNET_LSTM_PROJECTED_STREAMS_H_

#ifdef HAVE_KALDI
#include "cudamatrix/cu-math.h"
#include "nnet/nnet-component.h"
#include "nnet/nnet-various.h"
#else
#include "compat/kaldi-compat.h"
#endif
#include "lstmp_b200.h"

namespace kaldi {
namespace nnet1 {

// Raw device pointer of a CuMatrixBase.  Data() is protected in the reference's snapshot of
// cu-matrix.h (:446-461) and public in newer Kaldi; define B200_CUMATRIX_DATA_VIA_ROW to go through
// the public Row(0).Data() accessor instead ("raw pointer access").
template <class M>
inline const BaseFloat* B200DevPtr(const M& m) {
#ifdef B200_CUMATRIX_DATA_VIA_ROW
  return m.NumRows() ? m.Row(0).Data() : NULL;
#else
  return m.Data();
#endif
}
template <class M>
inline BaseFloat* B200DevPtr(M* m) {
  return const_cast<BaseFloat*>(B200DevPtr(static_cast<const M&>(*m)));
}

class B200LstmProjectedStreams : public UpdatableComponent {
 public:
  B200LstmProjectedStreams(int32 input_dim, int32 output_dim)  // LPS.h:27-33
      : UpdatableComponent(input_dim, output_dim), ncell_(0), nrecur_(output_dim), nstream_(0), max_frames_(20),
        engine_(NULL) {}
  ~B200LstmProjectedStreams() { lstmp_b200_destroy(engine_); }

  Component* Copy() const {  // deep copy incl. carried state and momentum buffers (LPS.h:38)
    B200LstmProjectedStreams* c = new B200LstmProjectedStreams(input_dim_, output_dim_);
    c->ncell_ = ncell_;
    c->nstream_ = nstream_;
    c->max_frames_ = max_frames_;
    c->opts_ = opts_;
    if (engine_) Check(lstmp_b200_clone(engine_, &c->engine_));
    return c;
  }
  ComponentType GetType() const { return kLstmProjectedStreams; }  // marker <LstmProjectedStreams>, google/nnet.proto:3

  void InitData(std::istream& is) {  // LPS.h:55-99
    float param_scale = 0.02;
    std::string token;
    while (!is.eof()) {
      ReadToken(is, false, &token);
      if (token == "<CellDim>") ReadBasicType(is, false, &ncell_);
      else if (token == "<NumStream>") ReadBasicType(is, false, &nstream_);
      else if (token == "<ParamScale>") ReadBasicType(is, false, &param_scale);
      else KALDI_ERR << "Unknown token " << token << ", a typo in config?"
                     << " (CellDim|NumStream|ParamScale)";
      is >> std::ws;
    }
    CreateEngine();
    // uniform in [-scale, +scale] (LPS.h:41-53); own generator, Kaldi's RandUniform stream is not reproducible
    Vector<BaseFloat> flat(NumParams());
    uint32_t s = 4321u;
    for (int32 i = 0; i < flat.Dim(); i++) {
      s = s * 1664525u + 1013904223u;
      flat(i) = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.0f * param_scale;
    }
    SetParams(flat);
  }

  void ReadData(std::istream& is, bool binary) {  // LPS.h:101-131, same on-disk order
    ExpectToken(is, binary, "<CellDim>");
    ReadBasicType(is, binary, &ncell_);
    ExpectToken(is, binary, "<NumStream>");
    ReadBasicType(is, binary, &nstream_);
    Matrix<BaseFloat> w_gifo_x, w_gifo_r, w_r_m;
    Vector<BaseFloat> bias, p_i, p_f, p_o;
    w_gifo_x.Read(is, binary);
    w_gifo_r.Read(is, binary);
    bias.Read(is, binary);
    p_i.Read(is, binary);
    p_f.Read(is, binary);
    p_o.Read(is, binary);
    w_r_m.Read(is, binary);
    KALDI_ASSERT(w_gifo_x.NumRows() == 4 * ncell_ && w_gifo_x.NumCols() == input_dim_);
    KALDI_ASSERT(w_gifo_r.NumRows() == 4 * ncell_ && w_gifo_r.NumCols() == nrecur_);
    KALDI_ASSERT(w_r_m.NumRows() == nrecur_ && w_r_m.NumCols() == ncell_);
    CreateEngine();  // zeroes state and momentum buffers, as LPS.h:119-130
    Check(lstmp_b200_set_params(engine_, w_gifo_x.Data(), w_gifo_x.Stride(), w_gifo_r.Data(), w_gifo_r.Stride(),
                                bias.Data(), p_i.Data(), p_f.Data(), p_o.Data(), w_r_m.Data(), w_r_m.Stride(), NULL));
  }

  void WriteData(std::ostream& os, bool binary) const {  // LPS.h:133-150
    WriteToken(os, binary, "<CellDim>");
    WriteBasicType(os, binary, ncell_);
    WriteToken(os, binary, "<NumStream>");
    WriteBasicType(os, binary, nstream_);
    Matrix<BaseFloat> w_gifo_x(4 * ncell_, input_dim_), w_gifo_r(4 * ncell_, nrecur_), w_r_m(nrecur_, ncell_);
    Vector<BaseFloat> bias(4 * ncell_), p_i(ncell_), p_f(ncell_), p_o(ncell_);
    Check(lstmp_b200_get_params(engine_, w_gifo_x.Data(), w_gifo_x.Stride(), w_gifo_r.Data(), w_gifo_r.Stride(),
                                bias.Data(), p_i.Data(), p_f.Data(), p_o.Data(), w_r_m.Data(), w_r_m.Stride(), NULL));
    w_gifo_x.Write(os, binary);
    w_gifo_r.Write(os, binary);
    bias.Write(os, binary);
    p_i.Write(os, binary);
    p_f.Write(os, binary);
    p_o.Write(os, binary);
    w_r_m.Write(os, binary);
  }

  int32 NumParams() const {  // LPS.h:152-160
    return 4 * ncell_ * input_dim_ + 4 * ncell_ * nrecur_ + 4 * ncell_ + 3 * ncell_ + nrecur_ * ncell_;
  }
  void GetParams(Vector<BaseFloat>* wei_copy) const {  // LPS.h:162-189 (same flat order)
    wei_copy->Resize(NumParams());
    Check(lstmp_b200_get_flat(engine_, 0, wei_copy->Data(), NULL));
  }
  void SetParams(const Vector<BaseFloat>& flat) {
    KALDI_ASSERT(flat.Dim() == NumParams());
    Check(lstmp_b200_set_flat(engine_, 0, flat.Data(), NULL));
  }
  void GetGradient(Vector<BaseFloat>* corr) const {  // the *_corr_ buffers (InfoGradient, LPS.h:201-210)
    corr->Resize(NumParams());
    Check(lstmp_b200_get_flat(engine_, 1, corr->Data(), NULL));
  }

  void Reset(std::vector<int>& stream_reset_flag) {  // LPS.h:212-220
    KALDI_ASSERT(nstream_ == (int32)stream_reset_flag.size());
    std::vector<int32_t> f(stream_reset_flag.begin(), stream_reset_flag.end());
    Check(lstmp_b200_reset(engine_, f.data(), (int)f.size(), NULL));
  }

  void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) {  // LPS.h:222-332
    KALDI_ASSERT(in.NumRows() % nstream_ == 0);  // :225
    EnsureFrames(in.NumRows() / nstream_);
    Check(lstmp_b200_propagate(engine_, B200DevPtr(in), in.Stride(), B200DevPtr(out), out->Stride(), in.NumRows(),
                               NULL));
  }
  void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out,
                        const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) {  // LPS.h:334-499
    (void)out;  // the reference reads its own propagate_buf_ too (:342-349)
    Check(lstmp_b200_backpropagate(engine_, B200DevPtr(in), in.Stride(), B200DevPtr(out_diff), out_diff.Stride(),
                                   in_diff ? B200DevPtr(in_diff) : NULL, in_diff ? in_diff->Stride() : 0,
                                   in.NumRows(), NULL));
  }
  // PAIRING CONTRACT: the reference accumulates momentum in BackpropagateFnc (corr = G + mmt*corr, LPS.h:465-487) and
  // Update only applies param -= lr*corr.  Here BackpropagateFnc leaves the FRESH gradient G (so that it can be
  // all-reduced) and Update does both steps: identical under the trainer's Propagate -> Backpropagate -> Update order;
  // GetGradient / InfoGradient between Backpropagate and Update still show the previous chunk's corr, and two
  // Backpropagate calls without an Update keep only the second gradient.  (INTEGRATION.md.)
  void Update(const CuMatrixBase<BaseFloat>&, const CuMatrixBase<BaseFloat>&) {  // LPS.h:465-487 + :501-512
    if (comm_) Check(lstmp_b200_allreduce_grads_nccl(engine_, comm_, NULL));  // data-parallel: sum fresh gradients
    if (max_grad_ > 0)  // the standard single-stream version's element-wise clip (nnet-lstm-projected.h:469-493)
      Check(lstmp_b200_update_clipped(engine_, opts_.learn_rate, opts_.momentum, max_grad_, NULL));
    else
      Check(lstmp_b200_update(engine_, opts_.learn_rate, opts_.momentum, NULL));
  }
  void SetMaxGrad(BaseFloat max_grad) { max_grad_ = max_grad; }

  // Data-parallel training: an ncclComm_t shared by the ranks that shard the streams.
  void SetNcclComm(void* comm) { comm_ = comm; }
  lstmp_b200_handle_t Engine() const { return engine_; }

 private:
  static void Check(int rc) {
    if (rc != 0) KALDI_ERR << "lstmp_b200: " << lstmp_b200_last_error() << " (code " << rc << ")";
  }
  void CreateEngine() {
    lstmp_b200_destroy(engine_);
    engine_ = NULL;
    int dev = 0;
    cudaGetDevice(&dev);  // Kaldi's CuDevice has already selected the GPU
    Check(lstmp_b200_create(input_dim_, ncell_, nrecur_, nstream_, max_frames_, dev, &engine_));
  }
  void EnsureFrames(int32 T) {  // the reference resizes its buffers per call (LPS.h:230)
    if (T <= max_frames_) return;
    Vector<BaseFloat> p(NumParams()), g(NumParams());
    Matrix<BaseFloat> c(nstream_, ncell_), r(nstream_, nrecur_);
    Check(lstmp_b200_get_flat(engine_, 0, p.Data(), NULL));
    Check(lstmp_b200_get_flat(engine_, 1, g.Data(), NULL));
    Check(lstmp_b200_get_state(engine_, c.Data(), c.Stride(), r.Data(), r.Stride(), NULL));
    max_frames_ = T;
    CreateEngine();
    Check(lstmp_b200_set_flat(engine_, 0, p.Data(), NULL));
    Check(lstmp_b200_set_flat(engine_, 1, g.Data(), NULL));
    Check(lstmp_b200_set_state(engine_, c.Data(), c.Stride(), r.Data(), r.Stride(), NULL));
  }

  int32 ncell_, nrecur_, nstream_, max_frames_;
  lstmp_b200_handle_t engine_;
  void* comm_ = NULL;           // per process; NOT carried over by Copy() -- call SetNcclComm on the copy
  BaseFloat max_grad_ = 0;
};

}  // namespace nnet1
}  // namespace kaldi
#endif
