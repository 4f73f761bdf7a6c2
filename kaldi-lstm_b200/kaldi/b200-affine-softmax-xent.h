// kaldi/b200-affine-softmax-xent.h
//
// The output tail of the reference network as ONE Kaldi-side object:
// `<AffineTransform> num_pdf input_dim` + `<Softmax>` (google/nnet.proto:4-5) + the objective the trainer evaluates on
// them, `Xent::EvalMasked` (google/nnet/nnet-loss.cc:76-164; called at google/nnetbin/bd-nnet-train-lstm-streams.cc:219),
// over the C ABI's lstmp_b200_tail_* entry points (include/lstmp_b200.h).  In the trainer it stands where
// `nnet.Propagate` reaches the last two components, `xent.EvalMasked` and the first two steps of `nnet.Backpropagate`:
//
//   tail.PropagateEval(lstm_out, frame_mask, target);      // Affine + Softmax + EvalMasked (diff kept inside)
//   tail.Backpropagate(lstm_out, &lstm_out_diff);          // Affine backward + its weight / bias gradients
//   tail.Update();                                         // momentum + SGD ([upstream] AffineTransform::Update)
//
// Parameters are Kaldi's: linearity_ [num_pdf x input_dim] and bias_ [num_pdf] (AffineTransform::ReadData order).
#ifndef B200_KALDI_AFFINE_SOFTMAX_XENT_H_
#define B200_KALDI_AFFINE_SOFTMAX_XENT_H_

#ifdef HAVE_KALDI
#include "cudamatrix/cu-matrix.h"
#include "hmm/posterior.h"
#include "nnet/nnet-trnopts.h"
#else
#include "compat/kaldi-compat.h"
#endif
#include <sstream>
#include <string>
#include <vector>

#include "lstmp_b200.h"

namespace kaldi {
namespace nnet1 {

class B200AffineSoftmaxXent {
 public:
  B200AffineSoftmaxXent(int32 input_dim, int32 output_dim, int32 max_frames)
      : input_dim_(input_dim), output_dim_(output_dim), max_frames_(max_frames), engine_(NULL) {
    int dev = 0;
    CU_SAFE_CALL(cudaGetDevice(&dev));
    Check(lstmp_b200_tail_create(input_dim_, output_dim_, max_frames_, dev, &engine_));
  }
  ~B200AffineSoftmaxXent() { lstmp_b200_tail_destroy(engine_); }
  B200AffineSoftmaxXent(const B200AffineSoftmaxXent&) = delete;
  B200AffineSoftmaxXent& operator=(const B200AffineSoftmaxXent&) = delete;

  int32 InputDim() const { return input_dim_; }
  int32 OutputDim() const { return output_dim_; }
  int32 NumParams() const { return output_dim_ * input_dim_ + output_dim_; }
  void SetTrainOptions(const NnetTrainOptions& opts) { opts_ = opts; }

  /// linearity_ [output_dim x input_dim], bias_ [output_dim]  (AffineTransform::ReadData order)
  void SetParams(const Matrix<BaseFloat>& linearity, const Vector<BaseFloat>& bias) {
    KALDI_ASSERT(linearity.NumRows() == output_dim_ && linearity.NumCols() == input_dim_ && bias.Dim() == output_dim_);
    std::vector<BaseFloat> flat(NumParams());
    for (int32 r = 0; r < output_dim_; r++)
      for (int32 c = 0; c < input_dim_; c++) flat[(size_t)r * input_dim_ + c] = linearity(r, c);
    for (int32 r = 0; r < output_dim_; r++) flat[(size_t)output_dim_ * input_dim_ + r] = bias(r);
    Check(lstmp_b200_tail_set_flat(engine_, 0, flat.data(), NULL));
  }
  void GetParams(Matrix<BaseFloat>* linearity, Vector<BaseFloat>* bias) const {
    std::vector<BaseFloat> flat(NumParams());
    Check(lstmp_b200_tail_get_flat(engine_, 0, flat.data(), NULL));
    linearity->Resize(output_dim_, input_dim_);
    bias->Resize(output_dim_);
    for (int32 r = 0; r < output_dim_; r++)
      for (int32 c = 0; c < input_dim_; c++) (*linearity)(r, c) = flat[(size_t)r * input_dim_ + c];
    for (int32 r = 0; r < output_dim_; r++) (*bias)(r) = flat[(size_t)output_dim_ * input_dim_ + r];
  }

  /// Affine + Softmax + Xent::EvalMasked.  posteriors: optional [frames x output_dim] soft-max outputs (nnet_out).
  void PropagateEval(const CuMatrixBase<BaseFloat>& in, const VectorBase<BaseFloat>& frame_mask_host, const Posterior& post,
                     CuMatrix<BaseFloat>* posteriors = NULL) {
    const int32 num_frames = in.NumRows();
    KALDI_ASSERT(in.NumCols() == input_dim_);
    KALDI_ASSERT(num_frames == static_cast<int32>(post.size()));        // nnet-loss.cc:80
    KALDI_ASSERT(frame_mask_host.Dim() == num_frames);
    KALDI_ASSERT(num_frames <= max_frames_);
    row_ptr_.assign(num_frames + 1, 0);
    pdf_.clear();
    weight_.clear();
    for (int32 t = 0; t < num_frames; t++) {
      for (size_t i = 0; i < post[t].size(); i++) {
        pdf_.push_back(post[t][i].first);
        weight_.push_back(post[t][i].second);
      }
      row_ptr_[t + 1] = static_cast<int32>(pdf_.size());
    }
    if (posteriors) posteriors->Resize(num_frames, output_dim_, kUndefined);
    Check(lstmp_b200_tail_propagate_eval(engine_, in.Data(), in.Stride(), num_frames, frame_mask_host.Data(),
                                         row_ptr_.data(), pdf_.empty() ? NULL : pdf_.data(),
                                         weight_.empty() ? NULL : weight_.data(), posteriors ? posteriors->Data() : NULL,
                                         posteriors ? posteriors->Stride() : 0, NULL));
  }

  /// in_diff (nullable) = diff * linearity_; the fresh weight / bias gradients stay inside until Update().
  void Backpropagate(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* in_diff) {
    if (in_diff) in_diff->Resize(in.NumRows(), input_dim_, kUndefined);
    Check(lstmp_b200_tail_backpropagate(engine_, in.Data(), in.Stride(), in_diff ? in_diff->Data() : NULL,
                                        in_diff ? in_diff->Stride() : 0, in.NumRows(), NULL));
  }

  void Update() {
    if (comm_) Check(lstmp_b200_tail_allreduce_grads_nccl(engine_, comm_, NULL));  // data-parallel: sum fresh gradients
    Check(lstmp_b200_tail_update(engine_, opts_.learn_rate, opts_.momentum, NULL));
  }
  void SetNcclComm(void* comm) { comm_ = comm; }

  /// Xent::Report (nnet-loss.cc:293-307, without the progress vector)
  std::string Report() const {
    lstmp_b200_xent_stats_t s;
    Check(lstmp_b200_tail_get_stats(engine_, &s, NULL));
    const double f = s.frames > 0 ? static_cast<double>(s.frames) : 1.0;
    std::ostringstream oss;
    oss << "AvgLoss: " << (s.loss - s.entropy) / f << " (Xent), "
        << "[AvgXent: " << s.loss / f << ", AvgTargetEnt: " << s.entropy / f << "]" << std::endl;
    oss << "\nFRAME_ACCURACY >> " << 100.0 * s.correct / f << "% <<";
    return oss.str();
  }
  lstmp_b200_xent_stats_t Stats() const {
    lstmp_b200_xent_stats_t s;
    Check(lstmp_b200_tail_get_stats(engine_, &s, NULL));
    return s;
  }
  lstmp_b200_tail_handle_t Engine() const { return engine_; }

 private:
  static void Check(int rc) {
    if (rc != 0) KALDI_ERR << "lstmp_b200 tail: " << lstmp_b200_last_error() << " (code " << rc << ")";
  }
  int32 input_dim_, output_dim_, max_frames_;
  lstmp_b200_tail_handle_t engine_;
  NnetTrainOptions opts_;
  void* comm_ = NULL;
  std::vector<int32> row_ptr_, pdf_;
  std::vector<BaseFloat> weight_;
};

}  // namespace nnet1
}  // namespace kaldi
#endif
