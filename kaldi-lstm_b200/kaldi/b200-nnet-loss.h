// kaldi/b200-nnet-loss.h
//
// Drop-in replacement for the masked cross-entropy the multi-stream trainer uses
// (kaldi::nnet1::Xent::EvalMasked + Report, google/nnet/nnet-loss.cc:76-164, :293-307; called from
// google/nnetbin/bd-nnet-train-lstm-streams.cc:219): same signature, the arithmetic behind the C ABI of
// include/lstmp_b200.h (lstmp_b200_xent_*).  The Posterior is flattened to CSR on the host (a few bytes per frame)
// instead of the reference's dense [frames x num_pdf] host matrix + H2D copy; the statistics stay on the device until
// Report() / the accessors are called.  No CPU fallback.
#ifndef B200_KALDI_NNET_LOSS_H_
#define B200_KALDI_NNET_LOSS_H_

#ifdef HAVE_KALDI
#include "cudamatrix/cu-matrix.h"
#include "hmm/posterior.h"
#else
#include "compat/kaldi-compat.h"
#endif
#include <sstream>
#include <string>
#include <vector>

#include "lstmp_b200.h"

namespace kaldi {
namespace nnet1 {

class B200Xent {
 public:
  B200Xent() : engine_(NULL), max_frames_(0) {}
  ~B200Xent() { lstmp_b200_xent_destroy(engine_); }
  B200Xent(const B200Xent&) = delete;
  B200Xent& operator=(const B200Xent&) = delete;

  /// Evaluate cross entropy with frames masked (nnet-loss.h:50-54): diff = frame_mask * (net_out - target);
  /// loss_, entropy_, correct_, frames_ accumulate (on the device).
  void EvalMasked(const VectorBase<BaseFloat>& frame_mask_host, const CuMatrixBase<BaseFloat>& net_out,
                  const Posterior& post, CuMatrix<BaseFloat>* diff) {
    const int32 num_frames = net_out.NumRows(), num_pdf = net_out.NumCols();
    KALDI_ASSERT(num_frames == static_cast<int32>(post.size()));        // nnet-loss.cc:80
    KALDI_ASSERT(frame_mask_host.Dim() == num_frames);
    if (!engine_ || num_frames > max_frames_) {
      if (engine_) {  // keep the statistics across the re-creation
        lstmp_b200_xent_stats_t s;
        Check(lstmp_b200_xent_get_stats(engine_, &s, NULL));
        carried_.loss += s.loss; carried_.entropy += s.entropy; carried_.correct += s.correct; carried_.frames += s.frames;
        lstmp_b200_xent_destroy(engine_);
        engine_ = NULL;
      }
      int dev = 0;
      CU_SAFE_CALL(cudaGetDevice(&dev));
      Check(lstmp_b200_xent_create(num_frames, dev, &engine_));
      max_frames_ = num_frames;
    }
    // Posterior -> CSR                                                   (replaces nnet-loss.cc:82-96)
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
    diff->Resize(num_frames, num_pdf, kUndefined);                       // fully written by the kernel
    // the default (NULL) stream Kaldi's CuDevice uses: ordered with the neighbouring components' work
    Check(lstmp_b200_xent_eval_masked(engine_, frame_mask_host.Data(), DevPtr(net_out), net_out.Stride(), num_frames,
                                      num_pdf, row_ptr_.data(), pdf_.empty() ? NULL : pdf_.data(),
                                      weight_.empty() ? NULL : weight_.data(), DevPtr(*diff), diff->Stride(), NULL));
  }

  struct Stats { double loss, entropy; long long correct, frames; };
  Stats GetStats() {
    Stats r = {carried_.loss, carried_.entropy, carried_.correct, carried_.frames};
    if (engine_) {
      lstmp_b200_xent_stats_t s;
      Check(lstmp_b200_xent_get_stats(engine_, &s, NULL));
      r.loss += s.loss; r.entropy += s.entropy; r.correct += s.correct; r.frames += s.frames;
    }
    return r;
  }

  /// Generate string with error report (nnet-loss.cc:293-307, without the progress vector)
  std::string Report() {
    Stats s = GetStats();
    std::ostringstream oss;
    oss << "AvgLoss: " << (s.loss - s.entropy) / s.frames << " (Xent), "
        << "[AvgXent: " << s.loss / s.frames << ", AvgTargetEnt: " << s.entropy / s.frames << "]" << std::endl;
    oss << "\nFRAME_ACCURACY >> " << 100.0 * s.correct / s.frames << "% <<";
    return oss.str();
  }

 private:
  static void Check(int rc) {
    if (rc != 0) KALDI_ERR << "lstmp_b200 error " << rc << ": " << lstmp_b200_last_error();
  }
  template <class M>
  static BaseFloat* DevPtr(const M& m) {
#ifdef B200_CUMATRIX_DATA_VIA_ROW
    return const_cast<BaseFloat*>(m.NumRows() ? m.Row(0).Data() : NULL);
#else
    return const_cast<BaseFloat*>(m.Data());
#endif
  }
  lstmp_b200_xent_handle_t engine_;
  int32 max_frames_;
  Stats carried_ = {0.0, 0.0, 0, 0};
  std::vector<int32> row_ptr_, pdf_;
  std::vector<BaseFloat> weight_;
};

}  // namespace nnet1
}  // namespace kaldi
#endif  // B200_KALDI_NNET_LOSS_H_
