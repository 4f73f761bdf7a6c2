// kaldi/b200-stream-dispatch.h
//
// The multi-stream dispatch of the reference trainer (google/nnetbin/bd-nnet-train-lstm-streams.cc:128-212) as a class:
// the same book-keeping (keys / feats / targets / curt / lent / new_utt_flags, :128-134), the same refill rule
// (:146-174: skip utterances without targets or with a length mismatch), the same termination test (:177-181) and
// the same batch fill (:187-206) -- except that the FEATURE part of the fill, the chunk's host->device copy (:212) and
// the AddShift + Rescale feature transform run on the GPU behind the C ABI (lstmp_b200_dispatch_*, include/lstmp_b200.h):
// an utterance crosses PCIe once, when its stream takes it; per chunk 3*S ints go down and one gather kernel writes
// the [T*S x D] CuMatrix the network reads.  frame_mask / target / new_utt_flags are host objects as in the reference.
//
// FeatureReader / TargetReader are the Kaldi table-reader concepts the trainer uses
// (SequentialBaseFloatMatrixReader: Done(), Key(), Value(), Next(); RandomAccessPosteriorReader: HasKey(), Value(key)).
#ifndef B200_KALDI_STREAM_DISPATCH_H_
#define B200_KALDI_STREAM_DISPATCH_H_

#ifdef HAVE_KALDI
#include "cudamatrix/cu-matrix.h"
#include "hmm/posterior.h"
#else
#include "compat/kaldi-compat.h"
#endif
#include <string>
#include <vector>

#include "lstmp_b200.h"

namespace kaldi {
namespace nnet1 {

class B200StreamDispatch {
 public:
  B200StreamDispatch(int32 num_stream, int32 batch_size, int32 targets_delay, int32 feat_dim, int32 max_utt_frames = 4096)
      : S_(num_stream), T_(batch_size), delay_(targets_delay), D_(feat_dim), engine_(NULL), keys_(num_stream),
        feats_(num_stream), targets_(num_stream), curt_(num_stream, 0), lent_(num_stream, 0),
        new_utt_flags_(num_stream, 0), num_done_(0), num_no_tgt_mat_(0), num_other_error_(0) {
    int dev = 0;
    CU_SAFE_CALL(cudaGetDevice(&dev));
    Check(lstmp_b200_dispatch_create(S_, T_, delay_, D_, max_utt_frames, dev, &engine_));
  }
  ~B200StreamDispatch() { lstmp_b200_dispatch_destroy(engine_); }
  B200StreamDispatch(const B200StreamDispatch&) = delete;
  B200StreamDispatch& operator=(const B200StreamDispatch&) = delete;

  /// <AddShift> / <Rescale> of the feature transform (google/feature_transform.nnet.txt:2-5); NULL = absent.
  void SetTransform(const Vector<BaseFloat>* shift, const Vector<BaseFloat>* scale) {
    KALDI_ASSERT(!shift || shift->Dim() == D_);
    KALDI_ASSERT(!scale || scale->Dim() == D_);
    Check(lstmp_b200_dispatch_set_transform(engine_, shift ? shift->Data() : NULL, scale ? scale->Data() : NULL));
  }

  /// One iteration of the trainer's while(1) loop up to (not including) nnet.Reset (:143-206 + :212 transform).
  /// Returns false when every stream is exhausted (:177-181).  feat is resized to [T*S x D] on the device.
  template <class FeatureReader, class TargetReader>
  bool NextChunk(FeatureReader* feature_reader, TargetReader* target_reader, CuMatrix<BaseFloat>* feat,
                 Vector<BaseFloat>* frame_mask, Posterior* target, std::vector<int32>* new_utt_flags,
                 cudaStream_t stream = 0) {
    for (int32 s = 0; s < S_; s++) {                       // :146
      if (curt_[s] < lent_[s]) {                           // :148
        new_utt_flags_[s] = 0;
        continue;
      }
      while (!feature_reader->Done()) {                    // :153
        // One deliberate difference: the reference assigns feats[s] / targets[s] BEFORE the two checks (:154-162), so
        // a skipped utterance at the end of the data leaves an exhausted stream padding from the wrong matrix
        // (possibly out of range).  Here a stream keeps its own utterance until it accepts a new one; the rows
        // concerned are padding (mask 0, no gradient).
        const std::string key = feature_reader->Key();
        if (!target_reader->HasKey(key)) {                 // :156
          KALDI_WARN << key << ", missing targets";
          num_no_tgt_mat_++;
          feature_reader->Next();
          continue;
        }
        if (feature_reader->Value().NumRows() != static_cast<int32>(target_reader->Value(key).size())) {   // :163
          KALDI_WARN << key << ", length miss-match between feats and targets, skip";
          num_other_error_++;
          feature_reader->Next();
          continue;
        }
        keys_[s] = key;
        feats_[s] = feature_reader->Value();
        targets_[s] = target_reader->Value(key);
        curt_[s] = 0;                                      // :168
        lent_[s] = feats_[s].NumRows();
        new_utt_flags_[s] = 1;                             // :170
        KALDI_ASSERT(feats_[s].NumCols() == D_);
        Check(lstmp_b200_dispatch_load_utt(engine_, s, feats_[s].Data(), feats_[s].Stride(), feats_[s].NumRows()));
        num_done_++;
        feature_reader->Next();
        break;
      }
    }
    int done = 1;                                          // :177
    for (int32 s = 0; s < S_; s++)
      if (curt_[s] < lent_[s]) done = 0;
    if (done) return false;

    // feature part of the fill + H2D + transform: one gather kernel on the device       (:198-202, :212)
    feat->Resize(T_ * S_, D_, kUndefined);
    Check(lstmp_b200_dispatch_assemble(engine_, curt_.data(), lent_.data(), feat->Data(), feat->Stride(), stream));
    // frame_mask & targets padding stay on the host                                       (:190-196)
    if (frame_mask->Dim() != T_ * S_) frame_mask->Resize(T_ * S_);
    target->resize(T_ * S_);
    for (int32 t = 0; t < T_; t++) {
      for (int32 s = 0; s < S_; s++) {
        if (lent_[s] == 0) {  // a stream that never got an utterance: the reference would index targets[s][-1]
          (*frame_mask)(t * S_ + s) = 0;
          (*target)[t * S_ + s].clear();
        } else if (curt_[s] < lent_[s]) {
          (*frame_mask)(t * S_ + s) = 1;
          (*target)[t * S_ + s] = targets_[s][curt_[s]];
        } else {
          (*frame_mask)(t * S_ + s) = 0;
          (*target)[t * S_ + s] = targets_[s][lent_[s] - 1];
        }
        curt_[s]++;                                        // :204
      }
    }
    *new_utt_flags = new_utt_flags_;
    return true;
  }

  int32 NumDone() const { return num_done_; }
  int32 NumNoTgtMat() const { return num_no_tgt_mat_; }
  int32 NumOtherError() const { return num_other_error_; }
  lstmp_b200_dispatch_stats_t Stats() const {
    lstmp_b200_dispatch_stats_t s;
    Check(lstmp_b200_dispatch_get_stats(engine_, &s));
    return s;
  }

 private:
  static void Check(int rc) {
    if (rc != 0) KALDI_ERR << "lstmp_b200 dispatch error " << rc << ": " << lstmp_b200_last_error();
  }
  int32 S_, T_, delay_, D_;
  lstmp_b200_dispatch_handle_t engine_;
  std::vector<std::string> keys_;
  std::vector<Matrix<BaseFloat> > feats_;
  std::vector<Posterior> targets_;
  std::vector<int32> curt_, lent_, new_utt_flags_;
  int32 num_done_, num_no_tgt_mat_, num_other_error_;
};

/// TimeShift::PropagateFnc on the device (standard/nnet/nnet-time-shift.h:42-51).
inline void B200TimeShiftPropagate(const CuMatrixBase<BaseFloat>& in, int32 shift, CuMatrixBase<BaseFloat>* out,
                                   cudaStream_t stream = 0) {
  KALDI_ASSERT(in.NumRows() == out->NumRows() && in.NumCols() == out->NumCols());
  if (lstmp_b200_time_shift(in.Data(), in.Stride(), out->Data(), out->Stride(), in.NumRows(), in.NumCols(), shift,
                            stream) != 0)
    KALDI_ERR << "lstmp_b200_time_shift: " << lstmp_b200_last_error();
}

}  // namespace nnet1
}  // namespace kaldi
#endif
