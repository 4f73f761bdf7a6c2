// A miniature of the reference trainer's main loop (google/nnetbin/bd-nnet-train-lstm-streams.cc:143-229) written
// against the three Kaldi-surface classes of this repository, C++ only:
//
//   B200StreamDispatch      (keys / feats / targets / curt / lent / new_utt_flags; device-side chunk assembly + CMVN)
//   B200LstmProjectedStreams (nnet.Reset / Propagate / Backpropagate / Update of the LSTM layer)
//   B200AffineSoftmaxXent    (AffineTransform + Softmax + Xent::EvalMasked + their backward / update)
//
// Synthetic data with a learnable rule (the target of a frame is a function of the feature frame `targets_delay`
// frames later is NOT used here; the target depends on the CURRENT utterance id), two epochs: checks that every valid
// frame is counted exactly once per epoch, that cross-validation (no Backpropagate) leaves the parameters untouched,
// and that the training loss goes down (the same set-up run through the CPU oracles goes 2.53 -> 1.57 per frame in four epochs).
#include <cmath>
#include <cstdio>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "b200-affine-softmax-xent.h"
#include "b200-lstm-projected-streams.h"
#include "b200-stream-dispatch.h"

using namespace kaldi;
using namespace kaldi::nnet1;

struct Utt {
  std::string key;
  Matrix<BaseFloat> feats;
};
struct SeqReader {
  const std::vector<Utt>* u;
  size_t i;
  bool Done() const { return i >= u->size(); }
  const std::string& Key() const { return (*u)[i].key; }
  const Matrix<BaseFloat>& Value() const { return (*u)[i].feats; }
  void Next() { ++i; }
};
struct TgtReader {
  std::map<std::string, Posterior> m;
  bool HasKey(const std::string& k) const { return m.count(k) != 0; }
  const Posterior& Value(const std::string& k) const { return m.at(k); }
};

int main() {
  try {
    const int32 num_stream = 4, batch_size = 10, targets_delay = 2, D = 8, R = 16, P = 12;
    const int NU = 24;
    std::mt19937 rng(11);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<Utt> utts(NU);
    TgtReader tr;
    long valid_frames = 0;
    for (int i = 0; i < NU; i++) {
      const int L = 15 + (int)(rng() % 40), cls = i % P;
      utts[i].key = "utt" + std::to_string(i);
      utts[i].feats.Resize(L, D);
      Posterior p(L);
      for (int r = 0; r < L; r++) {
        for (int d = 0; d < D; d++) utts[i].feats(r, d) = 0.3f * nd(rng) + ((d == cls % D) ? 2.0f : 0.f) + (cls >= D ? 1.0f : 0.f);
        p[r].push_back(std::make_pair((int32)cls, 1.0f));
      }
      tr.m[utts[i].key] = p;
      valid_frames += L;
    }
    B200LstmProjectedStreams lstm(D, R);
    {
      std::istringstream cfg("<CellDim> 32 <NumStream> 4 <ParamScale> 0.1");
      lstm.InitData(cfg);
    }
    B200AffineSoftmaxXent tail(R, P, batch_size * num_stream);
    {
      Matrix<BaseFloat> W(P, R);
      Vector<BaseFloat> b(P);
      for (int r = 0; r < P; r++)
        for (int c = 0; c < R; c++) W(r, c) = 0.1f * nd(rng);
      tail.SetParams(W, b);
    }
    NnetTrainOptions opts;
    opts.learn_rate = 0.002f;   // gradients are SUMS over the 40 frames of a chunk (LPS.h:468-487): 0.02 diverges
    opts.momentum = 0.9f;
    lstm.SetTrainOptions(opts);
    tail.SetTrainOptions(opts);
    Vector<BaseFloat> shift(D), scale(D);
    for (int d = 0; d < D; d++) { shift(d) = -0.5f; scale(d) = 0.8f; }

    double first_epoch_loss = 0, last_epoch_loss = 0;
    Vector<BaseFloat> params_before_cv;
    for (int epoch = 0; epoch < 5; epoch++) {
      const bool crossvalidate = (epoch == 4);
      if (crossvalidate) lstm.GetParams(&params_before_cv);
      B200StreamDispatch dispatch(num_stream, batch_size, targets_delay, D, 64);
      dispatch.SetTransform(&shift, &scale);
      SeqReader feature_reader{&utts, 0};
      CuMatrix<BaseFloat> feat_transf, lstm_out, lstm_out_diff;
      Vector<BaseFloat> frame_mask;
      Posterior target;
      std::vector<int32> new_utt_flags;
      lstmp_b200_tail_reset_stats(tail.Engine(), NULL);
      // a fresh epoch starts every stream from zero history
      std::vector<int> all(num_stream, 1);
      lstm.Reset(all);
      while (dispatch.NextChunk(&feature_reader, &tr, &feat_transf, &frame_mask, &target, &new_utt_flags)) {
        lstm.Reset(new_utt_flags);                                   // nnet.Reset(new_utt_flags)           :209
        lstm.Propagate(feat_transf, &lstm_out);                      // nnet.Propagate                      :215
        tail.PropagateEval(lstm_out, frame_mask, target);            // ... + xent.EvalMasked               :219
        if (!crossvalidate) {                                        // nnet.Backpropagate(obj_diff, NULL)  :228
          tail.Backpropagate(lstm_out, &lstm_out_diff);
          lstm.Backpropagate(feat_transf, lstm_out, lstm_out_diff, NULL);
          tail.Update();
          lstm.Update(feat_transf, lstm_out_diff);
        }
      }
      lstmp_b200_xent_stats_t st = tail.Stats();
      printf("epoch %d (%s): %s\n", epoch, crossvalidate ? "CROSS-VALIDATION" : "TRAINING", tail.Report().c_str());
      if (st.frames != valid_frames) { printf("FAIL: %lld frames counted, %ld valid\n", (long long)st.frames, valid_frames); return 1; }
      if (dispatch.NumDone() != NU) { printf("FAIL: %d utterances done\n", dispatch.NumDone()); return 1; }
      if (epoch == 0) first_epoch_loss = st.loss / st.frames;
      if (epoch == 3) last_epoch_loss = st.loss / st.frames;
      if (crossvalidate) {
        Vector<BaseFloat> after;
        lstm.GetParams(&after);
        for (int32 i = 0; i < after.Dim(); i++)
          if (after(i) != params_before_cv(i)) { printf("FAIL: cross-validation changed the parameters\n"); return 1; }
        if (!(st.loss / st.frames < first_epoch_loss)) { printf("FAIL: CV loss %g vs first epoch %g\n", st.loss / st.frames, first_epoch_loss); return 1; }
      }
    }
    if (!(last_epoch_loss < 0.85 * first_epoch_loss) || !std::isfinite(last_epoch_loss)) {
      printf("FAIL: loss did not go down: %g -> %g\n", first_epoch_loss, last_epoch_loss);
      return 1;
    }
    printf("loss per frame %g -> %g\nPASS\n", first_epoch_loss, last_epoch_loss);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "trainer_test: %s\n", e.what());
    return 2;
  }
}
